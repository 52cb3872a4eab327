"""K3 on the reference's own trained networks: runs the committed weight blobs (tests/golden/net_model{10,5}.npz,
imported from data/model/Checkers_Model{10,5}_*.h5 by ckb200.h5lite) through the device network and compares
policy LOGITS, pre-tanh value, softmax and tanh outputs with the float64 restatement's values stored next to
them.  The tower is selected by the environment like everywhere else (CK_TOWER=ss, CK_TS_TILES=1|2,
CK_HEADS=simt); `--impl simt` runs the fp32 CUDA-core path.  Prints one JSON line per model.

    python scripts/check_trained.py [--impl tc|simt] [--tol 1e-5]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
from ckb200 import lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="tc")
ap.add_argument("--tol", type=float, default=1e-5)
args = ap.parse_args()
bad = 0
for it in (10, 5):
    g = np.load(os.path.join(ROOT, "tests", "golden", "net_model%d.npz" % it))
    leaves = np.zeros(len(g["leaves"]), dtype=L.LEAF_DTYPE)
    u = g["leaves"]
    leaves["p1"], leaves["p2"], leaves["k"], leaves["info"] = u[:, 0], u[:, 1], u[:, 2], u[:, 3]
    leaves["mask"] = u[:, 4:12]
    net = L.Net(0, args.impl)
    net.set_weights(g["blob"])
    errs = {}
    for n in (len(leaves), 97, 3):                       # two tiles per CTA, one tile per CTA, a ragged tail
        pol, val, logits, vpre = net.forward_logits(leaves[:n])
        e = dict(logits=float(np.abs(logits - g["logits"][:n]).max()), value_pre=float(np.abs(vpre - g["value_pre"][:n]).max()),
                 policy=float(np.abs(pol - g["policy"][:n]).max()), value=float(np.abs(val - g["value"][:n]).max()))
        p2, v2 = net.forward(leaves[:n])
        assert p2.tobytes() == pol.tobytes() and v2.tobytes() == val.tobytes()
        for k, v in e.items():
            errs[k] = max(errs.get(k, 0.0), v)
    ok = all(v < args.tol for v in errs.values())
    bad += not ok
    print(json.dumps(dict(model=str(g["source"]), impl=args.impl, tower=os.environ.get("CK_TOWER", "ts"),
                          tiles=os.environ.get("CK_TS_TILES", "auto"), heads=os.environ.get("CK_HEADS", "tc"),
                          positions=len(leaves), max_abs_logit=float(np.abs(g["logits"]).max()), max_err=errs,
                          tol=args.tol, ok=bool(ok))), flush=True)
    net.close()
sys.exit(1 if bad else 0)
