"""Where does the tensor-core network's error against float64 come from?  For the committed trained weights
(tests/golden/net_model*.npz): error statistics of the tower outputs (pflat, vconv), the logits and the pre-tanh
value, plus the least-squares slope of error on value (a systematic shrink shows up as a negative slope with high
R^2: round-toward-zero accumulation in the tensor pipe).  The float64 features are recomputed here with the test
oracle (CPU, a few seconds).  One JSON line per model; tower / heads variants through CK_TOWER, CK_TS_TILES, CK_HEADS."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
from ckb200 import codec  # noqa: E402
from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402
from oracle import net_oracle as NO  # noqa: E402


def stats(got, ref):
    e = (got.astype(np.float64) - ref).reshape(-1)
    r = ref.reshape(-1)
    slope = float((e * r).sum() / max((r * r).sum(), 1e-300))
    resid = e - slope * r
    return dict(max=float(np.abs(e).max()), rms=float(np.sqrt((e * e).mean())), ref_max=float(np.abs(r).max()),
                slope=slope, r2=float(1 - (resid * resid).sum() / max((e * e).sum(), 1e-300)),
                max_after_slope=float(np.abs(resid).max()))


for it in (10, 5):
    g = np.load(os.path.join(ROOT, "tests", "golden", "net_model%d.npz" % it))
    u = g["leaves"]
    leaves = np.zeros(len(u), dtype=L.LEAF_DTYPE)
    leaves["p1"], leaves["p2"], leaves["k"], leaves["info"], leaves["mask"] = u[:, 0], u[:, 1], u[:, 2], u[:, 3], u[:, 4:12]
    planes = np.stack([codec.nn_input_planes((int(r[0]), int(r[1]), int(r[2]), int(r[3]) & 1), [int(v) for v in r[4:12]], int(r[3]) >> 8) for r in u])
    ref = NO.forward(N.unpack(g["blob"]), planes, features=True)
    out = dict(model=it, tower=os.environ.get("CK_TOWER", "ts"), heads=os.environ.get("CK_HEADS", "tc"), impl=os.environ.get("PROBE_IMPL", "tc"))
    net = L.Net(0, out["impl"])
    net.set_weights(g["blob"])
    pol, val, logits, vpre = net.forward_logits(leaves)
    out["logits"] = stats(logits, ref[2])
    out["value_pre"] = stats(vpre, ref[3])
    out["policy"] = stats(pol, ref[0])
    out["value"] = stats(val, ref[1])
    if out["impl"] == "tc" and out["tower"] != "ss":
        pflat, vconv = net.last_features(len(leaves))
        out["pflat"] = stats(pflat, ref[4])
        out["vconv"] = stats(vconv, ref[5])
    print(json.dumps(out), flush=True)
    net.close()
