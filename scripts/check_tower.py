"""Compares the tensor-core tower (default: weights-in-TMEM kernel; CK_TOWER=ss: shared-memory kernel)
with the fp32 CUDA-core path on the same leaves, at batch sizes that exercise one tile pair per CTA,
ragged tails and the persistent multi-pair loop.  Usage: python scripts/check_tower.py [sizes...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 5, 97, 592, 593, 1187, 4096, 4099]
nmax = max(sizes)
rng = np.random.RandomState(0)
pos = np.zeros(nmax, dtype=L.POS_DTYPE)
pos["p1"], pos["p2"] = 0x00000FFF, 0xFFF00000
plies = rng.randint(0, 60, size=nmax)
for k in range(60):
    out = L.movegen(pos)
    pick = (rng.rand(nmax) * np.maximum(out["counts"], 1)).astype(np.int64)
    nxt = out["children"][np.arange(nmax), pick]
    alive = (out["status"] == 0) & (out["counts"] > 0) & (plies > k)
    pos = np.where(alive, nxt, pos)
out = L.movegen(pos, want_children=False)
leaves = np.zeros(nmax, dtype=L.LEAF_DTYPE)
leaves["p1"], leaves["p2"], leaves["k"] = pos["p1"], pos["p2"], pos["k"]
leaves["info"] = (pos["meta"] & 1) | (rng.randint(0, 81, size=nmax).astype(np.uint32) << 8)
leaves["mask"] = out["masks"]
blob = N.random_init_blob(1, 0.2)
ref = L.Net(0, "simt"); ref.set_weights(blob)
net = L.Net(0, "tc"); net.set_weights(blob)
bad = 0
for n in sizes:
    rp, rv = ref.forward(leaves[:n])
    t0 = time.time()
    pol, val = net.forward(leaves[:n])
    dt = time.time() - t0
    ep, ev = float(np.abs(pol - rp).max()), float(np.abs(val - rv).max())
    ok = ep < 1e-5 and ev < 1e-5 and np.isfinite(pol).all()
    bad += not ok
    print("n=%5d  max|dpolicy|=%.3e  max|dvalue|=%.3e  %s  (%.1f ms incl. copies)" % (n, ep, ev, "ok" if ok else "MISMATCH", dt * 1e3), flush=True)
sys.exit(1 if bad else 0)
