"""Runs the network forward on a resident batch of leaf positions a few times (profiling target
for `ncu -k regex:tower_tc`).  Usage: python scripts/profile_tower.py [n_positions] [iters] [impl]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
impl = sys.argv[3] if len(sys.argv) > 3 else "tc"
rng = np.random.RandomState(0)
# synthetic random legal-looking positions: random play from the start via the movegen kernel
pos = np.zeros(n, dtype=L.POS_DTYPE)
pos["p1"], pos["p2"] = 0x00000FFF, 0xFFF00000
for _ in range(12):
    out = L.movegen(pos)
    pick = (rng.rand(n) * np.maximum(out["counts"], 1)).astype(np.int64)
    nxt = out["children"][np.arange(n), pick]
    alive = (out["status"] == 0) & (out["counts"] > 0)
    pos = np.where(alive, nxt, pos)
out = L.movegen(pos, want_children=False)
leaves = np.zeros(n, dtype=L.LEAF_DTYPE)
leaves["p1"], leaves["p2"], leaves["k"] = pos["p1"], pos["p2"], pos["k"]
leaves["info"] = (pos["meta"] & 1) | (out["plane5"].astype(np.uint32) << 8)
leaves["mask"] = out["masks"]
net = L.Net(0, impl)
net.set_weights(N.random_init_blob(0))
for _ in range(iters):
    pol, val = net.forward(leaves)
print("ok", pol.shape, float(pol.sum(1).mean()), float(val.mean()))
