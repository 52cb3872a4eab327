"""Times ck_net_forward_device (tower + heads) on a resident batch with CUDA events; reports the median
over iterations, L2 flushed between them.  Usage: python scripts/time_tower.py [n] [iters]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rng = np.random.RandomState(0)
pos = np.zeros(n, dtype=L.POS_DTYPE)
pos["p1"], pos["p2"] = 0x00000FFF, 0xFFF00000
for _ in range(16):
    out = L.movegen(pos)
    pick = (rng.rand(n) * np.maximum(out["counts"], 1)).astype(np.int64)
    nxt = out["children"][np.arange(n), pick]
    alive = (out["status"] == 0) & (out["counts"] > 0)
    pos = np.where(alive, nxt, pos)
out = L.movegen(pos, want_children=False)
leaves = np.zeros(n, dtype=L.LEAF_DTYPE)
leaves["p1"], leaves["p2"], leaves["k"] = pos["p1"], pos["p2"], pos["k"]
leaves["info"] = (pos["meta"] & 1)
leaves["mask"] = out["masks"]
net = L.Net(0, "tc")
net.set_weights(N.random_init_blob(0))
lib = L.raw()
d_leaves = torch.from_numpy(leaves.view(np.uint8).reshape(n, -1).copy()).cuda()
d_pol = torch.empty((n, 512), dtype=torch.float32, device="cuda")
d_val = torch.empty(n, dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def launch():
    L.check(lib.ck_net_forward_device(net._h, C.c_void_p(d_leaves.data_ptr()), n, C.c_void_p(d_pol.data_ptr()),
                                      C.c_void_p(d_val.data_ptr()), C.c_void_p(stream)))


for _ in range(3):
    launch()
torch.cuda.synchronize()
ts = []
for i in range(iters):
    if os.environ.get("FLUSH", "0") == "1":
        flush.fill_(i & 255)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    launch()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
# back-to-back (sustained clocks)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    launch()
e1.record()
torch.cuda.synchronize()
print("n=%d forward (tower+heads): median %.4f ms, min %.4f ms; 200 back-to-back: %.4f ms each" %
      (n, float(np.median(ts)), min(ts), e0.elapsed_time(e1) / 200))
