"""Device-timed K4 (`ck_rollout_device`): N playouts from the start position or from random mid-game positions.
Usage: python scripts/time_rollout.py [log2_n ...]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from ckb200 import codec, lib as L  # noqa: E402

L.require_device()
lib = L._lib
stream = torch.cuda.current_stream().cuda_stream
for lg in [int(v) for v in sys.argv[1:]] or [12, 16, 20, 22]:
    n = 1 << lg
    start = np.zeros(n, dtype=L.POS_DTYPE)
    for f, v in zip(("p1", "p2", "k", "meta"), codec.start_position()):
        start[f] = v
    d_pos = torch.from_numpy(start.view(np.uint32).reshape(n, 4).view(np.int32)).cuda()
    d_o = torch.zeros(n, dtype=torch.uint8, device="cuda")
    d_p = torch.zeros(n, dtype=torch.int32, device="cuda")
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.ck_rollout_device(C.c_void_p(d_pos.data_ptr()), n, 1, 0, C.c_void_p(d_o.data_ptr()),
                                      C.c_void_p(d_p.data_ptr()), C.c_void_p(stream)))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts[1:])[2]
    pl = d_p.float()
    print({"playouts": n, "kernel_ms": ms, "playouts_per_sec": n / (ms / 1e3), "plies_per_sec": float(pl.sum()) / (ms / 1e3),
           "plies_mean": float(pl.mean()), "plies_max": int(d_p.max()), "outcomes": torch.bincount(d_o.int(), minlength=4).tolist()}, flush=True)
