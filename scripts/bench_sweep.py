"""cfg4 of BASELINE.json: move-gen / rollout-only sweep (no network), 1 k - 1 M concurrent
bitboard positions; positions/s and playouts/s against the HBM roofline, with the CPU oracle
timed on the host next to it.  Prints one JSON line per size.

    python scripts/bench_sweep.py [--max-log2 20]
Positions are drawn by playing k ~ U[0,60] uniformly random legal plies from the start
(SURVEY 8d cfg4), generated on the GPU with the library's own movegen kernel and a fixed seed.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from ckb200 import lib as L  # noqa: E402


def make_positions(n, seed=0):
    rng = np.random.RandomState(seed)
    pos = np.zeros(n, dtype=L.POS_DTYPE)
    pos["p1"], pos["p2"] = 0x00000FFF, 0xFFF00000
    plies = rng.randint(0, 61, size=n)
    for k in range(60):
        out = L.movegen(pos)
        pick = (rng.rand(n) * np.maximum(out["counts"], 1)).astype(np.int64)
        nxt = out["children"][np.arange(n), pick]
        go = (out["status"] == 0) & (out["counts"] > 0) & (plies > k)
        pos = np.where(go, nxt, pos)
    return pos


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=20)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    L.require_device()
    lib = L.raw()
    peaks = None
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks["hbm_gbs"] if peaks else 6650.0
    full = make_positions(1 << args.max_log2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > L2 (126 MB)
    stream = torch.cuda.current_stream().cuda_stream
    for lg in range(10, args.max_log2 + 1, 2):
        n = 1 << lg
        pos = full[:n]
        d_pos = torch.from_numpy(pos.view(np.uint32).reshape(n, 4).copy()).cuda()
        d_children = torch.zeros((n, 48, 4), dtype=torch.int32, device="cuda")
        d_counts = torch.zeros(n, dtype=torch.int32, device="cuda")
        d_masks = torch.zeros((n, 8), dtype=torch.int32, device="cuda")
        d_status = torch.zeros(n, dtype=torch.uint8, device="cuda")
        d_p5 = torch.zeros(n, dtype=torch.uint8, device="cuda")

        def launch():
            L.check(lib.ck_movegen_device(C.c_void_p(d_pos.data_ptr()), n, 48, C.c_void_p(d_children.data_ptr()),
                                          C.c_void_p(d_counts.data_ptr()), C.c_void_p(d_masks.data_ptr()),
                                          C.c_void_p(d_status.data_ptr()), C.c_void_p(d_p5.data_ptr()), C.c_void_p(stream)))
        for _ in range(3):
            launch()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.iters):
            flush.fill_(1)                                                   # L2 flush between timed iterations
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            launch()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms_strided = float(np.median(times))
        b = float(d_counts.float().mean().item())
        bytes_alg = n * (16 + 4 + 32 + 1 + 1 + 16 * b)
        # packed (CSR) output: the layout the roofline is quoted on
        total = int(d_counts.sum().item())
        d_packed = torch.zeros((max(total, 1), 4), dtype=torch.int32, device="cuda")
        d_off = torch.zeros(n + 1, dtype=torch.int32, device="cuda")

        def launch_csr():
            L.check(lib.ck_movegen_csr_device(C.c_void_p(d_pos.data_ptr()), n, C.c_void_p(d_packed.data_ptr()), total,
                                              C.c_void_p(d_off.data_ptr()), C.c_void_p(d_masks.data_ptr()),
                                              C.c_void_p(d_status.data_ptr()), C.c_void_p(d_p5.data_ptr()), C.c_void_p(stream)))
        for _ in range(3):
            launch_csr()
        torch.cuda.synchronize()
        assert int(d_off[-1].item()) == total
        times = []
        for _ in range(args.iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            launch_csr()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        # random playouts (K4): host call once (results + end-to-end time), kernel timed on the device
        t0 = time.time()
        outcome, plies = L.rollout(pos, seed=1)
        roll_s = time.time() - t0
        d_outc = torch.zeros(n, dtype=torch.uint8, device="cuda")
        d_pl = torch.zeros(n, dtype=torch.int32, device="cuda")

        def launch_rollout():
            L.check(lib.ck_rollout_device(C.c_void_p(d_pos.data_ptr()), n, 1, 0, C.c_void_p(d_outc.data_ptr()),
                                          C.c_void_p(d_pl.data_ptr()), C.c_void_p(stream)))
        launch_rollout()
        torch.cuda.synchronize()
        assert (d_pl.cpu().numpy() == plies).all() and (d_outc.cpu().numpy() == outcome).all()
        times = []
        for _ in range(max(3, args.iters // 4)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            launch_rollout()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        roll_ms = float(np.median(times))
        line = {"workload": "cfg4 movegen/rollout sweep", "positions": n, "movegen_ms": ms, "movegen_strided_ms": ms_strided,
                "layout": "packed CSR (ck_movegen_csr_device; timing includes the workspace memset); strided = [n][48] ck_movegen_device",
                "positions_per_sec": n / (ms / 1e3), "mean_children": b,
                "roofline": {"bound": "hbm", "achieved": bytes_alg / (ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s",
                             "frac": bytes_alg / (ms / 1e3) / 1e9 / hbm, "alg_bytes_per_position": bytes_alg / n},
                "playout_kernel_ms": roll_ms, "playouts_per_sec": n / (roll_ms / 1e3),
                "plies_per_sec": float(plies.sum()) / (roll_ms / 1e3),
                "playouts_per_sec_e2e": n / roll_s, "playout_plies_mean": float(plies.mean()),
                "plies_per_sec_e2e": float(plies.sum()) / roll_s}
        if lg == 10:
            from oracle import oracle as O
            t0 = time.time()
            k = 0
            while time.time() - t0 < 3.0:
                for p in pos[:256]:
                    O.movegen(tuple(int(v) for v in p))
                k += 256
            line["cpu_oracle_positions_per_sec_1core"] = k / (time.time() - t0)
        print(json.dumps(line))


if __name__ == "__main__":
    main()
