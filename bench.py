#!/usr/bin/env python3
"""bench.py -- MCTS simulations/s of batched Checkers self-play on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: oracle port on the host cores

Workload (BASELINE.json configs[1]): 4096 concurrent self-play games per GPU, 400 sims/move,
UCT_C=4, Dirichlet(alpha=1, eps=0.25) at every node, tau=1 with decay, TERMINATE_CNT=200,
random-init network (seed 0), games start from the initial position and are refilled when they
finish.  One "step" = 400 lock-step rounds (tree kernel + batched network evaluation), i.e. about
one move of every game.  `value` = simulations completed / device time (CUDA events on the engine's
stream), inputs resident in HBM.  `e2e` = the same metric through the host-buffer API: every step
uploads the weight blob from pinned host memory and reads the finished games' records back.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "checkers-mcts_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "mcts_sims_per_sec"
UNIT = "sims/s"
SLOTS = 4096
BUDGET = 400
ROUNDS_PER_STEP = 400
TOWER_FLOP_PER_POS = 2 * (9 * 14 * 128 * 64 + 7 * 9 * 128 * 128 * 64 + 128 * 8 * 64)   # 134,316,032: the eight 3x3 convs + the policy conv1x1 the tower kernel evaluates
NET_FLOP_PER_POS = 134865024                                               # SURVEY 8(d), whole network
MCTS = dict(uct_c=4.0, training=True, alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10,
            terminate_cnt=200)


def workload_config(n_gpus):
    return {"workload": "cfg2: %d concurrent self-play games per GPU, %d sims/move, batched NN eval" % (SLOTS, BUDGET),
            "games_per_gpu": SLOTS, "sims_per_move": BUDGET, "rounds_per_step": ROUNDS_PER_STEP,
            "mcts": "UCT_C=4 alpha=1.0 eps=0.25 tau=1.0 decay=0.1 delay=10 TERMINATE_CNT=200",
            "net": "create_nn 7x[conv3x3(128)+ReLU+BN] + policy/value heads, random init seed 0",
            "sharding": "game g -> rank g mod %d, no data-path collective" % n_gpus,
            "l2": "no flush needed: per-GPU working set (tree pools + activations) exceeds the 126 MB L2"}


# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C restatement of the tree/game loop + PyTorch-CPU restatement of the
# network, one process per host core, batch-1 evaluation like the reference's Keras predict)
def _cpu_worker(args):
    seed, seconds, budget = args
    import torch
    torch.set_num_threads(1)
    from ckb200 import codec
    from ckb200 import net as N
    from oracle import net_oracle as NO
    from oracle import oracle as O
    model = NO.TorchKerasLike(N.random_init_blob(0))

    def ev(pos, mask, plane5):
        x = codec.nn_input_planes(pos, mask, plane5).reshape(1, 8, 8, 14)
        p, v = model.predict(x)
        return p[0], v[0, 0]

    cfg = O.make_cfg(uct_c=4.0, budget=budget, training=True, alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1,
                     tau_decay_delay=10, terminate_cnt=200, seed=seed + 1)
    game = O.Game(cfg, O.python_eval(ev))
    t0 = time.time()
    plies = 0
    while time.time() - t0 < seconds:
        if not game.play_ply():
            s0 = game.total_sims
            game.close()
            game = O.Game(cfg, O.python_eval(ev))
            game._carry = s0
        plies += 1
    el = time.time() - t0
    return game.total_sims + getattr(game, "_carry", 0), el, plies


def usable_cores():
    """worker processes for the CPU arm: every host core, bounded by memory (each worker imports
    torch, ~1.5 GB resident) and by 64 so that the pool start-up stays within the bench budget"""
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    try:
        import psutil
        cores = min(cores, max(1, int(psutil.virtual_memory().available / 1.5e9)))
    except Exception:
        pass
    return max(1, min(cores, 64))


def cpu_port_sample(n_procs, seconds, budget=BUDGET, pool=None):
    """-> (sims/s aggregate, cores, sims, seconds)"""
    import multiprocessing as mp
    if n_procs == 1 and pool is None:
        res = [_cpu_worker((0, seconds, budget))]
    elif pool is not None:
        res = pool.map(_cpu_worker, [(i, seconds, budget) for i in range(n_procs)])
    else:
        with mp.get_context("spawn").Pool(n_procs) as p2:
            res = p2.map(_cpu_worker, [(i, seconds, budget) for i in range(n_procs)])
    sims = sum(r[0] for r in res)
    el = max(r[1] for r in res)
    return sims / el, n_procs, sims, el


def run_reference(args, rank, world):
    """CPU arm: the reference's algorithm (oracle port) on all host cores, one process per core
    like the reference's own mp.Pool.map fan-out (training_pipeline.py:326-329)."""
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import oracle as O
    O.build()
    cores = usable_cores()
    per_step = 6.0
    with mp.get_context("spawn").Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_port_sample(cores, 1.0, pool=pool)
        vals = []
        t0 = time.time()
        for _ in range(args.steps):
            v, c, sims, el = cpu_port_sample(cores, per_step, pool=pool)
            vals.append(v)
        wall = time.time() - t0
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "oracle port (C tree + torch-CPU net, batch-1 eval), %d processes x %.0f s of self-play "
                                       "at %d sims/move per step" % (cores, per_step, BUDGET)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ckb200 import lib as L
    from ckb200 import net as N
    L.require_device()
    torch.cuda.set_device(local_rank)
    dev = local_rank

    blob = N.random_init_blob(0)
    pinned = torch.from_numpy(blob).pin_memory()
    net = L.Net(dev, args.net_impl)
    weights = torch.empty(blob.size, dtype=torch.float32, device="cuda:%d" % dev)
    weights.copy_(pinned)
    net.set_weights_device(weights.data_ptr(), weights.numel())

    from ckb200 import dist as D
    base, stride, _n = D.shard(args.slots * world, rank, world)          # game g -> rank g mod world
    cfg = L.make_cfg(n_slots=args.slots, budget=BUDGET, device=dev, evaluator="net", keep_records=True,
                     seed=20261017, game_id_base=base, game_id_stride=stride, **MCTS)
    eng = L.Engine(cfg)
    eng.set_net(0, net)
    n_games = args.slots * 8                     # staged games: enough refills for any bench length
    eng.begin(n_games)
    eng.set_profile(True)
    rec_buf = np.zeros(args.slots * 201, dtype=L.RECORD_DTYPE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.run(args.rounds)
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()

    # ---- timed region 1: device-resident throughput --------------------------------------------
    barrier()
    t0 = time.time()
    agg = dict(sims=0, nn_evals=0, gpu_ms=0.0, eval_ms=0.0, tower_ms=0.0, kernel_launches=0, games_finished=0, moves=0, steps=0)
    for _ in range(args.steps):
        st = eng.run(args.rounds)
        for k in agg:
            agg[k] += st[k]
    barrier()
    wall = time.time() - t0

    # ---- timed region 2: end to end through host buffers --------------------------------------
    barrier()
    e_t0 = time.time()
    e_sims, e_h2d, e_d2h, e_ms, e_launch = 0, 0, 0, 0.0, 0
    for _ in range(args.steps):
        s0 = time.time()
        weights.copy_(pinned, non_blocking=False)                       # H2D of this step's input (weights)
        net.set_weights_device(weights.data_ptr(), weights.numel())
        st = eng.run(args.rounds)
        nrec, ngames = eng.records_new(rec_buf)                          # D2H of this step's results
        e_sims += st["sims"]
        e_launch += st["kernel_launches"]
        e_h2d += blob.nbytes
        e_d2h += nrec * L.RECORD_DTYPE.itemsize + n_games * L.GAME_DTYPE.itemsize
        torch.cuda.synchronize()
        e_ms += 1000.0 * (time.time() - s0)
    barrier()
    e_wall = time.time() - e_t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- reduce over ranks: time = max, work = sum ----------------------------------------------
    t = torch.tensor([agg["gpu_ms"], e_ms], dtype=torch.float64, device="cuda:%d" % dev)
    w = torch.tensor([agg["sims"], agg["nn_evals"], agg["games_finished"], agg["moves"], e_sims, agg["kernel_launches"] + e_launch],
                     dtype=torch.float64, device="cuda:%d" % dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    gpu_ms, e2e_ms = t.tolist()
    sims, evals, games, moves, e2e_sims, launches = w.tolist()

    # pool finished records on rank 0 (the iteration-end NCCL gather; outside the timed regions)
    gather_ms, pooled = None, None
    if world > 1:
        from ckb200 import dist as D
        recs = eng.records()
        torch.cuda.synchronize()
        g0 = time.time()
        pooled = D.gather_records(recs, rank, world, device="cuda:%d" % dev)
        torch.cuda.synchronize()
        gather_ms = 1000.0 * (time.time() - g0)

    if rank != 0:
        return
    value = sims / (gpu_ms / 1000.0)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": gpu_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (split-fp16 tensor-core passes, fp32 accumulate)" if args.net_impl == "tc" else "f32",
            "data": "synthetic", "config": workload_config(world)}
    line["config"]["net_impl"] = args.net_impl
    line["wall_s"] = wall
    line["nn_evals_per_sec"] = evals / (gpu_ms / 1000.0)
    line["moves_per_sec"] = moves / (gpu_ms / 1000.0)
    line["games_finished"] = games
    line["games_per_sec_est"] = (moves / (gpu_ms / 1000.0)) / 150.0      # at the ~150 plies/game BASELINE.md assumes
    line["e2e"] = {"value": e2e_sims / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": e_h2d // max(args.steps, 1),
                   "d2h_bytes_per_step": e_d2h // max(args.steps, 1), "wall_s": e_wall}
    line["gpu_launches"] = int(launches)
    line["clocks"] = clocks
    if gather_ms is not None:
        line["records_gather_ms"] = gather_ms
        line["records_pooled"] = int(len(pooled))

    # roofline of the dominant kernel (the tcgen05 tower; rank 0's own launches)
    peaks = measured_peaks()
    n_launch = agg["steps"]
    if args.net_impl == "tc" and agg["tower_ms"] > 0:
        peak = peaks["bf16_tflops_sustained"] if peaks else 1400.0
        achieved = agg["nn_evals"] * TOWER_FLOP_PER_POS / (agg["tower_ms"] / 1000.0) / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "tower_ncu_summary.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        line["roofline"] = {"bound": "tensor", "kernel": "tower_ts_kernel" if (os.environ.get("CK_TOWER") or "ts")[0] != "s" else "tower_tc_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak, "traffic": traffic,
                            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400 (of fallback)",
                            "flop_per_position": TOWER_FLOP_PER_POS, "positions_per_launch": agg["nn_evals"] / max(n_launch, 1),
                            "avg_launch_ms": agg["tower_ms"] / max(n_launch, 1),
                            "share_of_step": agg["tower_ms"] / max(agg["gpu_ms"], 1e-9),
                            "issued_mma_tflops": 3 * achieved, "issued_mma_frac_of_peak": 3 * achieved / peak,
                            "note": "achieved/frac count useful FLOPs only (SURVEY 8d); the kernel issues 3 fp16 MMA passes per product "
                                    "(split hi/lo operands) to meet the 1e-5 accuracy contract, so against the pipe's own peak frac is "
                                    "bounded by 1/3; issued_mma_* is the tensor-pipe work actually executed.  The denominator is the "
                                    "power-capped cuBLAS bf16 rate, which this kernel can exceed (it holds a higher clock)"}
    else:
        peak = 75.0
        achieved = agg["nn_evals"] * NET_FLOP_PER_POS / (max(agg["eval_ms"], 1e-9) / 1000.0) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "conv3x3_simt_kernel (CUDA-core cross-check path)", "achieved": achieved,
                            "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                            "peak_source": "nominal fp32 CUDA-core rate; this path is not the product kernel"}

    # CPU baseline (oracle port) on this box's host cores, bounded sample
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        v, c, s, el = cpu_port_sample(1, 12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": c, "kind": "port",
                                "sample": "oracle port (C tree + torch-CPU net, batch-1 eval): 1 process, %.0f s of self-play at %d "
                                          "sims/move (%d sims)" % (el, BUDGET, s)}
    emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner to fd 1)
    are sent to stderr for the whole run and the result line is written to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr


def emit(line):
    out = _REAL_STDOUT or sys.__stdout__
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--net-impl", dest="net_impl", default=os.environ.get("CK_NET_IMPL", "tc"), choices=["tc", "simt"])
    ap.add_argument("--slots", type=int, default=SLOTS)
    ap.add_argument("--rounds", type=int, default=ROUNDS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
