#!/usr/bin/env python3
"""bench.py -- MCTS simulations/s and self-play games/s of batched Checkers self-play on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the reference's own code on the host cores

Workload (BASELINE.json configs[1]): 4096 concurrent self-play games per GPU, 400 sims/move, UCT_C=4,
Dirichlet(alpha=1, eps=0.25) at every node, tau=1 with decay, TERMINATE_CNT=200, random-init network (seed 0); slots
are refilled when their game ends.  One "step" = 400 lock-step rounds (tree kernel + batched network evaluation).

Steady state: the slots are desynchronised by a warm start (the first game a slot plays (game g) plays its first hash(g) mod 140 plies
at 8 sims/move, everything after that at 400) and an untimed pre-roll of about one game length at full budget, so the
timed region sees games at every stage with trees and evaluation caches as deep as in a long run, finishing games
included, instead of 4096 openings in lock step.  Then 2K steps alternate:
  even steps -> `value`: simulations / device time (CUDA events on the engine's stream), inputs resident in HBM;
  odd steps  -> `e2e`: the same loop through the host-buffer C ABI, wall clock: the weight blob goes up from pinned host
                memory (H2D), the step runs, the finished games' records and results come back (D2H).
`games_per_sec` = games finished in the even steps / their device time.  `e2e_pipeline` (rank-local, after the timed
region) = wall clock of the drop-in call a user of the reference makes, `generate_Checkers_data(...).generate_data()`
(training_pipeline.py:312-332), for one batch of games played to the end INCLUDING the reference-format pickle files.
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
PREROLL_BUDGET = 8            # sims/move of the opening plies a slot's first game plays before full-budget play (warm start)
PREROLL_PLIES = 140           # game g does that for hash(g) mod 140 plies: stages spread over a typical game length
PREROLL_ROUNDS = 24000        # untimed rounds (~28 s): about one game length at full budget, so trees and evaluation caches are as deep as in a long run
REF_PLIES = 2                 # plies per game of a reference-arm sample (the reference's TERMINATE_CNT knob)
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


def reference_sample(cpus, plies, budget=BUDGET, timeout=900):
    """one bounded sample of the reference's OWN self-play code (oracle/_ref byte code, see oracle/ref_arm.py) in a fresh
    process: `cpus` worker processes through its mp.Pool fan-out, one game each, adjudicated after `plies` plies.
    -> dict(sims, seconds, sims_per_sec, cores, ...) or None when the compiled reference is not available"""
    from oracle import build_ref
    if not build_ref.available() and not os.path.isfile(os.path.join(build_ref.REFERENCE_DIR, "Checkers.py")):
        return None
    r = subprocess.run([sys.executable, "-m", "oracle.ref_arm", "--cpus", str(cpus), "--plies", str(plies), "--budget", str(budget)],
                       cwd=ROOT, capture_output=True, text=True, timeout=timeout)
    for line in reversed(r.stdout.splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    sys.stderr.write("reference sample failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return None


def run_reference(args, rank, world):
    """CPU arm.  Preferred: the reference's own implementation (kind "reference"): its generate_data() with its own
    mp.Pool.map fan-out over every host core (training_pipeline.py:323-332), game loop, MCTS and rules, byte-compiled
    from /root/reference into oracle/_ref at build time; only Keras is replaced by a torch-CPU stand-in.  Each step
    is a bounded sample: one game per core cut after REF_PLIES plies (the reference's TERMINATE_CNT) at 400 sims/move.
    Fallback when oracle/_ref was not built: the oracle port (kind "port")."""
    if rank != 0:
        return
    cores = usable_cores()
    vals, kind, sample = [], "reference", ""
    t0 = time.time()
    for i in range(args.warmup + args.steps):
        if i == args.warmup:
            t0 = time.time()
        r = reference_sample(cores, REF_PLIES)
        if r is None:
            kind = "port"
            break
        if i >= args.warmup:
            vals.append(r["sims_per_sec"])
        cores = r["cores"]
        sample = ("the reference's own generate_data(): %d worker processes (mp.Pool.map) x 1 self-play game cut after %d plies "
                  "(TERMINATE_CNT) at %d sims/move, torch-CPU stand-in for Keras (1 thread per worker, random-init weights); "
                  "%d sims in %.1f s per step" % (r["cores"], REF_PLIES, BUDGET, r["sims"], r["seconds"]))
    if kind == "port":
        import multiprocessing as mp
        from oracle import oracle as O
        O.build()
        per_step = 6.0
        with mp.get_context("spawn").Pool(cores) as pool:
            for _ in range(args.warmup):
                cpu_port_sample(cores, 1.0, pool=pool)
            t0 = time.time()
            for _ in range(args.steps):
                v, c, sims, el = cpu_port_sample(cores, per_step, pool=pool)
                vals.append(v)
        sample = ("oracle port (C tree + torch-CPU net, batch-1 eval), %d processes x %.0f s of self-play at %d sims/move per step"
                  % (cores, per_step, BUDGET))
    wall = time.time() - t0
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------
def pipeline_e2e(dev, rank, world, n_games, num_cpus):
    """wall clock of the reference-facing call, generate_Checkers_data(...).generate_data() (training_pipeline.py:312-332):
    `n_games` self-play games at cfg2 settings played to the end on this rank's GPU, records converted to the reference's
    [state, probs, q, z] lists and pickled into data/training_data/ like the reference's workers do.  Rank-local (every
    rank plays its own batch: the weak-scaling shape of the bench), files go to a scratch directory and are removed."""
    import shutil
    import tempfile
    import training_pipeline as TP
    from ckb200 import net as N
    work = tempfile.mkdtemp(prefix="ckb200_pipe_%d_" % rank)
    free = shutil.disk_usage(work).free / max(world, 1)     # every rank writes its own files at the same time
    need = n_games * 150 * 12000 * 1.5                      # ~150 records per game x 11.8 KB in the reference's format, with headroom
    if free < need:
        n_games = max(num_cpus, int(n_games * free / need / num_cpus) * num_cpus)
    os.makedirs(os.path.join(work, "data", "training_data"))
    os.makedirs(os.path.join(work, "data", "model"))
    cwd = os.getcwd()
    os.chdir(work)
    try:
        fn = TP.save_blob(N.random_init_blob(0), "data/model/Checkers_Model0_bench.npy")
        sp = dict(NUM_SELFPLAY_GAMES=n_games // num_cpus, TRAINING_ITERATION=0, TERMINATE_CNT=MCTS["terminate_cnt"], NUM_CPUS=num_cpus,
                  NN_FN=fn, DEVICE=dev, SEED=20261017 + rank, MAX_CONCURRENT_GAMES=SLOTS, SINGLE_PROCESS=True)
        mk = dict(UCT_C=4, CONSTRAINT='rollout', BUDGET=BUDGET, MULTIPROC=False, NEURAL_NET=True, VERBOSE=False, TRAINING=True,
                  DIRICHLET_ALPHA=1.0, DIRICHLET_EPSILON=0.25, TEMPERATURE_TAU=1.0, TEMPERATURE_DECAY=0.1, TEMP_DECAY_DELAY=10)
        gen = TP.generate_Checkers_data(sp, mk)
        t0 = time.time()
        with open(os.devnull, "w") as devnull:
            saved = sys.stdout
            sys.stdout = devnull                             # the reference prints one line per finished game
            try:
                files = gen.generate_data()
            finally:
                sys.stdout = saved
        wall = time.time() - t0
        files = [files] if isinstance(files, str) else files
        nbytes = sum(os.path.getsize(f) for f in files)
        st = gen.stats
        return {"value": st["sims"] / wall, "unit": UNIT, "call": "generate_Checkers_data(selfplay_kwargs, mcts_kwargs).generate_data()",
                "games": n_games, "workers_NUM_CPUS": num_cpus, "wall_s": wall, "gpu_s": st["gpu_ms"] / 1e3,
                "host_s_after_gpu": gen.host_seconds, "host_share_of_gpu_time": gen.host_seconds / max(st["gpu_ms"] / 1e3, 1e-9),
                "sims": st["sims"], "games_per_sec": n_games / wall, "plies_per_game": st["moves"] / max(n_games, 1), "records": gen.n_records, "pickle_bytes": nbytes, "files": len(files)}
    finally:
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)


def verify_sharding(dev, rank, world):
    """world > 1, outside the timed region: V games per rank at a small budget, sharded game g -> rank g mod world as
    in the bench, pooled on rank 0 with the packed device-side gather; rank 0 then replays ALL of them in one
    single-process engine and compares the pooled records byte for byte (a game's random streams are keyed by its
    global index, so the shards of an N-rank run must equal a 1-rank run)."""
    from ckb200 import dist as D
    from ckb200 import lib as L
    per_rank, budget, term = 32, 64, 24
    total = per_rank * world
    kw = dict(budget=budget, device=dev, evaluator="hash_salted", keep_records=True, seed=77, uct_c=4.0, training=True, alpha=1.0,
              epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10, terminate_cnt=term)
    base, stride, n_local = D.shard(total, rank, world)
    eng = L.Engine(L.make_cfg(n_slots=per_rank, game_id_base=base, game_id_stride=stride, **kw))
    eng.selfplay(n_local)
    pooled, _ms = D.gather_engine_records(eng, rank, world, "cuda:%d" % dev)
    eng.close()
    if rank != 0:
        return None
    eng = L.Engine(L.make_cfg(n_slots=total, **kw))
    eng.selfplay(total)
    single = eng.records()
    eng.close()
    key = lambda r: r[np.lexsort((r["ply"], r["game"]))]
    a, b = key(pooled), key(single)
    return {"games": total, "records": int(len(b)), "identical": bool(len(a) == len(b) and a.tobytes() == b.tobytes())}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from ckb200 import dist as D
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

    base, stride, _n = D.shard(args.slots * world, rank, world)          # game g -> rank g mod world
    cfg = L.make_cfg(n_slots=args.slots, budget=BUDGET, device=dev, evaluator="net", keep_records=True,
                     seed=20261017, game_id_base=base, game_id_stride=stride,
                     eval_cache_entries=-1 if args.no_eval_cache else args.cache_entries, max_chain_per_step=args.max_chain,
                     stagger_budget=PREROLL_BUDGET if args.preroll > 0 else 0, stagger_plies=PREROLL_PLIES if args.preroll > 0 else 0, **MCTS)
    eng = L.Engine(cfg)
    eng.set_net(0, net)
    n_games = args.slots * 16                    # staged games: enough refills for the pre-roll and any bench length
    eng.begin(n_games)
    rec_buf = np.zeros(args.slots * 201, dtype=L.RECORD_DTYPE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- untimed: desynchronise the slots (pre-roll at a tiny budget), then warm up at the real one ----------
    pre = dict(rounds=0, games=0, moves=0)
    if args.preroll > 0:
        st = eng.run(args.preroll)
        pre = dict(rounds=int(st["steps"]), games=int(st["games_finished"]), moves=int(st["moves"]))
        while eng.records_new(rec_buf)[1] > 0:    # the pre-roll's records are not the bench's (drained in buffer-sized pieces)
            pass
    for _ in range(args.warmup):
        eng.run(args.rounds)
        while eng.records_new(rec_buf)[1] > 0:
            pass
    eng.set_profile(True)
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()

    # ---- timed region: device-resident steps and end-to-end steps alternate, so both see the same game phases ----
    keys = ("sims", "nn_evals", "cache_hits", "gpu_ms", "eval_ms", "tower_ms", "kernel_launches", "games_finished", "moves", "steps")
    agg = {k: 0 for k in keys}
    e = dict(sims=0, h2d=0, d2h=0, ms=0.0, launches=0, games=0, records=0)
    barrier()
    t0 = time.time()
    def device_step():
        st = eng.run(args.rounds)                                           # device-resident step
        for k in keys:
            agg[k] += st[k]
        torch.cuda.synchronize()

    def e2e_step():
        s0 = time.time()                                                    # end-to-end step through host buffers
        weights.copy_(pinned, non_blocking=False)                           # H2D: this step's input (the weights)
        net.set_weights_device(weights.data_ptr(), weights.numel())
        st = eng.run(args.rounds)
        nrec, ngames = eng.records_new(rec_buf)                             # D2H: this step's results
        torch.cuda.synchronize()
        e["ms"] += 1000.0 * (time.time() - s0)
        e["sims"] += st["sims"]; e["launches"] += st["kernel_launches"]; e["games"] += st["games_finished"]; e["records"] += nrec
        e["h2d"] += blob.nbytes
        e["d2h"] += nrec * L.RECORD_DTYPE.itemsize + n_games * L.GAME_DTYPE.itemsize

    for i in range(args.steps):                  # the order flips every step, so a slow drift of the game mix cancels between the two
        for fn in ((device_step, e2e_step) if i % 2 == 0 else (e2e_step, device_step)):
            fn()
    barrier()
    wall = time.time() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- reduce over ranks: time = max, work = sum ----------------------------------------------
    t = torch.tensor([agg["gpu_ms"], e["ms"]], dtype=torch.float64, device="cuda:%d" % dev)
    w = torch.tensor([agg["sims"], agg["nn_evals"], agg["games_finished"], agg["moves"], e["sims"], agg["kernel_launches"] + e["launches"],
                      agg["cache_hits"], e["games"]], dtype=torch.float64, device="cuda:%d" % dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    gpu_ms, e2e_ms = t.tolist()
    sims, evals, games, moves, e2e_sims, launches, hits, e2e_games = w.tolist()

    # ---- outside the timed region ------------------------------------------------------------------
    gather_ms, pooled_n, gather_bytes, shard_check, decode_ms = None, None, None, None, None
    if world > 1:
        # the iteration-end NCCL gather of the finished games' records, packed on the device (ckb200.dist)
        D.warm_up_p2p(rank, world, "cuda:%d" % dev)               # NCCL opens point-to-point channels at first use
        pooled, info = D.gather_engine_records(eng, rank, world, "cuda:%d" % dev, want_info=True)
        if rank == 0:
            gather_ms, pooled_n, gather_bytes, decode_ms = info["gather_ms"], int(len(pooled)), info["bytes"], info["decode_ms"]
        del pooled
    eng.close()
    if world > 1:
        shard_check = verify_sharding(dev, rank, world)
    pipe = None
    if not args.no_pipeline:
        pipe = pipeline_e2e(dev, rank, world, args.pipeline_games, args.pipeline_workers)
        if world > 1:
            pv = torch.tensor([pipe["wall_s"]], dtype=torch.float64, device="cuda:%d" % dev)
            pw = torch.tensor([pipe["sims"], pipe["games"], pipe["records"]], dtype=torch.float64, device="cuda:%d" % dev)
            dist.all_reduce(pv, op=dist.ReduceOp.MAX)
            dist.all_reduce(pw, op=dist.ReduceOp.SUM)
            pipe["wall_s_max_over_ranks"] = float(pv.item())
            pipe["value"] = float(pw[0].item()) / float(pv.item())
            pipe["games_all_ranks"], pipe["records_all_ranks"] = int(pw[1].item()), int(pw[2].item())
            pipe["games_per_sec"] = pipe["games_all_ranks"] / float(pv.item())

    if rank != 0:
        return
    value = sims / (gpu_ms / 1000.0)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": gpu_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (split-fp16 tensor-core passes, fp32 accumulate)" if args.net_impl == "tc" else "f32",
            "data": "synthetic", "config": workload_config(world)}
    # how this arm ran the workload (kept out of `config`, which is the workload both arms share)
    line["engine"] = {"net_impl": args.net_impl, "eval_cache": "off" if args.no_eval_cache else "on (per-slot, %s entries, chain cap %s)" % (args.cache_entries or "default", args.max_chain or "default")}
    line["engine"]["steady_state"] = ("warm start: the first game a slot plays (game g) plays its first hash(g) mod %d plies at %d sims/move and everything "
                                      "after that at %d, so slots reach full-budget play at scattered game stages; untimed pre-roll of %d rounds "
                                      "(%d moves, %d games ended), then %d warm-up steps" %
                                      (PREROLL_PLIES, PREROLL_BUDGET, BUDGET, pre["rounds"], pre["moves"], pre["games"], args.warmup))
    line["wall_s"] = wall
    line["nn_evals_per_sec"] = (evals - hits) / (gpu_ms / 1000.0)
    line["expansions_per_sec"] = evals / (gpu_ms / 1000.0)
    line["eval_cache_hit_rate"] = hits / max(evals, 1.0)
    line["moves_per_sec"] = moves / (gpu_ms / 1000.0)
    line["games_finished"] = games
    line["games_per_sec"] = games / (gpu_ms / 1000.0)
    line["games_per_sec_est"] = (moves / (gpu_ms / 1000.0)) / 150.0      # at the ~150 plies/game BASELINE.md assumes
    if pipe is not None and pipe.get("plies_per_game"):
        # stationary rate: moves/s over the mean length of complete games (the e2e_pipeline batch, every game played to its end)
        line["games_per_sec_at_mean_length"] = (moves / (gpu_ms / 1000.0)) / pipe["plies_per_game"]
    line["e2e"] = {"value": e2e_sims / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": e["h2d"] // max(args.steps, 1),
                   "d2h_bytes_per_step": e["d2h"] // max(args.steps, 1), "games_per_sec": e2e_games / (e2e_ms / 1000.0),
                   "games_finished": e2e_games, "how": "alternates with the device-timed steps; wall clock of H2D weights + 400 rounds + D2H records"}
    if pipe is not None:
        line["e2e_pipeline"] = pipe
    line["gpu_launches"] = int(launches)
    line["clocks"] = clocks
    if gather_ms is not None:
        line["records_gather_ms"] = gather_ms              # device-side packing + exact-size NCCL send/recv into rank 0's HBM
        line["records_decode_ms"] = decode_ms              # rank 0: D2H + numpy decode into ck_record structs
        line["records_pooled"] = pooled_n
        line["records_gather_bytes"] = gather_bytes
    if shard_check is not None:
        line["sharding_check"] = shard_check

    # roofline of the dominant kernel (the tcgen05 tower; rank 0's own launches)
    peaks = measured_peaks()
    n_launch = agg["steps"]
    net_evals = agg["nn_evals"] - agg["cache_hits"]
    if args.net_impl == "tc" and agg["tower_ms"] > 0:
        peak = peaks["bf16_tflops_sustained"] if peaks else 1400.0
        achieved = net_evals * TOWER_FLOP_PER_POS / (agg["tower_ms"] / 1000.0) / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "tower_ncu_summary.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass
        line["roofline"] = {"bound": "tensor", "kernel": "tower_ts_kernel" if (os.environ.get("CK_TOWER") or "ts")[0] != "s" else "tower_tc_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak, "traffic": traffic,
                            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1400 (of fallback)",
                            "flop_per_position": TOWER_FLOP_PER_POS, "positions_per_launch": net_evals / max(n_launch, 1),
                            "avg_launch_ms": agg["tower_ms"] / max(n_launch, 1),
                            "share_of_step": agg["tower_ms"] / max(agg["gpu_ms"], 1e-9),
                            "issued_mma_tflops": 3 * achieved, "issued_mma_frac_of_peak": 3 * achieved / peak,
                            "note": "achieved/frac count useful FLOPs of EVALUATED positions only (SURVEY 8d); expansions served by the "
                                    "evaluation cache cost no FLOPs and are not counted.  The kernel issues 3 fp16 MMA passes per product "
                                    "(split hi/lo operands) for fp32-grade accuracy, so against the pipe's own peak frac is bounded by 1/3; "
                                    "issued_mma_* is the tensor-pipe work actually executed.  The denominator is the power-capped cuBLAS "
                                    "bf16 rate, which this kernel can exceed (it holds a higher clock)"}
    else:
        peak = 75.0
        achieved = net_evals * NET_FLOP_PER_POS / (max(agg["eval_ms"], 1e-9) / 1000.0) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "conv3x3_simt_kernel (CUDA-core cross-check path)", "achieved": achieved,
                            "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                            "peak_source": "nominal fp32 CUDA-core rate; this path is not the product kernel"}

    # CPU baseline on this box's host cores, bounded samples: the reference's own code on all cores and on one
    if world == 1 and not args.no_cpu_baseline:
        cores = usable_cores()
        r_all = reference_sample(cores, REF_PLIES)
        if r_all is not None:
            r_one = reference_sample(1, 1)
            line["cpu_baseline"] = {"value": r_all["sims_per_sec"], "unit": UNIT, "cores": r_all["cores"], "kind": "reference",
                                    "sample": "the reference's own generate_data() (oracle/_ref byte code, torch-CPU stand-in for Keras, 1 thread "
                                              "per worker): %d worker processes x 1 game cut after %d plies at %d sims/move = %d sims in %.1f s"
                                              % (r_all["cores"], REF_PLIES, BUDGET, r_all["sims"], r_all["seconds"]),
                                    "one_core": None if r_one is None else {"value": r_one["sims_per_sec"], "cores": 1, "sims": r_one["sims"],
                                                                            "seconds": r_one["seconds"]}}
        else:
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
    ap.add_argument("--no-pipeline", action="store_true", help="skip the generate_data() wall-clock measurement")
    ap.add_argument("--no-eval-cache", action="store_true", help="evaluate every leaf (A/B of the evaluation cache)")
    ap.add_argument("--cache-entries", type=int, default=0, help="evaluation cache entries per slot (0: the engine's default)")
    ap.add_argument("--max-chain", type=int, default=0, help="simulations a slot may chain inside one round (0: the engine's default)")
    ap.add_argument("--preroll", type=int, default=PREROLL_ROUNDS)
    ap.add_argument("--pipeline-games", type=int, default=SLOTS)
    ap.add_argument("--pipeline-workers", type=int, default=32, help="NUM_CPUS of the generate_data() call (files written)")
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
