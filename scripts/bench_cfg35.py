"""BASELINE.json configs[2] and configs[4] AT THEIR STATED SIZE, one process per GPU under torchrun:

  cfg3  32768 concurrent self-play games, 800 sims/move, game g on rank g mod world (4096 per GPU at world = 8), every
        game played TO THE END (TERMINATE_CNT 200); records pooled on rank 0 with the packed device-side gather
  cfg5  arena: 1024 evaluator games new-net (seed 0) vs old-net (seed 1), 400 sims/move, eps 0.25, tau 0, no ply cap,
        net A is player 1 in games < 512 (training_pipeline.py:523-528), games sharded over the ranks; W/L/D pooled

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_cfg35.py [--games3 32768] [--games5 1024]

Time = max over ranks of the CUDA-event time of the run; work = sum over ranks.  One JSON line per configuration on rank 0."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from ckb200 import dist as D  # noqa: E402
from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402


def reduce(dev, times, sums):
    t = torch.tensor(times, dtype=torch.float64, device=dev)
    s = torch.tensor(sums, dtype=torch.float64, device=dev)
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return t.tolist(), s.tolist()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games3", type=int, default=32768)
    ap.add_argument("--sims3", type=int, default=800)
    ap.add_argument("--games5", type=int, default=1024)
    ap.add_argument("--skip", default="")
    ap.add_argument("--chain", type=int, default=0, help="max_chain_per_step of both engines (0: the engine's default)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    nets = []
    for seed in (0, 1):
        net = L.Net(local)
        net.set_weights(N.random_init_blob(seed))
        nets.append(net)
    out = sys.__stdout__

    if "3" not in args.skip:
        base, stride, n_local = D.shard(args.games3, rank, world)
        eng = L.Engine(L.make_cfg(n_slots=n_local, budget=args.sims3, device=local, training=True, terminate_cnt=200, evaluator="net",
                                  keep_records=True, uct_c=4.0, alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10,
                                  seed=3, game_id_base=base, game_id_stride=stride, max_chain_per_step=args.chain))
        eng.set_net(0, nets[0])
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        st = eng.selfplay(n_local)
        torch.cuda.synchronize()
        wall = time.time() - t0
        games = eng.games()
        D.warm_up_p2p(rank, world, dev)
        pooled, info = D.gather_engine_records(eng, rank, world, dev, want_info=True)
        (gpu_ms, wall_max), (sims, evals, hits, moves, ngames, nrec) = reduce(
            dev, [st["gpu_ms"], wall], [st["sims"], st["nn_evals"], st["cache_hits"], st["moves"], len(games), eng.records_packed_sizes()[0]])
        oc = np.bincount(np.asarray(games["outcome"]), minlength=4)[1:4].astype(np.float64)
        (_,), oc = reduce(dev, [0.0], oc.tolist())
        if rank == 0:
            out.write(json.dumps({"workload": "cfg3: %d concurrent self-play games, %d sims/move, sharded by game index over %d B200, played to the end"
                                              % (args.games3, args.sims3, world), "n_gpus": world, "games": int(ngames),
                                  "sims_per_sec": sims / (gpu_ms / 1e3), "games_per_sec": ngames / (gpu_ms / 1e3), "gpu_s_max_over_ranks": gpu_ms / 1e3,
                                  "wall_s_max_over_ranks": wall_max, "sims": sims, "moves": moves, "plies_per_game": moves / max(ngames, 1),
                                  "eval_cache_hit_rate": hits / max(evals, 1), "network_evals": evals - hits, "records": int(nrec),
                                  "records_pooled_on_rank0": int(len(pooled)), "records_gather_ms": info["gather_ms"], "records_decode_ms": info["decode_ms"], "records_gather_bytes": info["bytes"],
                                  "outcomes_p1_p2_draw": [int(v) for v in oc]}) + "\n")
            out.flush()
        del pooled
        eng.close()

    if "5" not in args.skip:
        base, stride, n_local = D.shard(args.games5, rank, world)
        eng = L.Engine(L.make_cfg(n_slots=n_local, budget=400, device=local, training=False, terminate_cnt=0, evaluator="net", arena=True,
                                  keep_records=False, uct_c=4.0, alpha=1.0, epsilon=0.25, tau=0.0, seed=5, max_plies=1024,
                                  game_id_base=base, game_id_stride=stride, max_chain_per_step=args.chain))
        eng.set_net(0, nets[0])
        eng.set_net(1, nets[1])
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.time()
        st = eng.arena(n_local)
        torch.cuda.synchronize()
        wall = time.time() - t0
        games = eng.games()
        a_is_p1 = np.asarray(games["p1_net"]) == 0
        o = np.asarray(games["outcome"])
        a_wins = int(((o == 1) & a_is_p1).sum() + ((o == 2) & ~a_is_p1).sum())
        b_wins = int(((o == 2) & a_is_p1).sum() + ((o == 1) & ~a_is_p1).sum())
        plies = np.asarray(games["move_count"]).astype(np.float64)
        (gpu_ms, wall_max, pmax), (sims, evals, hits, ngames, aw, bw, dr, psum) = reduce(
            dev, [st["gpu_ms"], wall, float(plies.max())],
            [st["sims"], st["nn_evals"], st["cache_hits"], len(games), a_wins, b_wins, int((o == 3).sum()), float(plies.sum())])
        if rank == 0:
            out.write(json.dumps({"workload": "cfg5 arena: %d games net A (seed 0) vs net B (seed 1), 400 sims/move, eps 0.25, tau 0, sharded over %d B200"
                                              % (args.games5, world), "n_gpus": world, "games": int(ngames), "games_per_sec": ngames / (gpu_ms / 1e3),
                                  "sims_per_sec": sims / (gpu_ms / 1e3), "gpu_s_max_over_ranks": gpu_ms / 1e3, "wall_s_max_over_ranks": wall_max,
                                  "A_wins": int(aw), "B_wins": int(bw), "draws": int(dr), "plies_mean": psum / max(ngames, 1), "plies_max": int(pmax),
                                  "eval_cache_hit_rate": hits / max(evals, 1)}) + "\n")
            out.flush()
        eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
