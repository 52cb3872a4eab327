"""Secondary BASELINE.json configurations on one GPU (parity-test cases, not bench lines):

  cfg1  ONE self-play game at 50 sims/move (the reference's own CPU-runnable case) played to the end through the
        engine: latency of a lock-step round with a batch of one position
  cfg3  4096 concurrent self-play games per GPU at 800 sims/move (the per-GPU share of the
        32768-game / 8-GPU configuration), a bounded number of lock-step rounds
  cfg5  arena: evaluator games between two independent random-init networks (seeds 0 and 1), net A
        is player 1 in the first half, eps = 0.25, tau = 0, 400 sims/move, no ply cap, played to the
        end.  BASELINE asks for 1024 games over 8 GPUs (128 per GPU); --arena-games sets the count.

Prints one JSON line per configuration.  Usage: python scripts/bench_configs.py [--arena-games 128]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arena-games", type=int, default=128)
    ap.add_argument("--rounds", type=int, default=1600)
    args = ap.parse_args()
    L.require_device()
    nets = []
    for seed in (0, 1):
        net = L.Net(0)
        net.set_weights(N.random_init_blob(seed))
        nets.append(net)

    # ---- cfg1 ----
    eng = L.Engine(L.make_cfg(n_slots=1, budget=50, training=True, terminate_cnt=200, evaluator="net", keep_records=True,
                              uct_c=4.0, alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=1))
    eng.set_net(0, nets[0])
    eng.selfplay(1)                                   # warm-up game
    st = eng.selfplay(1)
    g = eng.games()
    print(json.dumps({"workload": "cfg1: 1 self-play game, 50 sims/move, random-init net, TERMINATE_CNT 200",
                      "sims_per_sec": st["sims"] / (st["gpu_ms"] / 1e3), "games_per_sec": 1e3 / st["gpu_ms"],
                      "ms_per_round": st["gpu_ms"] / max(st["steps"], 1), "sims": st["sims"], "gpu_ms": st["gpu_ms"],
                      "plies": int(g["move_count"][0])}), flush=True)
    eng.close()

    # ---- cfg3 ----
    eng = L.Engine(L.make_cfg(n_slots=4096, budget=800, training=True, terminate_cnt=200, evaluator="net", keep_records=True,
                              uct_c=4.0, alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=3))
    eng.set_net(0, nets[0])
    eng.begin(4096 * 4)
    eng.run(200)
    st = eng.run(args.rounds)
    print(json.dumps({"workload": "cfg3 (per-GPU share): 4096 concurrent self-play games, 800 sims/move", "rounds": args.rounds,
                      "sims_per_sec": st["sims"] / (st["gpu_ms"] / 1e3), "moves_per_sec": st["moves"] / (st["gpu_ms"] / 1e3),
                      "nn_evals": st["nn_evals"], "sims": st["sims"], "gpu_ms": st["gpu_ms"], "compactions": st["compactions"],
                      "nodes_created": st["nodes_created"]}), flush=True)
    eng.close()

    # ---- cfg5 ----
    g = args.arena_games
    eng = L.Engine(L.make_cfg(n_slots=g, budget=400, training=False, terminate_cnt=0, evaluator="net", arena=True, keep_records=False,
                              uct_c=4.0, alpha=1.0, epsilon=0.25, tau=0.0, seed=5, max_plies=1024))
    eng.set_net(0, nets[0])
    eng.set_net(1, nets[1])
    st = eng.arena(g)
    games = eng.games()
    out = np.asarray(games["outcome"])
    plies = np.asarray(games["move_count"])
    line = {"workload": "cfg5 arena: %d games, net A (seed 0) vs net B (seed 1), 400 sims/move, eps 0.25, tau 0" % g,
            "games_finished": int(st["games_finished"]), "games_per_sec": st["games_finished"] / (st["gpu_ms"] / 1e3),
            "sims_per_sec": st["sims"] / (st["gpu_ms"] / 1e3), "sims": st["sims"], "gpu_ms": st["gpu_ms"],
            "outcomes": {str(int(k)): int((out == k).sum()) for k in np.unique(out)}}
    if plies is not None:
        line["plies_mean"] = float(plies.mean()); line["plies_max"] = int(plies.max())
    print(json.dumps(line), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
