"""Throughput of the playout mode (NEURAL_NET=False: UCT + one random playout per simulation, the reference's
iteration-0 self-play, train_Checkers.py:78,96): N concurrent games, BUDGET simulations per move, lock-step
rounds of tree_step_kernel<uct> + playout_eval_kernel.  Prints one JSON line per slot count.
Usage: python scripts/bench_uct.py [budget] [steps] [slots ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
from ckb200 import lib as L  # noqa: E402

budget = int(sys.argv[1]) if len(sys.argv) > 1 else 200
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
slot_list = [int(v) for v in sys.argv[3:]] or [4096, 16384, 65536]
L.require_device()
for slots in slot_list:
    eng = L.Engine(L.make_cfg(n_slots=slots, budget=budget, training=True, terminate_cnt=200, evaluator="rollout",
                              keep_records=True, uct_c=4.0, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=20261017,
                              pool_cap=8192))          # a 200-simulation UCT search adds ~1.1 k nodes
    eng.begin(slots * 2)
    eng.set_profile(True)
    eng.run(budget)                                   # warm-up: one move of every game
    sims, ms, ev_ms, games, moves = 0, 0.0, 0.0, 0, 0
    for _ in range(steps):
        st = eng.run(budget)
        sims += st["sims"]; ms += st["gpu_ms"]; ev_ms += st["eval_ms"]; games += st["games_finished"]; moves += st["moves"]
    rounds = steps * budget
    print(json.dumps({"mode": "uct_playouts", "slots": slots, "budget": budget, "rounds": rounds,
                      "sims_per_sec": sims / (ms / 1e3), "ms_per_round": ms / rounds,
                      "playout_ms_per_round": ev_ms / rounds, "tree_ms_per_round": (ms - ev_ms) / rounds,
                      "moves": moves, "games_finished": games}), flush=True)
    eng.close()
