"""Steady-state check of the cfg2 loop: per-step simulations/s, evaluator share, finished games and tree
compactions over many steps (one step = 400 lock-step rounds = about one move of every game).
Usage: python scripts/long_run.py [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
max_term = int(sys.argv[2]) if len(sys.argv) > 2 else 0
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 0
pool_cap = int(os.environ.get("LR_POOL_CAP", "0"))
eps = float(os.environ.get("LR_EPS", "0.25"))
net = L.Net(0)
net.set_weights(N.random_init_blob(0))
eng = L.Engine(L.make_cfg(n_slots=4096, budget=400, training=True, terminate_cnt=200, evaluator="net", keep_records=True,
                          uct_c=4.0, alpha=1.0, epsilon=eps, tau=1.0, pool_cap=pool_cap, tau_decay=0.1, tau_decay_delay=10, seed=20261017,
                          max_terminal_sims_per_step=max_term))
eng.set_net(0, net)
eng.begin(4096 * 8)
eng.set_profile(True)
tot_sims, tot_ms, games = 0, 0.0, 0
for s in range(steps):
    st = eng.run(400)
    if s >= warm:
        tot_sims += st["sims"]; tot_ms += st["gpu_ms"]; games += st["games_finished"]
    if s % 5 == 4 or s == steps - 1:
        print(json.dumps({"step": s + 1, "sims_per_sec": st["sims"] / (st["gpu_ms"] / 1e3), "ms_per_round": st["gpu_ms"] / 400,
                          "eval_ms_per_round": st["eval_ms"] / 400, "tower_ms_per_round": st["tower_ms"] / 400,
                          "nn_evals_per_round": st["nn_evals"] / 400, "games_finished_total": games,
                          "compactions": st["compactions"], "moves": st["moves"]}), flush=True)
print(json.dumps({"steps": steps, "warm": warm, "max_terminal_sims_per_step": max_term, "sims_per_sec_overall": tot_sims / (tot_ms / 1e3), "games_finished": games,
                  "games_per_sec": games / (tot_ms / 1e3)}))
