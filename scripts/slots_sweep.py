"""Throughput against the number of concurrent games per GPU (400 sims/move, self-play settings of cfg2):
simulations/s over `rounds` lock-step rounds after a short warm-up.  Usage: python scripts/slots_sweep.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402

net = L.Net(0)
net.set_weights(N.random_init_blob(0))
for slots in (128, 296, 592, 1024, 2048, 4096, 8192, 16384):
    eng = L.Engine(L.make_cfg(n_slots=slots, budget=400, training=True, terminate_cnt=200, evaluator="net", keep_records=False,
                              uct_c=4.0, alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=7))
    eng.set_net(0, net)
    eng.begin(slots * 4)
    eng.set_profile(True)
    eng.run(200)
    st = eng.run(800)
    print(json.dumps({"slots": slots, "sims_per_sec": st["sims"] / (st["gpu_ms"] / 1e3), "ms_per_round": st["gpu_ms"] / 800,
                      "tower_ms_per_round": st["tower_ms"] / 800, "eval_ms_per_round": st["eval_ms"] / 800}), flush=True)
    eng.close()
