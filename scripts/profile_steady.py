"""Profiling target: the cfg2 loop brought to its steady state the way bench.py does it (warm-start stagger + pre-roll),
then `rounds` more lock-step rounds.  Usage (under ncu, -s skips the pre-roll's launches of the kernel of interest):
    ncu --set full -k regex:tree_step_kernel -s 1400 -c 1 ... python scripts/profile_steady.py [total_rounds]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
net = L.Net(0)
net.set_weights(N.random_init_blob(0))
eng = L.Engine(L.make_cfg(n_slots=4096, budget=400, training=True, terminate_cnt=200, evaluator="net", keep_records=True,
                          uct_c=4.0, alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=20261017,
                          stagger_budget=8, stagger_plies=140))
eng.set_net(0, net)
eng.begin(4096 * 16)
st = eng.run(rounds)
print("ok", st["sims"], st["nn_evals"], st["cache_hits"], st["gpu_ms"])
