"""How full is the tower's last wave?  Brings the cfg2 loop to its steady state the way bench.py does (warm-start stagger +
pre-roll), then runs single rounds and records the size of every evaluator batch (network evaluations expanded by the
next round) with the tower time of its launch.  Prints one JSON line: histogram of the batch size in waves of
4 x SMs positions, the mean tower time per iteration count, and the mean fill of the last wave.
    [CK_BATCH_WAVES=0] python scripts/batch_hist.py [preroll_rounds] [rounds]"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402

preroll = int(sys.argv[1]) if len(sys.argv) > 1 else 24000
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
WAVE = 592
net = L.Net(0)
net.set_weights(N.random_init_blob(0))
eng = L.Engine(L.make_cfg(n_slots=4096, budget=400, training=True, terminate_cnt=200, evaluator="net", keep_records=False,
                          uct_c=4.0, alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=20261017,
                          stagger_budget=8, stagger_plies=140))
eng.set_net(0, net)
eng.begin(4096 * 16)
eng.run(preroll)
miss, tower, sims = [], [], []
for _ in range(rounds + 1):
    st = eng.run(1)
    miss.append(st["nn_evals"] - st["cache_hits"])
    tower.append(st["tower_ms"])
    sims.append(st["sims"])
n = np.array(miss[1:], dtype=np.int64)            # the batch round r staged is expanded (and counted) by round r + 1
t = np.array(tower[:-1])
its = np.ceil(n / WAVE).astype(int)
out = {"batch_waves_env": os.environ.get("CK_BATCH_WAVES", ""), "preroll": preroll, "rounds": rounds,
       "batch_mean": float(n.mean()), "batch_p5_p50_p95": [int(x) for x in np.percentile(n, [5, 50, 95])],
       "sims_per_round": float(np.mean(sims)), "tower_ms_mean": float(t.mean()),
       "last_wave_fill_mean": float(np.mean(n / WAVE - (its - 1))),
       "by_iterations": {str(k): {"launches": int((its == k).sum()), "tower_ms": float(t[its == k].mean()),
                                  "batch_mean": float(n[its == k].mean())} for k in sorted(set(its.tolist()))},
       "exact_whole_waves": int((n % WAVE == 0).sum()),
       "excess_hist_tenths_of_wave": np.bincount(np.minimum(((n % WAVE) * 10 // WAVE), 9), minlength=10).tolist()}
print(json.dumps(out))
