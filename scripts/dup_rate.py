#!/usr/bin/env python3
"""How many leaf evaluations of a self-play game repeat an earlier evaluation of the SAME network input
(position + side to move + plane 5 + legal-action planes)?  The two per-colour trees of a game
(training_pipeline.py:353,372) search nearly the same positions, and games share their openings.
CPU measurement with the oracle port (test infrastructure) and the random-init network, cfg2 constants.

    python scripts/dup_rate.py [games] [budget] [max_plies]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

from ckb200 import codec
from ckb200 import net as N
from oracle import net_oracle as NO
from oracle import oracle as O

games = int(sys.argv[1]) if len(sys.argv) > 1 else 2
budget = int(sys.argv[2]) if len(sys.argv) > 2 else 400
max_plies = int(sys.argv[3]) if len(sys.argv) > 3 else 200
torch.set_num_threads(8)
model = NO.TorchKerasLike(N.random_init_blob(0))
seen_global = {}
tot = dup_game = dup_global = 0
t0 = time.time()
for g in range(games):
    seen = {}
    window = []          # (eval index, key) for hit-distance statistics
    dist = []

    def ev(pos, mask, plane5, seen=seen, dist=dist):
        global tot, dup_game, dup_global
        key = (pos[0], pos[1], pos[2], pos[3] & 1, plane5)
        tot += 1
        if key in seen:
            dup_game += 1
            dist.append(tot - seen[key])
        elif key in seen_global:
            dup_global += 1
        seen[key] = tot
        seen_global[key] = 1
        x = codec.nn_input_planes(pos, mask, plane5).reshape(1, 8, 8, 14)
        p, v = model.predict(x)
        return p[0], v[0, 0]

    cfg = O.make_cfg(uct_c=4.0, budget=budget, training=True, alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1,
                     tau_decay_delay=10, terminate_cnt=max_plies, seed=100 + g)
    gm = O.Game(cfg, O.python_eval(ev))
    gm.play()
    d = np.array(dist) if dist else np.zeros(1)
    print("game %d: plies %d evals so far %d  same-game dups %.3f  cross-game dups %.3f  hit distance (evals) p50 %d p90 %d p99 %d  [%.0f s]"
          % (g, gm.move_count, tot, dup_game / tot, dup_global / tot, np.percentile(d, 50), np.percentile(d, 90), np.percentile(d, 99),
             time.time() - t0), flush=True)
    gm.close()
