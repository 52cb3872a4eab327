"""Arena between the reference's own trained networks (blobs committed under tests/golden/, imported from
data/model/Checkers_Model{10,5}_*.h5) at the reference's tournament settings (train_Checkers.py:180-202: BUDGET 200,
eps 0.25, tau 0): Model10 vs Model5, and each of them against an untrained (random-init) network, 100 games per pairing.
The reference's records for comparison (tests/golden/tournament_results.json): every trained iteration beats the
untrained Model0 (Model1 10/0/0; final round-robin row 0: -19 of -20); Model10 vs Model5 in the final round-robin: +1
for Model10 over two games.  One JSON line per pairing."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402

games = int(sys.argv[1]) if len(sys.argv) > 1 else 100
blobs = {"Model10": np.load(os.path.join(ROOT, "tests", "golden", "net_model10.npz"))["blob"],
         "Model5": np.load(os.path.join(ROOT, "tests", "golden", "net_model5.npz"))["blob"],
         "untrained": N.random_init_blob(0)}
for a, b in (("Model10", "Model5"), ("Model10", "untrained"), ("Model5", "untrained")):
    na, nb = L.Net(0), L.Net(0)
    na.set_weights(blobs[a])
    nb.set_weights(blobs[b])
    eng = L.Engine(L.make_cfg(n_slots=games, budget=200, training=False, alpha=1.0, epsilon=0.25, tau=0.0, arena=True,
                              keep_records=False, seed=2026, max_plies=1024))
    eng.set_net(0, na)
    eng.set_net(1, nb)
    st = eng.arena(games)
    g = eng.games()
    a_p1 = np.asarray(g["p1_net"]) == 0
    o = np.asarray(g["outcome"])
    wins = int(((o == 1) & a_p1).sum() + ((o == 2) & ~a_p1).sum())
    losses = int(((o == 2) & a_p1).sum() + ((o == 1) & ~a_p1).sum())
    print(json.dumps({"new": a, "old": b, "games": games, "wins_losses_draws": [wins, losses, games - wins - losses],
                      "plies_mean": float(np.mean(g["move_count"])), "plies_max": int(np.max(g["move_count"])),
                      "sims_per_sec": st["sims"] / (st["gpu_ms"] / 1e3), "seconds": st["gpu_ms"] / 1e3}), flush=True)
    eng.close()
    na.close()
    nb.close()
