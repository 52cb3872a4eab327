"""cfg2 loop (4096 games, 400 sims/move) with the evaluation cache: simulations/s, network evaluations per round and
the tree kernel's share for different cache sizes and per-round chain caps.  Every configuration starts from the
same seed with bench.py's warm start (stagger + pre-roll of `warm` steps), and is then timed over `steps` steps of 400
rounds.  CK_BATCH_WAVES=0 in the environment switches the wave-shaped batches off (A/B).
Usage: python scripts/cache_sweep.py [steps] [warm]   (one JSON line per configuration)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
from ckb200 import lib as L  # noqa: E402
from ckb200 import net as N  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 6
configs = [(-1, 0), (4096, 2), (4096, 4), (4096, 8), (4096, 16), (1024, 8), (16384, 8)]
if len(sys.argv) > 3:
    configs = [tuple(int(v) for v in c.split(":")) for c in sys.argv[3:]]
slots = int(os.environ.get("SWEEP_SLOTS", "4096"))
net = L.Net(0)
net.set_weights(N.random_init_blob(0))
for cache, chain in configs:
    eng = L.Engine(L.make_cfg(n_slots=slots, budget=400, training=True, terminate_cnt=200, evaluator="net", keep_records=True,
                              uct_c=4.0, alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=20261017,
                              eval_cache_entries=cache, max_chain_per_step=chain, stagger_budget=8, stagger_plies=140))
    eng.set_net(0, net)
    eng.begin(slots * 8)
    eng.set_profile(True)
    for _ in range(warm):
        eng.run(400)
    agg = dict(sims=0, nn_evals=0, cache_hits=0, gpu_ms=0.0, eval_ms=0.0, tower_ms=0.0, moves=0, steps=0)
    for _ in range(steps):
        st = eng.run(400)
        for k in agg:
            agg[k] += st[k]
    r = agg["steps"]
    print(json.dumps({"slots": slots, "overlap": os.environ.get("CK_OVERLAP", "default"), "batch_waves": os.environ.get("CK_BATCH_WAVES", "auto"), "cache_entries": cache, "max_chain": chain, "sims_per_sec": agg["sims"] / (agg["gpu_ms"] / 1e3),
                      "ms_per_round": agg["gpu_ms"] / r, "tower_ms_per_round": agg["tower_ms"] / r,
                      "tree_ms_per_round": (agg["gpu_ms"] - agg["eval_ms"]) / r, "heads_ms_per_round": (agg["eval_ms"] - agg["tower_ms"]) / r,
                      "sims_per_round": agg["sims"] / r, "net_evals_per_round": (agg["nn_evals"] - agg["cache_hits"]) / r,
                      "hit_rate": agg["cache_hits"] / max(agg["nn_evals"], 1), "moves_per_sec": agg["moves"] / (agg["gpu_ms"] / 1e3)}), flush=True)
    eng.close()
