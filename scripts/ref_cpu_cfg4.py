"""cfg4 on the CPU with the REFERENCE's own code (SURVEY 8d "reference CPU path timed beside it"): one core,
  (i)  movegen only: Checkers.get_legal_next_states (Checkers.py:77-92 -> _check_moves :94-304) over positions reached by
       k ~ U[0, 60] uniformly random legal plies from the start position;
  (ii) random playouts: MCTS.default_policy with NEURAL_NET=False (MCTS.py:132-143) from the start position.
Uses the byte-compiled copy in oracle/_ref when it exists (the GPU box), else the source tree (this container).
CPU only, test/measurement infrastructure -- nothing here is on the product path.
    python scripts/ref_cpu_cfg4.py [seconds_per_leg]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from oracle import ref_harness as H  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
compiled = H.compiled_reference_available()
with H.reference_modules(compiled=compiled) as ref:
    rng = np.random.RandomState(4)
    # positions: random walks with the reference's own environment
    histories = []
    while len(histories) < 256:
        env = ref.Checkers.Checkers(None)
        k = rng.randint(0, 61)
        for _ in range(k):
            if env.done:
                break
            env.step(env.legal_next_states[rng.randint(len(env.legal_next_states))])
        if not env.done:
            histories.append([s.copy() for s in env.history])
    env = ref.Checkers.Checkers(None)
    t0 = time.time()
    n = succ = 0
    while time.time() - t0 < seconds:
        for h in histories:
            succ += len(env.get_legal_next_states(h))
        n += len(histories)
    mg = time.time() - t0
    # playouts
    env = ref.Checkers.Checkers(None)
    ref.MCTS.MCTS(GAME_ENV=env, UCT_C=4, CONSTRAINT='rollout', BUDGET=1, MULTIPROC=False, NEURAL_NET=False, VERBOSE=False,
                  TRAINING=False, DIRICHLET_ALPHA=1.0, DIRICHLET_EPSILON=0.25, TEMPERATURE_TAU=0, TEMPERATURE_DECAY=0,
                  TEMP_DECAY_DELAY=0)
    node = ref.MCTS.MCTS_Node(env.state)
    np.random.seed(1)
    t0 = time.time()
    playouts = 0
    outcomes = {}
    while time.time() - t0 < seconds:
        outcome, _player = ref.MCTS.MCTS.default_policy(node)
        outcomes[outcome] = outcomes.get(outcome, 0) + 1
        playouts += 1
    po = time.time() - t0
print(json.dumps({"workload": "cfg4 on the CPU, the reference's own code, 1 core", "kind": "reference (compiled copy)" if compiled else "reference (source tree)",
                  "cpu": os.uname().machine, "cores_used": 1,
                  "movegen_positions_per_sec": n / mg, "movegen_successors_per_position": succ / max(n, 1), "movegen_positions": n,
                  "playouts_per_sec": playouts / po, "playouts": playouts, "playout_outcomes": outcomes}))
