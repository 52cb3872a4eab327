"""BASELINE INFRASTRUCTURE ONLY -- times the reference's OWN self-play code on the host cores.

Runs ``generate_Checkers_data(selfplay_kwargs, mcts_kwargs).generate_data()`` of the byte-compiled reference
(oracle/_ref, see build_ref.py) verbatim: its own ``mp.Pool.map`` fan-out over ``NUM_CPUS`` worker processes
(training_pipeline.py:323-332), its own game loop, MCTS and Checkers rules.  Only TensorFlow/Keras are replaced:
``load_model`` returns the PyTorch-CPU restatement of ``create_nn`` (oracle/net_oracle.py, one thread per worker)
with random-init weights, as BASELINE.md section 3 prescribes.  The sample is bounded through the reference's own
``TERMINATE_CNT`` knob: every worker plays one game that is adjudicated after ``plies`` plies, i.e. ``plies``
searches of ``budget`` simulations each.

    python -m oracle.ref_arm --cpus 16 --plies 4 --budget 400      # prints one JSON line

Called by bench.py (``--impl reference`` and the ``cpu_baseline`` leg) as a subprocess, so that the fork-based pool
starts from a process that has no CUDA context and has not run any multi-threaded torch code.
"""
import argparse
import json
import os
import pickle
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _load_model(_path):
    import torch
    torch.set_num_threads(1)
    from ckb200 import net as N
    from oracle import net_oracle as NO
    return NO.TorchKerasLike(N.random_init_blob(0))


def selfplay_sample(cpus, plies, budget, games_per_worker=1):
    from oracle import ref_harness as H
    H.set_load_model(_load_model)
    work = tempfile.mkdtemp(prefix="ckrefarm_")
    os.makedirs(os.path.join(work, "data", "training_data"))
    cwd = os.getcwd()
    os.chdir(work)
    try:
        with H.reference_modules(with_pipeline=True, compiled=not H.reference_available()) as ref:
            sp = dict(NUM_SELFPLAY_GAMES=games_per_worker, TRAINING_ITERATION=0, TERMINATE_CNT=plies, NUM_CPUS=cpus, NN_FN="random-init")
            mk = dict(GAME_ENV=None, UCT_C=4, CONSTRAINT="rollout", BUDGET=budget, MULTIPROC=False, NEURAL_NET=True, VERBOSE=False,
                      TRAINING=True, DIRICHLET_ALPHA=1.0, DIRICHLET_EPSILON=0.25, TEMPERATURE_TAU=1.0, TEMPERATURE_DECAY=0.1,
                      TEMP_DECAY_DELAY=10)
            gen = ref.training_pipeline.generate_Checkers_data(sp, mk)
            used = gen.num_cpus                                  # the reference clamps to mp.cpu_count() (:320-321)
            t0 = time.time()
            fns = gen.generate_data()
            dt = time.time() - t0
        fns = [fns] if isinstance(fns, str) else list(fns)
        moves = 0
        for fn in fns:
            for rec in pickle.load(open(fn, "rb")):
                moves += 1 if float(rec[1].sum()) > 0 else 0     # records of searched moves (terminal records carry zero planes)
    finally:
        os.chdir(cwd)
        shutil.rmtree(work, ignore_errors=True)
    sims = moves * budget                                        # CONSTRAINT='rollout': exactly BUDGET simulations per move (MCTS.py:188-201)
    return dict(sims=sims, seconds=dt, sims_per_sec=sims / dt, games=len(fns) * games_per_worker, moves=moves, cores=used,
                plies_per_game=plies, budget=budget, kind="reference")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpus", type=int, default=1)
    ap.add_argument("--plies", type=int, default=4)
    ap.add_argument("--budget", type=int, default=400)
    a = ap.parse_args()
    out = selfplay_sample(a.cpus, a.plies, a.budget)
    sys.stdout.write(json.dumps(out) + "\n")
