"""Multi-GPU parity of the pipeline entry points: generate_Checkers_data / tournament_Checkers under torchrun
with two ranks (games sharded by index, records pooled on rank 0 over NCCL) must write exactly the files a
single process writes for the same seed -- a game's random streams depend on its global index only."""
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu

MCTS = dict(UCT_C=4, CONSTRAINT='rollout', BUDGET=24, MULTIPROC=False, NEURAL_NET=True, VERBOSE=False, TRAINING=True,
            DIRICHLET_ALPHA=1.0, DIRICHLET_EPSILON=0.25, TEMPERATURE_TAU=1.0, TEMPERATURE_DECAY=0.1, TEMP_DECAY_DELAY=10)
SPEC = dict(selfplay=dict(NUM_SELFPLAY_GAMES=5, TRAINING_ITERATION=3, TERMINATE_CNT=40, NUM_CPUS=2, NN_FN='stub:hash_salted', SEED=77),
            mcts=MCTS,
            tourney=dict(NEW_NN_FN='stub:hash', OLD_NN_FN='stub:uniform_material', TOURNEY_GAMES=4, NUM_CPUS=1, SEED=5),
            tourney_mcts=dict(MCTS, TRAINING=False, TEMPERATURE_TAU=0, TEMPERATURE_DECAY=0, TEMP_DECAY_DELAY=0))


def _run(cwd, launcher):
    os.makedirs(cwd)
    worker = os.path.join(ROOT, "tests", "dist_pipeline_worker.py")
    subprocess.run(launcher + [worker, json.dumps(SPEC)], cwd=cwd, check=True, timeout=600,
                   stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    res = json.load(open(os.path.join(cwd, "result.json")))
    data = [pickle.load(open(os.path.join(cwd, fn), "rb")) for fn in res["data_fns"]]
    table = open(os.path.join(cwd, res["tourney_fn"]), encoding="utf-8").read()
    return data, table


def test_two_ranks_write_the_single_process_files(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    one, table1 = _run(str(tmp_path / "one"), [sys.executable])
    two, table2 = _run(str(tmp_path / "two"), [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                                               "--master-addr", "127.0.0.1", "--master-port", str(29700 + os.getpid() % 200)])
    assert len(one) == len(two) == 2
    for a, b in zip(one, two):
        assert len(a) == len(b) > 0
        for ra, rb in zip(a, b):
            assert (ra[0] == rb[0]).all() and (ra[1] == rb[1]).all() and ra[2] == rb[2] and ra[3] == rb[3]
    assert table1 == table2 and 'Wins/Losses/Draws' in table1
