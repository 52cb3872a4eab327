"""Runs the unmodified reference next to the oracle (build container only; skipped where
/root/reference is not mounted -- the committed golden vectors cover that case)."""
import numpy as np
import pytest

from helpers import KerasLikeStub, codec
from oracle import oracle as O
from oracle import ref_harness as H

pytestmark = pytest.mark.needs_reference


def test_random_games_movegen_and_draw_plane():
    ref = H.load_reference()
    env = ref.Checkers.Checkers()
    rng = np.random.RandomState(99)
    npos = 0
    for _ in range(12):
        env.reset()
        rev, ply = 0, 0
        while True:
            st = env.state
            pos = codec.encode_state(st, rev, ply)
            kids, mask, status, p5 = O.movegen(pos)
            raw = env._check_moves(env.history)
            done, outcome = env.determine_outcome(env.history, legal_moves=raw)
            npos += 1
            assert len(raw) == len(kids)
            for r, c in zip(raw, kids):
                e = codec.encode_state(r)
                assert e[:3] == c[:3]
                assert codec.meta_player(e[3]) == codec.meta_player(c[3])
                assert codec.meta_action(e[3]) == codec.meta_action(c[3])
            assert [codec.plane_to_bits(st[6 + i]) for i in range(8)] == mask
            assert codec.OUTCOME_NAMES[status] == outcome
            assert st[5, 0, 0] == p5 / 80
            assert (codec.decode_state(pos, mask, p5) == st).all()
            if done:
                break
            i = rng.randint(len(raw))
            env.step(raw[i])
            rev, ply = codec.meta_rev(kids[i][3]), codec.meta_ply(kids[i][3])
            assert ply == len(env.history) - 1
    assert npos > 500


def test_tree_search_matches_reference_from_midgame():
    """search from a position reached by random play, hash evaluator (non-trivial floats)."""
    with H.reference_modules() as ref:
        env = ref.Checkers.Checkers(KerasLikeStub("hash"))
        rng = np.random.RandomState(5)
        rev = ply = 0
        pos = codec.encode_state(env.state)
        for _ in range(24):
            kids, _, _, _ = O.movegen(pos)
            i = rng.randint(len(env.legal_next_states))
            env.step(env.legal_next_states[i])
            pos = kids[i]
        assert not env.done
        ref.MCTS.MCTS(GAME_ENV=env, UCT_C=4, CONSTRAINT='rollout', BUDGET=300, MULTIPROC=False, NEURAL_NET=True,
                      VERBOSE=False, TRAINING=False, DIRICHLET_ALPHA=1.0, DIRICHLET_EPSILON=0.0,
                      TEMPERATURE_TAU=0, TEMPERATURE_DECAY=0, TEMP_DECAY_DELAY=0)
        root = ref.MCTS.MCTS_Node(env.state)
        root.history = list(env.history)        # same history length as the game (ply index)
        ref.MCTS.MCTS.begin_tree_search(root)
        parent_player = int(env.history[-2][4, 0, 0])
        t = O.Tree(pos, O.make_cfg(budget=300), "hash", parent_player=parent_player)
        t.search(300)
        n, w = t.root_stats()
        assert (n, float(w)) == (root.n, float(root.w))
        got = t.root_children()
        assert [(c["n"], float(c["w"]), float(c["p"])) for c in got] == \
               [(c.n, float(c.w), float(c.p)) for c in root.children]
