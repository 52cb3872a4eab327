"""Runs the unmodified reference next to the oracle (build container only; skipped where
/root/reference is not mounted -- the committed golden vectors cover that case)."""
import os

import numpy as np
import pytest

from helpers import KerasLikeStub, codec, hashed_playouts
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


@pytest.mark.parametrize("plies,budget", [(24, 160), (70, 120)])
def test_uct_playout_search_matches_reference_from_midgame(plies, budget):
    """NEURAL_NET=False (UCT, one child per visit, a playout per simulation) from positions reached by random play;
    np.random.randint inside the reference's default_policy is replaced by the position hash the oracle uses"""
    with H.reference_modules() as ref:
        env = ref.Checkers.Checkers(None)
        rng = np.random.RandomState(plies)
        pos = codec.encode_state(env.state)
        for _ in range(plies):
            kids, _, status, _ = O.movegen(pos)
            if status != 0 or env.done:
                break
            i = rng.randint(len(env.legal_next_states))
            env.step(env.legal_next_states[i])
            pos = kids[i]
        assert not env.done
        ref.MCTS.MCTS(GAME_ENV=env, UCT_C=4, CONSTRAINT='rollout', BUDGET=budget, MULTIPROC=False, NEURAL_NET=False,
                      VERBOSE=False, TRAINING=False, DIRICHLET_ALPHA=1.0, DIRICHLET_EPSILON=0.0,
                      TEMPERATURE_TAU=0, TEMPERATURE_DECAY=0, TEMP_DECAY_DELAY=0)
        root = ref.MCTS.MCTS_Node(env.state)
        root.history = list(env.history)
        with hashed_playouts():
            ref.MCTS.MCTS.begin_tree_search(root)
        parent_player = int(env.history[-2][4, 0, 0])
        t = O.Tree(pos, O.make_cfg(budget=budget, rollout="hash"), parent_player=parent_player)
        t.search(budget)
        n, w = t.root_stats()
        assert (n, float(w)) == (root.n, float(root.w))
        assert [(c["n"], float(c["w"])) for c in t.root_children()] == [(c.n, float(c.w)) for c in root.children]
        assert [codec.meta_action(c["pos"][3]) for c in t.root_children()] == \
               [codec.action_id(*[int(v) for v in c.state[14, 0, 0:3]]) for c in root.children]


def _dict_literals(path, names):
    """dict literals assigned to `names` in a reference script, without executing it; values that are plain
    names (TRAINING_ITERATION, NN_FN, ...) come back as the name in angle brackets"""
    import ast
    out = {}
    for node in ast.walk(ast.parse(open(path).read())):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) \
                and node.targets[0].id in names and isinstance(node.value, ast.Dict):
            d = {}
            for k, v in zip(node.value.keys, node.value.values):
                d[ast.literal_eval(k)] = '<%s>' % v.id if isinstance(v, ast.Name) else ast.literal_eval(v)
            out[node.targets[0].id] = d
    return out


@pytest.mark.needs_reference
def test_driver_defaults_equal_the_reference_scripts():
    """train_Checkers.py / play_Checkers.py of this package carry the reference scripts' settings
    (train_Checkers.py:80-127,179-201; play_Checkers.py:88-103) as defaults"""
    import play_Checkers as P
    import train_Checkers as TC
    ref = _dict_literals(os.path.join(H.REFERENCE_DIR, "train_Checkers.py"),
                         ("selfplay_kwargs", "mcts_kwargs", "training_kwargs", "tourney_kwargs", "tourney_mcts_kwargs"))
    names = {'<TRAINING_ITERATION>': 7, '<NN_FN>': 'old.h5', '<NEW_NN_FN>': 'new.h5', '<NEURAL_NET>': True}

    def subst(d):
        return {k: names.get(v, v) if isinstance(v, str) else v for k, v in d.items()}
    assert TC.default_selfplay_kwargs(7, 'old.h5') == subst(ref["selfplay_kwargs"])
    assert TC.default_mcts_kwargs(7) == subst(ref["mcts_kwargs"])
    assert TC.default_mcts_kwargs(0)['NEURAL_NET'] is False               # train_Checkers.py:78
    assert TC.default_training_kwargs(7) == subst(ref["training_kwargs"])
    assert TC.default_tourney_kwargs(7, 'old.h5', 'new.h5') == subst(ref["tourney_kwargs"])
    assert TC.default_tourney_mcts_kwargs('new.h5') == subst(ref["tourney_mcts_kwargs"])
    play = _dict_literals(os.path.join(H.REFERENCE_DIR, "play_Checkers.py"), ("mcts_kwargs",))["mcts_kwargs"]
    play.pop('NN_FN')
    assert P.DEFAULT_MCTS_KWARGS == play
