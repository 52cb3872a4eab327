"""GPU tests of the reference-facing Python surface (Checkers / MCTS / training_pipeline drop-ins):
same names, arguments, file formats and errors as the reference, results identical to the golden
vectors generated from the unmodified reference."""
import json
import os
import pickle

import numpy as np
import pytest

from helpers import GOLDEN, codec
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from ckb200 import lib as L
    L.require_device()
    import Checkers as C
    import MCTS as M
    import training_pipeline as T
    assert os.path.dirname(C.__file__).endswith("checkers-mcts_b200")
    return C, M, T


MCTS_KW = dict(UCT_C=4, CONSTRAINT='rollout', MULTIPROC=False, NEURAL_NET=True, VERBOSE=False, TRAINING=False,
               DIRICHLET_ALPHA=1.0, DIRICHLET_EPSILON=0.0, TEMPERATURE_TAU=0, TEMPERATURE_DECAY=0, TEMP_DECAY_DELAY=0)


def test_checkers_env_matches_oracle(mods):
    C, _, _ = mods
    env = C.Checkers()
    assert env.state.shape == (15, 8, 8) and env.state.dtype == np.float64
    assert len(env.legal_next_states) == 7 and env.current_player(env.state) == 'player1'
    rng = np.random.RandomState(3)
    for _game in range(3):
        env.reset()
        pos = O.start_position()
        while not env.done:
            kids, mask, status, p5 = O.movegen(pos)
            assert len(env.legal_next_states) == len(kids)
            for st, k in zip(env.legal_next_states, kids):
                assert (st == codec.decode_state(k)).all()
            assert [codec.plane_to_bits(env.state[6 + i]) for i in range(8)] == mask
            assert env.state[5, 0, 0] == p5 / 80
            i = rng.randint(len(kids))
            state, outcome, done = env.step(env.legal_next_states[i])
            pos = kids[i]
            assert state is env.state and env.history[-1] is state
        assert env.outcome in ('player1_wins', 'player2_wins', 'draw')
        assert env.get_legal_next_states(env.history) == []
    env.reset()
    with pytest.raises(ValueError, match='Illegal next state'):
        env.step(np.zeros((15, 8, 8)))


def test_mcts_shim_first_search_and_reroot(mods):
    C, M, _ = mods
    from ckb200.net import StubNet
    kat = json.load(open(os.path.join(GOLDEN, "mcts_kat.json")))["hash_first_search"]
    env = C.Checkers(StubNet("hash"))
    M.MCTS(GAME_ENV=env, BUDGET=kat["budget"], **MCTS_KW)
    root = M.MCTS_Node(env.state)
    assert root.player == 'player1' and not root.terminal and root.depth == 1
    M.MCTS.begin_tree_search(root)
    assert root.n == kat["root_n"] and float(root.w) == kat["root_w"]
    for c, r in zip(root.children, kat["children"]):
        assert (int(c.state[14, 0, 0]) - 6) * 64 + int(c.state[14, 0, 1]) * 8 + int(c.state[14, 0, 2]) == r["action"]
        assert c.n == r["n"] and float(c.w) == r["w"] and float(c.p) == r["p"] and c.parent is root
    best = M.MCTS.best_child(root)
    assert best.n == max(c.n for c in root.children)
    env.step(best.state)
    reply = M.MCTS.best_child(best) if best.children else None
    assert reply is not None
    env.step(env.legal_next_states[[i for i, s in enumerate(env.legal_next_states)
                                    if (s[:5] == reply.state[:5]).all()][0]])
    new_root = M.MCTS.new_root_node(best)
    assert (new_root.state[:5] == env.state[:5]).all() and new_root.parent is None
    inherited = new_root.n
    M.MCTS.begin_tree_search(new_root)
    assert new_root.n == inherited + kat["budget"]            # BUDGET new sims on top of inherited statistics
    assert sum(c.n for c in new_root.children) == new_root.n - 1
    # a move that is not in the tree raises like the reference
    env2 = C.Checkers(StubNet("hash"))
    M.MCTS(GAME_ENV=env2, BUDGET=8, **MCTS_KW)
    r2 = M.MCTS_Node(env2.state)
    M.MCTS.begin_tree_search(r2)
    b2 = M.MCTS.best_child(r2)
    env2.step(b2.state)
    env2.step(env2.legal_next_states[-1])
    env2.step(env2.legal_next_states[-1])
    with pytest.raises(ValueError, match='All child nodes should be visited'):
        M.MCTS.new_root_node(b2)
    with pytest.raises(TypeError):
        env3 = C.Checkers(object())
        M.MCTS(GAME_ENV=env3, BUDGET=8, **MCTS_KW)
        M.MCTS.begin_tree_search(M.MCTS_Node(env3.state))


def _action(state):
    return (int(state[14, 0, 0]) - 6) * 64 + int(state[14, 0, 1]) * 8 + int(state[14, 0, 2])


@pytest.mark.parametrize("engine_options", [{}, {"compact_always": True}], ids=["default", "compact-every-reroot"])
def test_play_loop_two_trees_matches_reference_game(mods, engine_options):
    """play_Checkers' game loop (one tree per player, re-rooted through the opponent's reply) through the MCTS /
    MCTS_Node views: root and child statistics of every search equal the reference's own game
    (mcts_kat.json: _generate_data run verbatim, two trees, 40 plies), also when the device tree is
    renumbered at every re-rooting."""
    import play_Checkers as P
    from ckb200.net import StubNet
    kat = json.load(open(os.path.join(GOLDEN, "mcts_kat.json")))["hash_game"]
    log = []

    def spy(root, best):
        log.append((root.n, float(root.w), [(_action(c.state), c.n, float(c.w), float(c.p), bool(c.terminal))
                                             for c in root.children], [int(v) for v in best.state[14, 0, 0:3]]))

    kw = dict(MCTS_KW, BUDGET=kat["budget"], TRAINING=True, ENGINE_OPTIONS=engine_options)
    outcome, plies = P.play(StubNet("hash"), kw, max_plies=kat["terminate_cnt"], quiet=True, on_search=spy)
    assert plies == kat["terminate_cnt"] and len(log) == len(kat["moves"]) and outcome is None
    for i, (got, want) in enumerate(zip(log, kat["moves"])):
        assert got[0] == want["root_n"] and got[1] == want["root_w"], i
        assert got[2] == [(c["action"], c["n"], c["w"], c["p"], c["terminal"]) for c in want["children"]], i
    assert [g[3] for g in log[:len(kat["chosen"])]] == kat["chosen"]
    assert len(mods[1].MCTS._trees) == 2                       # one device tree per player


def test_iteration0_selfplay_without_network(mods, tmp_path, monkeypatch):
    """train_Checkers.py iteration 0: NEURAL_NET=False self-play through generate_Checkers_data and the MCTS
    shim -- same records as the reference's own _generate_data game (uct_kat.json, hashed playouts)"""
    C, M, T = mods
    import play_Checkers as P
    from ckb200.net import StubNet
    gk = json.load(open(os.path.join(GOLDEN, "uct_kat.json")))["game"]
    monkeypatch.chdir(tmp_path)
    os.makedirs("data/training_data")
    sp = dict(NUM_SELFPLAY_GAMES=1, TRAINING_ITERATION=0, TERMINATE_CNT=gk["terminate_cnt"], NUM_CPUS=1, NN_FN=None, SEED=3)
    mk = dict(MCTS_KW, BUDGET=gk["budget"], TRAINING=True, NEURAL_NET=False, PLAYOUT_EVALUATOR='rollout_hash')
    data = pickle.load(open(T.generate_Checkers_data(sp, mk).generate_data(), 'rb'))
    assert len(data) == len(gk["q"])
    for e, m, q, z in zip(data, gk["moves"], gk["q"], gk["z"]):
        assert e[2] == q and e[3] == z and e[0].shape == (15, 8, 8) and e[1].shape == (8, 8, 8)
        total = sum(c["n"] for c in m["children"])
        for c in m["children"]:
            assert e[1].reshape(512)[c["action"]] == c["n"] / total
    # the same game through MCTS / MCTS_Node views (two trees), children lists grow one node per visit
    log = []
    kw = dict(MCTS_KW, BUDGET=gk["budget"], TRAINING=True, NEURAL_NET=False)
    P.play(StubNet("rollout_hash"), kw, max_plies=12, quiet=True,
           on_search=lambda root, best: log.append((root.n, float(root.w), [(_action(c.state), c.n, float(c.w)) for c in root.children])))
    for got, want in zip(log, gk["moves"]):
        assert got == (want["root_n"], want["root_w"], [(c["action"], c["n"], c["w"]) for c in want["children"]])
    # CONSTRAINT='time': BUDGET seconds of searching
    env = C.Checkers(StubNet("hash"))
    M.MCTS(GAME_ENV=env, BUDGET=0.2, **dict(MCTS_KW, CONSTRAINT='time'))
    root = M.MCTS_Node(env.state)
    M.MCTS.begin_tree_search(root)
    assert root.n == M.MCTS.rollout_count >= M.MCTS.time_check_sims and root.n % M.MCTS.time_check_sims == 0
    with pytest.raises(ValueError, match='Invalid MCTS computational constraint'):
        M.MCTS(GAME_ENV=env, BUDGET=3, **dict(MCTS_KW, CONSTRAINT='memory'))
    env = C.Checkers(None)
    M.MCTS(GAME_ENV=env, BUDGET=3, **dict(MCTS_KW, NEURAL_NET=False))
    root = M.MCTS_Node(env.state)
    M.MCTS.begin_tree_search(root)
    assert root.n == 3 and len(root.children) == 3 and len(env.legal_next_states) == 7
    left = root.unvisited_child_states                          # the reference pops from the end of the legal list
    assert len(left) == 4 and all((a == b).all() for a, b in zip(left, env.legal_next_states[:4]))


def test_train_checkers_iteration_loop(mods, tmp_path, monkeypatch):
    """train_Checkers.py end to end at toy size: iteration 0 (playout self-play -> first network -> tournament
    against the untrained one), iteration 1 self-play with the trained network, final round-robin"""
    import train_Checkers as TC
    monkeypatch.chdir(tmp_path)
    small_mcts = {'BUDGET': 16}
    out = TC.run_iteration(0, SELFPLAY=True, TRAINING=True, EVALUATION=True,
                           selfplay_kwargs={'NUM_SELFPLAY_GAMES': 3, 'TERMINATE_CNT': 30, 'NUM_CPUS': 2, 'SEED': 11},
                           mcts_kwargs=small_mcts, training_kwargs={'EPOCHS': 2, 'BATCH_SIZE': 32},
                           tourney_kwargs={'TOURNEY_GAMES': 2, 'NUM_CPUS': 1, 'SEED': 5}, tourney_mcts_kwargs=small_mcts)
    assert len(out['data_fns']) == 2 and len(out['history']['loss']) == 2
    merged = [fn for fn in os.listdir('data/training_data') if 'Data0_' in fn]
    assert len(merged) == 1                                      # the per-worker files were merged into one
    data = pickle.load(open('data/training_data/' + merged[0], 'rb'))
    assert len(data) >= 6 * 20 and all(len(e) == 4 for e in data)
    for key in ('OLD_NN_FN', 'NEW_NN_FN', 'plot', 'tourney_fn'):
        assert os.path.isfile(out[key]), key
    assert 'Model0_' in out['OLD_NN_FN'] and 'Model1_' in out['NEW_NN_FN']
    assert 'Wins/Losses/Draws' in open(out['tourney_fn'], encoding='utf-8').read()
    out1 = TC.run_iteration(1, NN_FN=out['NEW_NN_FN'], SELFPLAY=True,
                            selfplay_kwargs={'NUM_SELFPLAY_GAMES': 2, 'TERMINATE_CNT': 12, 'SEED': 12}, mcts_kwargs=small_mcts)
    assert len(pickle.load(open(out1['data_fns'], 'rb'))) >= 2 * 12
    fe = TC.run_final_evaluation([0, 1], tourney_kwargs={'NUM_CPUS': 1, 'SEED': 9}, tourney_mcts_kwargs=small_mcts)
    assert os.path.isfile(fe) and 'Total' in open(fe, encoding='utf-8').read()


def test_play_loop_human_input(mods, capsys):
    import play_Checkers as P
    from ckb200.net import StubNet
    answers = iter(["99", "1", "2", "1", "1", "1", "1"])
    outcome, plies = P.play(StubNet("uniform_material"), dict(MCTS_KW, BUDGET=30), human_player2=True, max_plies=6,
                            print_trees=True, tree_depth=1, input_fn=lambda prompt: next(answers))
    out = capsys.readouterr().out
    assert plies == 6 and outcome is None
    assert 'Invalid selection!  Try again!' in out and 'Option #1: (' in out and '|- (' in out
    env = mods[0].Checkers(StubNet("hash"))
    assert P.states_to_piece_positions(env.state, env.legal_next_states)[0] == [(3, 2), (4, 3)]


def test_generate_data_pickle_matches_reference_golden(mods, tmp_path, monkeypatch):
    _, _, T = mods
    f = np.load(os.path.join(GOLDEN, "selfplay_hash.npz"))
    meta = json.loads(str(f["meta"]))
    monkeypatch.chdir(tmp_path)
    os.makedirs("data/training_data")
    kw = dict(MCTS_KW, BUDGET=meta["budget"], TRAINING=True)
    gen = T.generate_Checkers_data(dict(NUM_SELFPLAY_GAMES=1, TRAINING_ITERATION=3, TERMINATE_CNT=meta["terminate_cnt"],
                                        NUM_CPUS=1, NN_FN="stub:hash"), kw)
    fn = gen.generate_data()
    assert fn.startswith("data/training_data/Checkers_Data3_") and fn.endswith("_P0.pkl")
    data = T.load_training_data(fn)
    assert len(data) == len(f["q"])
    for i, (state, probs, q, z) in enumerate(data):
        assert state.shape == (15, 8, 8) and state.dtype == np.float64 and probs.shape == (8, 8, 8)
        planes = [codec.plane_to_bits(state[j]) for j in (0, 1, 2, 3, 6, 7, 8, 9, 10, 11, 12, 13)]
        assert planes == [int(v) for v in f["planes"][i]]
        assert int(state[4, 0, 0]) == f["player"][i] and (state[4] == state[4, 0, 0]).all()
        assert state[5, 0, 0] == f["plane5"][i] / 80
        assert [int(v) for v in state[14, 0, 0:3]] == [int(v) for v in f["action"][i]]
        assert probs.reshape(512).tobytes() == f["probs"][i].tobytes()
        assert float(q) == f["q"][i] and z == f["z"][i]
    # several workers -> one file per worker, games split as in the reference
    gen = T.generate_Checkers_data(dict(NUM_SELFPLAY_GAMES=2, TRAINING_ITERATION=0, TERMINATE_CNT=20, NUM_CPUS=3,
                                        NN_FN="stub:hash_salted"), dict(kw, BUDGET=20))
    fns = gen.generate_data()
    assert len(fns) == 3 and all(len(T.load_training_data(x)) == 2 * 20 for x in fns)
    T.record_params('selfplay', A=1)
    with pytest.raises(ValueError):
        T.record_params('nonsense')


def test_tournament_matches_reference_golden(mods, tmp_path, monkeypatch):
    _, _, T = mods
    t = json.load(open(os.path.join(GOLDEN, "tournament.json")))
    monkeypatch.chdir(tmp_path)
    os.makedirs("data/tournament_results")
    kinds = {v: k for k, v in {"A": "stub:" + t["nets"]["data/model/A"], "B": "stub:" + t["nets"]["data/model/B"]}.items()}
    tour = T.tournament_Checkers(dict(NEW_NN_FN="stub:" + t["nets"]["data/model/A"], OLD_NN_FN="stub:" + t["nets"]["data/model/B"],
                                      TOURNEY_GAMES=len(t["outcomes"]), NUM_CPUS=1), dict(MCTS_KW, BUDGET=t["budget"]))
    rows = tour._start_tournament()
    for row, ref in zip(rows, t["outcomes"]):
        assert [row[0], kinds[row[1]], kinds[row[2]], row[3], row[4]] == ref
    fn = tour.start_tournament()
    txt = open(fn, encoding="utf-8").read()
    assert fn.startswith("data/tournament_results/Tournament_") and "Wins/Losses/Draws" in txt and "Turn Count" in txt


def test_final_evaluation_round_robin(mods, tmp_path, monkeypatch):
    """final_evaluation (reference training_pipeline.py:603-719): three models, every pairing plays two
    games; the pairing table is antisymmetric and the points are its row sums."""
    _, _, T = mods
    from ckb200 import net as N
    monkeypatch.chdir(tmp_path)
    os.makedirs("data/model")
    for it in (0, 1, 2):
        np.save("data/model/Checkers_Model%d_01-Jan-2021(00:00:0%d).npy" % (it, it), N.random_init_blob(it))
    fe = T.final_evaluation([0, 1, 2], dict(TRAINING_ITERATION=2, OLD_NN_FN=None, NEW_NN_FN=None, TOURNEY_GAMES=2, NUM_CPUS=1, SEED=7),
                            dict(MCTS_KW, BUDGET=6, NEURAL_NET=True, TRAINING=False, DIRICHLET_EPSILON=0.25, TEMPERATURE_TAU=0))
    assert [f.split("_")[1] for f in fe.model_fn_list] == ["Model0", "Model1", "Model2"]
    fn = fe.start_evaluation(num_cpus=4)
    assert sum(len(g) for g in fe.game_outcomes) == 6                     # 3 pairings x 2 games
    assert (fe.table == -fe.table.T).all() and (fe.model_scores == fe.table.sum(1)).all()
    txt = open(fn, encoding="utf-8").read()
    assert fn.startswith("data/final_eval/Checkers_Final_Evaluation_") and "Total" in txt
    with pytest.raises(ValueError):
        T.final_evaluation([0, 5], dict(NUM_CPUS=1), MCTS_KW)


def test_generate_data_pooled_writers_equal_inline_conversion(mods, tmp_path, monkeypatch):
    """generate_data() at a size where the per-worker pickles are written by the process pool (ckb200.records.
    save_reference_pickles: > 20 000 records): every file loads as the reference's list format and equals the
    inline conversion of the same engine records, worker by worker."""
    from ckb200 import lib as L
    from ckb200 import records as R
    _, _, T = mods
    monkeypatch.chdir(tmp_path)
    os.makedirs("data/training_data")
    sp = dict(NUM_SELFPLAY_GAMES=80, TRAINING_ITERATION=2, TERMINATE_CNT=110, NUM_CPUS=4, NN_FN='stub:hash_salted', SEED=31)
    mk = dict(MCTS_KW, BUDGET=6, TRAINING=True, DIRICHLET_EPSILON=0.25, TEMPERATURE_TAU=1.0, TEMPERATURE_DECAY=0.1, TEMP_DECAY_DELAY=10)
    gen = T.generate_Checkers_data(sp, mk)
    fns = gen.generate_data()
    assert len(fns) == 4 and gen.n_records > 20000
    eng = L.Engine(L.make_cfg(n_slots=320, budget=6, training=True, terminate_cnt=110, evaluator="hash_salted", seed=31, uct_c=MCTS_KW['UCT_C'],
                              alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10))
    eng.selfplay(320)
    recs = eng.records()
    eng.close()
    recs = recs[np.argsort(recs["game"], kind="stable")]
    assert len(recs) == gen.n_records
    total = 0
    for p, fn in enumerate(fns):
        data = pickle.load(open(fn, "rb"))
        want = R.to_reference_list(recs[(recs["game"] >= 80 * p) & (recs["game"] < 80 * (p + 1))])
        assert len(data) == len(want) > 0
        for a, b in zip(data[:50] + data[-50:], want[:50] + want[-50:]):
            assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2] and a[3] == b[3] and type(a[2]) is type(b[2])
        total += len(data)
    assert total == gen.n_records
