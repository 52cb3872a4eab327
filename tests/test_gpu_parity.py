"""GPU parity tests: the CUDA path, called through the C ABI (ckb200.lib -> libckb200.so), against
the CPU oracle on the same inputs and against the committed golden vectors (generated from the
unmodified reference).  Integer / index work and the deterministic tree (epsilon = 0, tau = 0)
are bit-exact; the network is checked within 1e-5 (north_star tolerance)."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, PKG, ROOT, codec, record_planes
from oracle import oracle as O

pytestmark = pytest.mark.gpu

NET_TOL = 1e-5          # absolute, on softmax policy outputs and the tanh value


@pytest.fixture(scope="module")
def lib():
    from ckb200 import lib as L
    L.require_device()
    return L


def _pos_arr(L, positions):
    a = np.zeros(len(positions), dtype=L.POS_DTYPE)
    for i, p in enumerate(positions):
        a[i] = tuple(int(v) for v in p)
    return a


def _check_movegen(L, positions):
    out = L.movegen(_pos_arr(L, positions))
    for i, p in enumerate(positions):
        kids, mask, status, p5 = O.movegen(p)
        assert out["counts"][i] == len(kids)
        got = [tuple(int(v) for v in out["children"][i, j]) for j in range(len(kids))]
        assert got == kids
        assert [int(v) for v in out["masks"][i]] == mask
        assert int(out["status"][i]) == status and int(out["plane5"][i]) == p5


# ---- K1 -------------------------------------------------------------------------------------
def test_movegen_golden(lib):
    g = np.load(os.path.join(GOLDEN, "movegen_cases.npz"))
    pos = g["pos"]
    out = lib.movegen(_pos_arr(lib, pos))
    off = np.concatenate([[0], np.cumsum(g["counts"])])
    assert (out["counts"] == g["counts"]).all()
    assert (out["masks"] == g["mask"]).all()
    assert (out["status"] == g["status"]).all() and (out["plane5"] == g["plane5"]).all()
    for i in range(len(pos)):
        ref = g["kids"][off[i]:off[i + 1]]
        ch = out["children"][i, :len(ref)]
        assert (ch["p1"] == ref[:, 0]).all() and (ch["p2"] == ref[:, 1]).all() and (ch["k"] == ref[:, 2]).all()
        assert ((ch["meta"] & 1) == ref[:, 3]).all() and (((ch["meta"] >> 8) & 0x1FF) == ref[:, 4]).all()


def test_movegen_random_walks_vs_oracle(lib):
    rng = np.random.RandomState(123)
    positions = []
    for _ in range(80):
        pos = O.start_position()
        for _ply in range(300):
            positions.append(pos)
            kids, _, st, _ = O.movegen(pos)
            if st != codec.ONGOING:
                break
            pos = kids[rng.randint(len(kids))]
    assert len(positions) > 4000
    _check_movegen(lib, positions)


def test_movegen_edge_cases(lib):
    # empty batch, a side without pieces, a blocked side, maximum king mobility, late-ply draw window
    assert len(lib.movegen(np.zeros(0, dtype=lib.POS_DTYPE))["counts"]) == 0
    cases = [
        (0x00000FFF, 0, 0, codec.make_meta(1)),
        (0, 0xFFF00000, 0, codec.make_meta(0)),
        (0x1, 0x30, 0, codec.make_meta(0)),
        (0x0F0F0F00, 0x1, 0x0F0F0F01, codec.make_meta(0, 79, 0, 0, 120)),
        (0x00000400, 0x00200000, 0x00200400, codec.make_meta(1, 78, 0, 0, 79)),
        (0x00000400, 0x00200000, 0x00200400, codec.make_meta(1, 90, 0, 0, 300)),
    ]
    _check_movegen(lib, cases)
    out = lib.movegen(_pos_arr(lib, cases[3:4]))
    assert out["counts"][0] > 24       # twelve kings


def test_movegen_large_batch_properties(lib):
    """full-size sweep (2^20 positions): every position's outputs equal those of its duplicates and
    the batch result does not depend on the batch it was computed in."""
    rng = np.random.RandomState(5)
    base = []
    pos = O.start_position()
    for _ in range(64):
        base.append(pos)
        kids, _, st, _ = O.movegen(pos)
        if st != codec.ONGOING:
            pos = O.start_position()
        else:
            pos = kids[rng.randint(len(kids))]
    small = _pos_arr(lib, base)
    ref = lib.movegen(small)
    idx = rng.randint(0, len(base), size=1 << 20)
    big = lib.movegen(small[idx], want_children=False)
    assert (big["counts"] == ref["counts"][idx]).all()
    assert (big["masks"] == ref["masks"][idx]).all()
    assert (big["status"] == ref["status"][idx]).all()


def test_movegen_csr_matches_strided_and_oracle(lib):
    """packed (CSR) K1: same successors, in the same order, as the strided kernel and the oracle, for
    ragged batches (empty, one position, tile boundaries of 256, many tiles for the look-back chain)."""
    rng = np.random.RandomState(11)
    walk = []
    pos = O.start_position()
    for _ in range(700):
        walk.append(pos)
        kids, _, st, _ = O.movegen(pos)
        pos = O.start_position() if st != codec.ONGOING else kids[rng.randint(len(kids))]
    arr = _pos_arr(lib, walk)
    out0 = lib.movegen_csr(arr[:0])
    assert list(out0["offsets"]) == [0] and len(out0["children"]) == 0
    for n in (1, 255, 256, 257, 700):
        sub = arr[:n]
        ref = lib.movegen(sub)
        out = lib.movegen_csr(sub)
        off = out["offsets"].astype(np.int64)
        assert (np.diff(off) == ref["counts"]).all() and off[0] == 0 and off[-1] == len(out["children"])
        assert (out["masks"] == ref["masks"]).all() and (out["status"] == ref["status"]).all()
        assert (out["plane5"] == ref["plane5"]).all()
        for i in range(n):
            a = out["children"][off[i]:off[i + 1]]
            assert a.tobytes() == ref["children"][i, :ref["counts"][i]].tobytes()
    for i in rng.randint(0, 700, size=40):                       # and directly against the oracle
        kids, mask, st, p5 = O.movegen(walk[i])
        a = out["children"][off[i]:off[i + 1]]
        assert [tuple(int(v) for v in c) for c in a] == kids
    # 2^20 positions: 4096 tiles chained by the look-back; offsets must be the exact prefix sums
    idx = rng.randint(0, 700, size=1 << 20)
    big = lib.movegen_csr(arr[idx])
    cnt = ref["counts"][idx].astype(np.int64)
    exp = np.concatenate([[0], np.cumsum(cnt)])
    assert (big["offsets"].astype(np.int64) == exp).all()
    for i in rng.randint(0, 1 << 20, size=200):
        a = big["children"][exp[i]:exp[i + 1]]
        assert a.tobytes() == ref["children"][idx[i], :cnt[i]].tobytes()


# ---- Checkers.predict glue -----------------------------------------------------------------
def test_mask_renorm_golden(lib):
    g = np.load(os.path.join(GOLDEN, "predict_glue.npz"))
    out = lib.mask_renorm(g["policy"], g["mask"])
    assert out.tobytes() == np.ascontiguousarray(g["prior"], dtype=np.float32).tobytes()


# ---- K2: deterministic tree ------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["uniform_zero", "uniform_material", "hash"])
def test_first_search_golden(lib, kind):
    kat = json.load(open(os.path.join(GOLDEN, "mcts_kat.json")))[kind + "_first_search"]
    eng = lib.Engine(lib.make_cfg(n_slots=1, budget=kat["budget"], evaluator=kind, keep_records=False))
    eng.tree_set_root(O.start_position())
    eng.tree_search(kat["budget"])
    n, w, b = eng.tree_root()
    assert n == kat["root_n"] and float(w) == kat["root_w"] and b == len(kat["children"])
    for c, r in zip(eng.tree_root_children(), kat["children"]):
        assert codec.meta_action(c["pos"][3]) == r["action"]
        assert c["n"] == r["n"] and float(c["w"]) == r["w"] and float(c["p"]) == r["p"]
    eng.close()


def test_tree_api_advance_matches_oracle(lib):
    """search, advance to the robust child, search again (tree reuse) == oracle tree."""
    cfg = O.make_cfg(budget=150)
    t = O.Tree(O.start_position(), cfg, "hash")
    eng = lib.Engine(lib.make_cfg(n_slots=1, budget=150, evaluator="hash", keep_records=False))
    eng.tree_set_root(O.start_position())
    t.search(150)
    eng.tree_search(150)
    got, ref = eng.tree_root_children(), t.root_children()
    assert [(c["n"], float(c["w"]), float(c["p"])) for c in got] == [(c["n"], float(c["w"]), float(c["p"])) for c in ref]
    assert eng.tree_best_child() == t.best_child()
    assert eng.tree_node_count() == t.node_count()
    eng.close()


def _engine_records(lib, eng):
    recs = eng.records()
    games = eng.games()
    out = {}
    for r in recs:
        n = int(r["n_children"])
        out.setdefault(int(r["game"]), []).append(dict(
            pos=tuple(int(v) for v in r["pos"]), mask=[int(v) for v in r["mask"]], plane5=int(r["plane5"]),
            actions=[int(v) for v in r["action"][:n]], visits=[int(v) for v in r["visits"][:n]],
            q=np.float32(r["q"]), z=int(r["z"]), root_n=int(r["root_n"]), root_w=np.float32(r["root_w"]),
            chosen=int(r["chosen"])))
    return out, {int(g["game"]): g for g in games}


def _same_record(a, b):
    return (a["pos"] == b["pos"] and a["mask"] == b["mask"] and a["plane5"] == b["plane5"]
            and a["actions"] == b["actions"] and a["visits"] == b["visits"]
            and np.float32(a["q"]).tobytes() == np.float32(b["q"]).tobytes() and a["z"] == b["z"]
            and a["root_n"] == b["root_n"] and np.float32(a["root_w"]).tobytes() == np.float32(b["root_w"]).tobytes()
            and a["chosen"] == b["chosen"])


@pytest.mark.parametrize("kind", ["uniform_zero", "uniform_material", "hash"])
def test_game_opening_golden(lib, kind):
    kat = json.load(open(os.path.join(GOLDEN, "mcts_kat.json")))[kind + "_game"]
    eng = lib.Engine(lib.make_cfg(n_slots=1, budget=kat["budget"], training=True, terminate_cnt=kat["terminate_cnt"],
                                  evaluator=kind))
    eng.selfplay(1)
    recs, _ = _engine_records(lib, eng)
    recs = recs[0]
    assert len(recs) == len(kat["moves"])
    for r, m in zip(recs, kat["moves"]):
        assert r["root_n"] == m["root_n"] and float(r["root_w"]) == m["root_w"]
        assert r["actions"] == [c["action"] for c in m["children"]]
        assert r["visits"] == [c["n"] for c in m["children"]]
    eng.close()


@pytest.mark.parametrize("kind", ["hash", "uniform_material"])
def test_selfplay_records_golden(lib, kind):
    f = np.load(os.path.join(GOLDEN, "selfplay_%s.npz" % kind))
    meta = json.loads(str(f["meta"]))
    eng = lib.Engine(lib.make_cfg(n_slots=1, budget=meta["budget"], training=True,
                                  terminate_cnt=meta["terminate_cnt"], evaluator=kind))
    eng.selfplay(1)
    recs, games = _engine_records(lib, eng)
    recs = recs[0]
    assert len(recs) == len(f["q"]) and games[0]["reroot_misses"] == 0
    for i, r in enumerate(recs):
        state, probs = record_planes(r)
        pl = [codec.plane_to_bits(state[j]) for j in (0, 1, 2, 3, 6, 7, 8, 9, 10, 11, 12, 13)]
        assert pl == [int(v) for v in f["planes"][i]]
        assert int(state[4, 0, 0]) == f["player"][i] and r["plane5"] == f["plane5"][i]
        assert [int(v) for v in state[14, 0, 0:3]] == [int(v) for v in f["action"][i]]
        assert probs.reshape(512).tobytes() == f["probs"][i].tobytes()
        assert float(r["q"]) == f["q"][i] and r["z"] == f["z"][i]
    eng.close()


def test_tournament_golden(lib):
    t = json.load(open(os.path.join(GOLDEN, "tournament.json")))
    kinds = t["nets"]
    rows = t["outcomes"]
    first = rows[0]
    eng = lib.Engine(lib.make_cfg(n_slots=2, budget=t["budget"], training=False, arena=True, keep_records=False,
                                  evaluator=kinds["data/model/" + first[1]], evaluator_p2=kinds["data/model/" + first[2]]))
    eng.arena(len(rows))
    games = {int(g["game"]): g for g in eng.games()}
    names = {1: "player1_wins", 2: "player2_wins", 3: "draw"}
    for i, (_num, p1, _p2, outcome, move_count) in enumerate(rows):
        assert names[int(games[i]["outcome"])] == outcome and int(games[i]["move_count"]) == move_count
        assert int(games[i]["p1_net"]) == (0 if p1 == first[1] else 1)
    eng.close()


def _oracle_games(n_games, budget, terminate_cnt, kind="hash_salted", kind2=None, training=True):
    out = []
    for g in range(n_games):
        gm = O.Game(O.make_cfg(budget=budget, training=training, terminate_cnt=terminate_cnt), kind, kind2, salt=g)
        gm.play()
        out.append((gm.records(), gm.outcome, gm.move_count, gm.terminated, gm.total_sims, gm.nn_evals))
        gm.close()
    return out


@pytest.mark.parametrize("slots,compact", [(24, False), (7, True)])
def test_concurrent_games_vs_oracle(lib, slots, compact):
    """many different games in flight (per-game salted stub evaluator), fewer slots than games so
    that slots are refilled; every record of every game must equal the oracle's bit for bit.
    The second variant forces the re-root compaction (K5) on every move with a small pool."""
    n_games, budget, term = 40, 48, 70
    ref = _oracle_games(n_games, budget, term)
    eng = lib.Engine(lib.make_cfg(n_slots=slots, budget=budget, training=True, terminate_cnt=term,
                                  evaluator="hash_salted", compact_always=compact, pool_cap=8192 if compact else 0))
    st = eng.selfplay(n_games)
    recs, games = _engine_records(lib, eng)
    assert len(games) == n_games
    total_sims = 0
    for g in range(n_games):
        rr, outcome, move_count, terminated, sims, evals = ref[g]
        assert int(games[g]["outcome"]) == outcome and int(games[g]["move_count"]) == move_count
        assert bool(games[g]["terminated"]) == terminated and int(games[g]["reroot_misses"]) == 0
        assert int(games[g]["sims"]) == sims and int(games[g]["nn_evals"]) == evals
        assert len(recs[g]) == len(rr)
        for a, b in zip(recs[g], rr):
            assert _same_record(a, b)
        total_sims += sims
    assert st["sims"] == total_sims and st["games_finished"] == n_games
    if compact:
        assert st["compactions"] > 0
    eng.close()


@pytest.mark.parametrize("cache,chain", [(-1, 0), (16, 1), (64, 3), (0, 0), (0, 64)])
def test_eval_cache_is_transparent(lib, cache, chain):
    """the evaluation cache (a leaf whose network input was evaluated before in the same slot is expanded from the
    stored priors / value inside the round) must not change a single bit of any record: no cache, a 16-entry cache
    that evicts all the time, the default, different per-round chain caps -- all equal the oracle, which evaluates
    every leaf.  Games follow each other in a slot, so entries of a finished game are probed by the next one
    (salted evaluator: they must not hit)."""
    n_games, budget, term = 12, 64, 60
    ref = _oracle_games(n_games, budget, term)
    eng = lib.Engine(lib.make_cfg(n_slots=3, budget=budget, training=True, terminate_cnt=term, evaluator="hash_salted",
                                  eval_cache_entries=cache, max_chain_per_step=chain))
    st = eng.selfplay(n_games)
    recs, games = _engine_records(lib, eng)
    for g in range(n_games):
        rr, outcome, move_count, terminated, sims, evals = ref[g]
        assert int(games[g]["outcome"]) == outcome and int(games[g]["move_count"]) == move_count
        assert int(games[g]["sims"]) == sims and int(games[g]["nn_evals"]) == evals
        assert len(recs[g]) == len(rr) and all(_same_record(a, b) for a, b in zip(recs[g], rr))
    if cache < 0:
        assert st["cache_hits"] == 0
    else:
        assert 0 < st["cache_hits"] < st["nn_evals"]
        if cache == 0:
            assert st["cache_hits"] > 0.2 * st["nn_evals"]      # the two colours' trees repeat each other's evaluations
    eng.close()


def test_eval_cache_with_network_and_weight_change(lib):
    """network evaluator: identical records with and without the cache (noise and temperature on), and new weights on
    the same net object invalidate the cached evaluations"""
    from ckb200 import net as N
    net = lib.Net(0)

    def run(blob_seed, cache):
        net.set_weights(N.random_init_blob(blob_seed))
        eng = lib.Engine(lib.make_cfg(n_slots=16, budget=100, training=True, terminate_cnt=6, evaluator="net", uct_c=4.0, alpha=1.0,
                                      epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=5, eval_cache_entries=cache))
        eng.set_net(0, net)
        st = eng.selfplay(16)
        recs = eng.records()
        out = [st, recs[np.lexsort((recs["ply"], recs["game"]))]]
        # second batch of games on the SAME engine after a weight change: stale entries must not be used
        net.set_weights(N.random_init_blob(blob_seed + 1))
        st2 = eng.selfplay(16)
        recs = eng.records()
        out += [st2, recs[np.lexsort((recs["ply"], recs["game"]))]]
        eng.close()
        return out

    a = run(0, -1)
    b = run(0, 0)
    assert a[1].tobytes() == b[1].tobytes() and a[3].tobytes() == b[3].tobytes()
    assert a[1].tobytes() != a[3].tobytes()                      # the two weight sets do play differently
    assert a[0]["cache_hits"] == 0 and b[0]["cache_hits"] > 0 and b[2]["cache_hits"] > 0
    assert a[0]["nn_evals"] == b[0]["nn_evals"] and a[0]["sims"] == b[0]["sims"]
    net.close()


def test_baseline_size_records_vs_oracle(lib):
    """BASELINE configs[1] size -- 4096 concurrent games, 400 sims/move -- against the oracle: every game's evaluator is
    salted with its id, 64 sampled games (first, last, and spread over the slots) must equal the oracle's records bit
    for bit, i.e. nothing leaks between the 4096 trees, their caches and their batch rows at full size."""
    n_games, budget, term = 4096, 400, 5
    eng = lib.Engine(lib.make_cfg(n_slots=n_games, budget=budget, training=True, terminate_cnt=term, evaluator="hash_salted"))
    st = eng.selfplay(n_games)
    recs, games = _engine_records(lib, eng)
    eng.close()
    assert st["games_finished"] == n_games and st["sims"] == n_games * term * budget
    sample = sorted(set([0, 1, 2, 4095, 4094] + list(range(7, 4096, 69))))[:64]
    for g in sample:
        gm = O.Game(O.make_cfg(budget=budget, training=True, terminate_cnt=term), "hash_salted", None, salt=g)
        gm.play()
        rr = gm.records()
        assert int(games[g]["outcome"]) == gm.outcome and int(games[g]["move_count"]) == gm.move_count
        assert int(games[g]["sims"]) == gm.total_sims and int(games[g]["nn_evals"]) == gm.nn_evals
        assert len(recs[g]) == len(rr) and all(_same_record(a, b) for a, b in zip(recs[g], rr)), g
        gm.close()


def test_reference_tau_quirk(lib):
    """MCTS(**kwargs) runs once per worker process in the reference (training_pipeline.py:347), so the temperature is
    never reset: only a worker's first game samples with tau > 0, every later game is pure arg-max play.  With
    reference_tau_quirk the second game of a slot therefore equals the oracle's game at tau = 0; without it the
    temperature starts again at TEMPERATURE_TAU."""
    kw = dict(budget=60, training=True, terminate_cnt=30, evaluator="hash", tau=1.0, tau_decay=0.5, tau_decay_delay=2, seed=9)
    gm = O.Game(O.make_cfg(budget=60, training=True, terminate_cnt=30, tau=0.0), "hash", None)
    gm.play()
    ref = gm.records()
    gm.close()

    def second_game(quirk):
        eng = lib.Engine(lib.make_cfg(n_slots=1, reference_tau_quirk=quirk, **kw))
        eng.selfplay(2)
        recs, _games = _engine_records(lib, eng)
        eng.close()
        return recs[0], recs[1]

    first, second = second_game(True)
    assert len(second) == len(ref) and all(_same_record(a, b) for a, b in zip(second, ref))
    first2, second2 = second_game(False)
    assert all(_same_record(a, b) for a, b in zip(first, first2)) and len(first) == len(first2)     # the first game is the same either way
    chosen = lambda rr: [r["chosen"] for r in rr]
    assert chosen(second2)[:8] != chosen(ref)[:8] or chosen(second2) != chosen(ref)      # temperature back on: sampled moves


def test_batch_shaping_is_transparent():
    """the evaluator batch of a round is capped at whole tower waves and the leaves beyond the cap wait a round
    (round_begin_kernel / stage_leaf).  Forced here to a wave of 5 rows with the salted stub evaluator and 24 slots, so
    that a large share of the leaves is deferred every round: all records must still equal the oracle's bit for bit."""
    import subprocess
    import sys
    code = (
        "import sys, json; sys.path[:0] = %r\n"
        "import numpy as np\n"
        "from ckb200 import lib\n"
        "eng = lib.Engine(lib.make_cfg(n_slots=24, budget=48, training=True, terminate_cnt=70, evaluator='hash_salted'))\n"
        "st = eng.selfplay(40)\n"
        "r = eng.records(); r = r[np.lexsort((r['ply'], r['game']))]\n"
        "np.save(sys.argv[1], r); print(json.dumps(dict(steps=st['steps'], sims=st['sims'])))\n") % ([ROOT, PKG],)
    import tempfile
    outs = {}
    with tempfile.TemporaryDirectory() as tmp:
        for waves in ("0", "5"):
            fn = os.path.join(tmp, "r%s.npy" % waves)
            r = subprocess.run([sys.executable, "-c", code, fn], env=dict(os.environ, CK_BATCH_WAVES=waves), capture_output=True,
                               text=True, timeout=300)
            assert r.returncode == 0, r.stdout + r.stderr
            outs[waves] = (np.load(fn), json.loads(r.stdout.strip().splitlines()[-1]))
    a, b = outs["0"], outs["5"]
    assert a[0].tobytes() == b[0].tobytes() and a[1]["sims"] == b[1]["sims"]
    assert b[1]["steps"] > 1.1 * a[1]["steps"], (a[1], b[1])     # deferral really happened: the capped run needs more rounds
    ref = _oracle_games(40, 48, 70)
    k = 0
    for g in range(40):
        for rr in ref[g][0]:
            e = a[0][k]
            n = int(e["n_children"])
            assert tuple(int(v) for v in e["pos"]) == rr["pos"] and [int(v) for v in e["visits"][:n]] == rr["visits"]
            assert np.float32(e["root_w"]).tobytes() == np.float32(rr["root_w"]).tobytes()
            k += 1
    assert k == len(a[0])


def test_production_size_batch_shaping_is_transparent():
    """the shaping rule as it runs in production: 4096 slots, the network evaluator, the default 4 x SMs wave and 0.4-wave
    cut-back threshold (at cfg2 it caps every round of the steady state).  The records of 4096 warm-started games must be
    the same bytes with the rule switched off (CK_BATCH_WAVES=0) and with a different chain cap, and the shaped run must
    really have deferred leaves (it needs more rounds for the same games)."""
    import subprocess
    import sys
    code = (
        "import sys, json, hashlib; sys.path[:0] = %r\n"
        "import numpy as np\n"
        "from ckb200 import lib, net as N\n"
        "net = lib.Net(0); net.set_weights(N.random_init_blob(3))\n"
        "eng = lib.Engine(lib.make_cfg(n_slots=4096, budget=96, training=True, terminate_cnt=24, evaluator='net', seed=5,\n"
        "                              stagger_budget=8, stagger_plies=16, max_chain_per_step=int(sys.argv[1])))\n"
        "eng.set_net(0, net)\n"
        "st = eng.selfplay(4096)\n"
        "r = eng.records(); r = r[np.lexsort((r['ply'], r['game']))]\n"
        "print(json.dumps(dict(steps=st['steps'], sims=st['sims'], evals=st['nn_evals'] - st['cache_hits'], records=len(r),\n"
        "                      sha=hashlib.sha256(r.tobytes()).hexdigest())))\n") % ([ROOT, PKG],)
    outs = {}
    for name, waves, chain in (("off", "0", 0), ("default", None, 0), ("chain2", None, 2)):
        env = dict(os.environ)
        env.pop("CK_BATCH_WAVES", None)
        env.pop("CK_BATCH_SLACK10", None)
        if waves is not None:
            env["CK_BATCH_WAVES"] = waves
        r = subprocess.run([sys.executable, "-c", code, str(chain)], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[name] = json.loads(r.stdout.strip().splitlines()[-1])
    off, dflt, chain2 = outs["off"], outs["default"], outs["chain2"]
    assert off["records"] == dflt["records"] == chain2["records"] and off["records"] >= 4096 * 8
    assert off["sha"] == dflt["sha"] == chain2["sha"], outs
    assert off["sims"] == dflt["sims"] == chain2["sims"]
    assert dflt["steps"] > off["steps"], outs            # leaves were deferred: more rounds, each with fewer tower iterations


def test_overlapped_groups_are_transparent():
    """CK_OVERLAP=1: the slots form two groups and the tree kernel of one runs on a second stream next to the tower of the
    other (one-warp blocks, per-group batches and counters).  Same records, bit for bit, as the single-stream engine
    (network evaluator, noise and temperature on, 2560 slots = the smallest size that enables the option)."""
    import subprocess
    import sys
    import tempfile
    code = (
        "import sys, json; sys.path[:0] = %r\n"
        "import numpy as np\n"
        "from ckb200 import lib, net as N\n"
        "net = lib.Net(0); net.set_weights(N.random_init_blob(0))\n"
        "eng = lib.Engine(lib.make_cfg(n_slots=2560, budget=60, training=True, terminate_cnt=4, evaluator='net', uct_c=4.0, alpha=1.0,\n"
        "                              epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=21))\n"
        "eng.set_net(0, net)\n"
        "st = eng.selfplay(2560)\n"
        "r = eng.records(); r = r[np.lexsort((r['ply'], r['game']))]\n"
        "np.save(sys.argv[1], r); print(json.dumps(dict(sims=st['sims'], evals=st['nn_evals'], games=st['games_finished'])))\n") % ([ROOT, PKG],)
    outs = {}
    with tempfile.TemporaryDirectory() as tmp:
        for ov in ("0", "1"):
            fn = os.path.join(tmp, "r%s.npy" % ov)
            r = subprocess.run([sys.executable, "-c", code, fn], env=dict(os.environ, CK_OVERLAP=ov), capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stdout + r.stderr
            outs[ov] = (np.load(fn), json.loads(r.stdout.strip().splitlines()[-1]))
    assert outs["0"][1] == outs["1"][1] and outs["0"][1]["games"] == 2560
    assert outs["0"][0].tobytes() == outs["1"][0].tobytes()


def test_packed_records_equal_full_records(lib):
    """device-side packing (ck_records_pack_device / ck_records_fetch_packed: 40-byte headers + one word per child) loses
    nothing: unpacked on the host it is bit-identical to ck_records_fetch, also with unfinished games in the store, and
    equals the numpy twin of the pack kernel"""
    from ckb200 import records as R
    eng = lib.Engine(lib.make_cfg(n_slots=9, budget=40, training=True, terminate_cnt=0, evaluator="hash_salted", epsilon=0.25,
                                  alpha=1.0, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=3, max_plies=600))
    eng.begin(30)
    eng.run(1500)                                     # some games finished, some in flight, some not started
    full = eng.records()
    hdr, words = eng.records_packed()
    assert 0 < len(full) == len(hdr) and hdr.dtype.itemsize == 40
    assert R.unpack(hdr, words).tobytes() == full.tobytes()
    h2, w2 = R.pack(full)
    assert h2.tobytes() == hdr.tobytes() and w2.tobytes() == words.tobytes()
    assert hdr.nbytes + words.nbytes < full.nbytes / 4
    eng.run(0)                                        # to the end
    full = eng.records()
    assert R.unpack(*eng.records_packed()).tobytes() == full.tobytes() and eng.games_finished() == 30
    eng.close()


def test_set_budget_between_runs(lib):
    """ck_engine_set_budget: searches that start after the call use the new BUDGET (bench.py's pre-roll); every record's
    visit counts tell which budget its search ran with"""
    eng = lib.Engine(lib.make_cfg(n_slots=4, budget=50, training=True, terminate_cnt=30, evaluator="hash_salted"))
    eng.begin(4)
    eng.set_budget(5)
    eng.run(40)
    eng.set_budget(50)
    eng.run(0)
    recs = eng.records()
    searched = recs[recs["n_children"] > 0]
    assert len(searched) == 4 * 30
    small, large = searched[searched["ply"] < 3], searched[searched["ply"] > 25]
    assert (small["visits"].sum(axis=1) <= 5 + 5).all() and (large["visits"].sum(axis=1) >= 50).all()
    assert (searched["root_n"] >= 5).all()
    eng.close()


def test_uct_first_search_golden_and_midgame_vs_oracle(lib):
    """NEURAL_NET=False tree policy (MCTS.py:78-89,113-115): one child per visit, UCT in float64, one playout
    per simulation.  With hashed playouts the whole search is deterministic: the reference's own first search
    (uct_kat.json) and an oracle search from a mid-game position with tree reuse."""
    fs = json.load(open(os.path.join(GOLDEN, "uct_kat.json")))["first_search"]
    eng = lib.Engine(lib.make_cfg(n_slots=1, budget=fs["budget"], evaluator="rollout_hash", keep_records=False))
    eng.tree_set_root(O.start_position())
    eng.tree_search(fs["budget"])
    n, w, b = eng.tree_root()
    assert n == fs["root_n"] and float(w) == fs["root_w"] and b == len(fs["children"])
    got = [(codec.meta_action(c["pos"][3]), c["n"], float(c["w"]), bool(c["terminal"])) for c in eng.tree_root_children()]
    assert got == [(c["action"], c["n"], c["w"], c["terminal"]) for c in fs["children"]]
    # partially expanded root (budget below the number of moves), then reuse after advancing
    pos = O.start_position()
    rng = np.random.RandomState(5)
    for _ in range(23):
        kids = O.movegen(pos)[0]
        pos = kids[rng.randint(len(kids))]
    t = O.Tree(pos, O.make_cfg(budget=3, rollout="hash"))
    eng.tree_set_root(pos)
    for sims in (3, 90, 200):
        t.search(sims)
        eng.tree_search(sims)
        ref = t.root_children()
        assert eng.tree_root()[:2] == t.root_stats() and eng.tree_root()[2] == len(ref)
        assert [(c["pos"], c["n"], float(c["w"])) for c in eng.tree_root_children()] == \
               [(c["pos"], c["n"], float(c["w"])) for c in ref]
    eng.close()


def test_wide_node_more_than_32_children_vs_oracle(lib):
    """a root with 39 legal moves (twelve kings): the second set of child slots of every lane in PUCT / UCT
    selection, expansion, move selection and the record path -- searches in both modes against the oracle"""
    pos = (185011974, 1073741824, 1258753798, codec.make_meta(0, 0, 0, 0, 90))
    assert len(O.movegen(pos)[0]) == 39
    for evaluator, cfg in (("hash", O.make_cfg(budget=400)), ("rollout_hash", O.make_cfg(budget=400, rollout="hash"))):
        t = O.Tree(pos, cfg, "hash")
        eng = lib.Engine(lib.make_cfg(n_slots=1, budget=400, evaluator=evaluator, keep_records=False))
        eng.tree_set_root(pos)
        for sims in (20, 400):                                   # 20: the playout mode has added 20 of the 39 children
            t.search(sims)
            eng.tree_search(sims)
            ref = t.root_children()
            assert eng.tree_root()[:2] == t.root_stats() and eng.tree_root()[2] == len(ref)
            got = eng.tree_root_children()
            assert [(c["pos"], c["n"], float(c["w"])) for c in got] == [(c["pos"], c["n"], float(c["w"])) for c in ref]
            if evaluator == "hash":                              # priors exist with a network only
                assert [float(c["p"]) for c in got] == [float(c["p"]) for c in ref]
        assert len(ref) == 39 and eng.tree_best_child() == t.best_child()
        eng.close()


def test_uct_selfplay_game_golden(lib):
    """_generate_data with NEURAL_NET=False run verbatim (hashed playouts): every search of a 40-ply game"""
    gk = json.load(open(os.path.join(GOLDEN, "uct_kat.json")))["game"]
    eng = lib.Engine(lib.make_cfg(n_slots=1, budget=gk["budget"], training=True, terminate_cnt=gk["terminate_cnt"],
                                  evaluator="rollout_hash"))
    st = eng.selfplay(1)
    recs, games = _engine_records(lib, eng)
    recs = recs[0]
    assert len(recs) == len(gk["moves"]) and games[0]["reroot_misses"] == 0
    assert st["sims"] == gk["budget"] * len(recs)
    for r, m, q, z in zip(recs, gk["moves"], gk["q"], gk["z"]):
        assert r["root_n"] == m["root_n"] and float(r["root_w"]) == m["root_w"]
        assert r["actions"] == [c["action"] for c in m["children"]]
        assert r["visits"] == [c["n"] for c in m["children"]]
        assert r["z"] == z
    from ckb200 import records as R
    for ref_rec, q in zip(R.to_reference_list(eng.records(), playouts=True), gk["q"]):
        assert type(ref_rec[2]) is type(q) or q == 0
        assert ref_rec[2] == q                                  # float64 int / int quotient, exact
    eng.close()


@pytest.mark.parametrize("slots,compact", [(16, False), (5, True)])
def test_uct_concurrent_games_vs_oracle(lib, slots, compact):
    """many playout-mode games in flight (hashed playouts salted per game so that the games differ), slots
    refilled, low budget so that roots are only partially expanded and replies go missing from the re-used
    tree: every record, the miss counts and the simulation totals equal the oracle's"""
    n_games, budget, term = 24, 20, 60
    ref = []
    for g in range(n_games):
        gm = O.Game(O.make_cfg(budget=budget, training=True, terminate_cnt=term, rollout="hash"), salt=g)
        gm.play()
        ref.append((gm.records(), gm.outcome, gm.move_count, gm.terminated, gm.total_sims, gm.reroot_misses))
        gm.close()
    eng = lib.Engine(lib.make_cfg(n_slots=slots, budget=budget, training=True, terminate_cnt=term,
                                  evaluator="rollout_hash", compact_always=compact, pool_cap=8192 if compact else 0))
    st = eng.selfplay(n_games)
    recs, games = _engine_records(lib, eng)
    assert len(games) == n_games
    for g in range(n_games):
        rr, outcome, move_count, terminated, sims, misses = ref[g]
        assert int(games[g]["outcome"]) == outcome and int(games[g]["move_count"]) == move_count
        assert bool(games[g]["terminated"]) == terminated and int(games[g]["reroot_misses"]) == misses
        assert int(games[g]["sims"]) == sims
        assert len(recs[g]) == len(rr)
        for a, b in zip(recs[g], rr):
            assert _same_record(a, b)
    assert sum(r[5] for r in ref) > 0                           # the fresh-root path was exercised
    assert st["games_finished"] == n_games
    eng.close()


def test_uct_random_playouts_properties(lib):
    """CK_EVAL_ROLLOUT (Philox playouts, no bit parity with numpy): exact simulation accounting, reproducible per
    seed, different across seeds, integer rewards"""
    def run(seed):
        eng = lib.Engine(lib.make_cfg(n_slots=32, budget=60, training=True, tau=1.0, tau_decay=0.1, tau_decay_delay=10,
                                      terminate_cnt=50, evaluator="rollout", seed=seed))
        st = eng.selfplay(48)
        recs, games = _engine_records(lib, eng)
        eng.close()
        return st, recs, games
    st, recs, games = run(7)
    assert st["games_finished"] == 48 and len(games) == 48
    for g, rr in recs.items():
        searches = [r for r in rr if r["actions"]]
        assert int(games[g]["sims"]) == 60 * len(searches)
        for r in searches:
            assert float(r["root_w"]) == int(r["root_w"]) and abs(float(r["root_w"])) <= r["root_n"]
            assert sum(r["visits"]) <= r["root_n"] and all(v >= 1 for v in r["visits"])
            assert r["chosen"] in r["actions"]
    st2, recs2, _ = run(7)
    assert all(_same_record(a, b) for g in recs for a, b in zip(recs[g], recs2[g]))
    _, recs3, _ = run(8)
    assert any(not _same_record(a, b) for g in recs for a, b in zip(recs[g], recs3[g]))
    with pytest.raises(lib.CkError):
        lib.Engine(lib.make_cfg(n_slots=2, budget=10, arena=True, evaluator="rollout"))


def test_long_game_draw_rule_vs_oracle(lib):
    """no ply cap: games run into the 80-ply draw window / long endgames (arena semantics)."""
    ref = _oracle_games(6, 24, 0, training=False)
    eng = lib.Engine(lib.make_cfg(n_slots=6, budget=24, training=False, terminate_cnt=0, evaluator="hash_salted",
                                  keep_records=True, max_plies=4096))
    eng.selfplay(6)
    recs, games = _engine_records(lib, eng)
    for g in range(6):
        rr, outcome, move_count, _t, sims, _e = ref[g]
        assert (int(games[g]["outcome"]), int(games[g]["move_count"]), int(games[g]["sims"])) == (outcome, move_count, sims)
        assert all(_same_record(a, b) for a, b in zip(recs[g], rr)) and len(recs[g]) == len(rr)
    eng.close()


def test_noise_and_temperature_invariants(lib):
    """epsilon > 0 / tau > 0 use the engine's own Philox streams (no bit parity with numpy's RNG by
    design); check the structural invariants of the records instead."""
    eng = lib.Engine(lib.make_cfg(n_slots=16, budget=40, training=True, terminate_cnt=60, evaluator="hash_salted",
                                  epsilon=0.25, alpha=1.0, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=7))
    st = eng.selfplay(32)
    recs, games = _engine_records(lib, eng)
    assert st["games_finished"] == 32
    chosen_not_argmax = 0
    for g, rr in recs.items():
        outcome = int(games[g]["outcome"])
        for r in rr:
            if r["chosen"] < 0:
                continue
            assert sum(r["visits"]) == r["root_n"] - 1          # N(node) = 1 + sum N(children)
            assert r["chosen"] in r["actions"] and -1.0 <= float(r["q"]) <= 1.0
            player = r["pos"][3] & 1
            want = 0 if outcome == 3 else (1 if (outcome == 1) == (player == 0) else -1)
            assert r["z"] == want
            if r["visits"][r["actions"].index(r["chosen"])] != max(r["visits"]):
                chosen_not_argmax += 1
    assert chosen_not_argmax > 0                                 # temperature sampling happened
    # same seed -> same games; other seed -> different
    eng2 = lib.Engine(lib.make_cfg(n_slots=5, budget=40, training=True, terminate_cnt=60, evaluator="hash_salted",
                                   epsilon=0.25, alpha=1.0, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=7))
    eng2.selfplay(32)
    recs2, _ = _engine_records(lib, eng2)
    assert all(len(recs[g]) == len(recs2[g]) and all(_same_record(a, b) for a, b in zip(recs[g], recs2[g])) for g in recs)
    eng.close()
    eng2.close()


# ---- K4 -------------------------------------------------------------------------------------
def test_rollout_properties(lib):
    pos = _pos_arr(lib, [O.start_position()] * 4096)
    outcome, plies = lib.rollout(pos, seed=3)
    assert set(np.unique(outcome)) <= {1, 2, 3} and (plies > 10).all()
    o2, p2 = lib.rollout(pos, seed=3)
    assert (o2 == outcome).all() and (p2 == plies).all()
    ref = [O.random_playout(O.start_position(), 1000 + i) for i in range(600)]
    ref_plies = np.mean([r[1] for r in ref])
    assert abs(plies.mean() - ref_plies) < 0.1 * ref_plies      # same rules => same playout-length statistics
    capped, cp = lib.rollout(pos[:64], seed=3, max_plies=5)
    assert (cp <= 5).all()


# ---- K3 -------------------------------------------------------------------------------------
def _random_leaves(lib, n, seed):
    rng = np.random.RandomState(seed)
    leaves = np.zeros(n, dtype=lib.LEAF_DTYPE)
    planes = np.zeros((n, 8, 8, 14), dtype=np.float32)
    i = 0
    while i < n:
        pos = O.start_position()
        for _ply in range(rng.randint(1, 140)):
            kids, mask, st, p5 = O.movegen(pos)
            if st != codec.ONGOING:
                break
            pos = kids[rng.randint(len(kids))]
        kids, mask, st, p5 = O.movegen(pos)
        if rng.rand() < 0.3:
            p5 = rng.randint(0, 81)
        leaves[i] = (pos[0], pos[1], pos[2], (pos[3] & 1) | (p5 << 8), mask)
        planes[i] = codec.nn_input_planes(pos, mask, p5)
        i += 1
    return leaves, planes


@pytest.mark.parametrize("impl", ["simt", "tc"])
def test_net_forward_vs_torch(lib, impl):
    from ckb200 import net as N
    from oracle import net_oracle as NO
    for seed, jitter in ((0, 0.0), (1, 0.2)):
        blob = N.random_init_blob(seed, jitter)
        net = lib.Net(0, impl)
        net.set_weights(blob)
        leaves, planes = _random_leaves(lib, 97, seed)
        pol, val = net.forward(leaves)
        rp, rv = NO.forward(N.unpack(blob), planes)
        assert np.abs(pol - rp).max() < NET_TOL and np.abs(val - rv).max() < NET_TOL
        assert np.abs(pol.sum(1) - 1).max() < 1e-5
        kp, kv = net.predict(planes)                               # Keras-signature entry point
        assert kp.tobytes() == pol.tobytes() and kv.reshape(-1).tobytes() == val.tobytes()
        net.close()


def test_selfplay_with_network(lib):
    """whole path with the real network evaluator: games finish, records are consistent."""
    from ckb200 import net as N
    net = lib.Net(0)
    net.set_weights(N.random_init_blob(0))
    eng = lib.Engine(lib.make_cfg(n_slots=32, budget=30, training=True, terminate_cnt=40, evaluator="net",
                                  epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10))
    eng.set_net(0, net)
    st = eng.selfplay(32)
    recs, games = _engine_records(lib, eng)
    assert st["games_finished"] == 32 and st["nn_evals"] > 0 and st["sims"] >= 32 * 30
    for g, rr in recs.items():
        for r in rr:
            if r["chosen"] >= 0:
                assert sum(r["visits"]) == r["root_n"] - 1
    eng.close()
    net.close()


@pytest.mark.parametrize("tower", ["ts", "ts-one-tile", "ts-two-tiles", "ss"])
def test_tower_kernels_agree_with_fp32_path(tower):
    """both tcgen05 towers (weights-in-TMEM product kernel, shared-memory cross-check) against the fp32
    CUDA-core path on ragged batch sizes: single tile pair, odd tails, the persistent multi-pair loop.
    The tower is chosen once per process (CK_TOWER), hence the subprocess."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CK_TOWER=tower[:2])
    if tower == "ts-one-tile":
        env["CK_TS_TILES"] = "1"          # small-batch mode forced on every size (the default picks it up to 296)
    elif tower == "ts-two-tiles":
        env["CK_TS_TILES"] = "2"
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "check_tower.py"), "1", "2", "3", "5", "97", "593", "1187", "4099"],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count(" ok ") == 8 and "MISMATCH" not in r.stdout


def test_full_size_properties_and_batch_independence(lib):
    """cfg2 size (4096 concurrent games, 400 sims/move) with the network evaluator, noise and temperature on:
    (i) structural invariants of every record, (ii) the run is reproducible bit for bit although leaves
    get their batch rows in a different (atomic) order every time -- i.e. a position's evaluation does not
    depend on its row, its tile mates or the CTA that processed it, (iii) a game's records do not depend on
    how many games run next to it (the same game ids in a 592-slot and in a 128-slot engine)."""
    from ckb200 import net as N
    net = lib.Net(0)
    net.set_weights(N.random_init_blob(0))

    def run(slots):
        eng = lib.Engine(lib.make_cfg(n_slots=slots, budget=400, training=True, terminate_cnt=3, evaluator="net", keep_records=True,
                                      uct_c=4.0, alpha=1.0, epsilon=0.25, tau=1.0, tau_decay=0.1, tau_decay_delay=10, seed=99))
        eng.set_net(0, net)
        st = eng.selfplay(slots)                     # every game: three searched moves, then the ply-cap adjudication
        recs = eng.records()
        eng.close()
        return st, recs[np.lexsort((recs["ply"], recs["game"]))]

    st, a = run(4096)
    assert st["games_finished"] == 4096 and len(a) == 3 * 4096 and st["moves"] == 3 * 4096
    assert st["nn_evals"] <= st["sims"] and st["sims"] >= 3 * 400 * 4096
    n = a["n_children"].astype(np.int64)
    vis = a["visits"].astype(np.int64)
    col = np.arange(vis.shape[1])[None, :]
    assert (np.where(col < n[:, None], vis, 0).sum(1) == a["root_n"].astype(np.int64) - 1).all()      # :433-434
    assert (a["root_n"] >= 400).all() and (np.abs(a["q"]) <= 1).all() and (np.abs(a["z"]) <= 1).all()
    st2, b = run(4096)
    assert st2["sims"] == st["sims"] and a.tobytes() == b.tobytes()
    _st3, c = run(592)
    assert len(c) == 3 * 592 and a[a["game"] < 592].tobytes() == c.tobytes()
    _st4, d = run(128)                              # small batch: the tower runs one tile per CTA
    assert len(d) == 3 * 128 and a[a["game"] < 128].tobytes() == d.tobytes()
    net.close()
