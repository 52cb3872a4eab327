"""Pins the CPU oracle (oracle/ck_oracle.c) against the golden vectors generated from the
unmodified reference by tests/golden/make_golden.py.  CPU only."""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN, codec, record_planes
from oracle import oracle as O


def test_start_position_and_codec_roundtrip():
    sp = O.start_position()
    assert sp == codec.start_position()
    st = codec.decode_state(sp)
    assert st[0].sum() == 12 and st[2].sum() == 12 and st[4].sum() == 0
    assert codec.encode_state(st) == sp


def test_movegen_golden():
    g = np.load(os.path.join(GOLDEN, "movegen_cases.npz"))
    off = np.concatenate([[0], np.cumsum(g["counts"])])
    assert len(g["pos"]) > 3000
    for i in range(len(g["pos"])):
        kids, mask, status, p5 = O.movegen(g["pos"][i])
        ref = g["kids"][off[i]:off[i + 1]]
        assert len(kids) == len(ref)
        for k, r in zip(kids, ref):
            assert tuple(k[:3]) == tuple(int(v) for v in r[:3])
            assert codec.meta_player(k[3]) == r[3] and codec.meta_action(k[3]) == r[4]
        assert mask == [int(v) for v in g["mask"][i]]
        assert status == g["status"][i]
        assert p5 == g["plane5"][i]


def test_perft_golden():
    ref = json.load(open(os.path.join(GOLDEN, "perft.json")))
    sp = O.start_position()
    for d in range(1, 8):
        assert list(O.perft(sp, d)) == ref[str(d)]
    assert list(O.perft(sp, 8)) == ref["8_survey_reported"]


def test_predict_glue_golden():
    g = np.load(os.path.join(GOLDEN, "predict_glue.npz"))
    for pol, mask, prior in zip(g["policy"], g["mask"], g["prior"]):
        out = O.mask_renorm(pol, mask)
        assert out.tobytes() == prior.tobytes()       # bit-exact float32


@pytest.mark.parametrize("kind", ["uniform_zero", "uniform_material", "hash"])
def test_first_search_golden(kind):
    kat = json.load(open(os.path.join(GOLDEN, "mcts_kat.json")))[kind + "_first_search"]
    t = O.Tree(O.start_position(), O.make_cfg(budget=kat["budget"]), kind)
    t.search(kat["budget"])
    n, w = t.root_stats()
    assert n == kat["root_n"] and float(w) == kat["root_w"]
    got = t.root_children()
    assert len(got) == len(kat["children"])
    for c, r in zip(got, kat["children"]):
        assert codec.meta_action(c["pos"][3]) == r["action"]
        assert c["n"] == r["n"] and float(c["w"]) == r["w"] and float(c["p"]) == r["p"]


@pytest.mark.parametrize("kind", ["uniform_zero", "uniform_material", "hash"])
def test_game_opening_golden(kind):
    kat = json.load(open(os.path.join(GOLDEN, "mcts_kat.json")))[kind + "_game"]
    g = O.Game(O.make_cfg(budget=kat["budget"], training=True, terminate_cnt=kat["terminate_cnt"]), kind)
    g.play()
    recs = g.records()
    assert len(recs) == len(kat["moves"])
    for r, m in zip(recs, kat["moves"]):
        assert r["root_n"] == m["root_n"] and float(r["root_w"]) == m["root_w"]
        assert r["actions"] == [c["action"] for c in m["children"]]
        assert r["visits"] == [c["n"] for c in m["children"]]


def test_uct_rollout_mode_golden():
    """NEURAL_NET=False (MCTS.py:78-89,113-115,132-146): one child per visit, UCT selection, one playout per
    simulation -- first search and a 40-ply _generate_data game of the reference with hashed playouts."""
    kat = json.load(open(os.path.join(GOLDEN, "uct_kat.json")))
    fs = kat["first_search"]
    t = O.Tree(O.start_position(), O.make_cfg(budget=fs["budget"], rollout="hash"))
    t.search(fs["budget"])
    n, w = t.root_stats()
    assert n == fs["root_n"] and float(w) == fs["root_w"]
    got = [(codec.meta_action(c["pos"][3]), c["n"], float(c["w"]), bool(c["terminal"])) for c in t.root_children()]
    assert got == [(c["action"], c["n"], c["w"], c["terminal"]) for c in fs["children"]]
    gk = kat["game"]
    g = O.Game(O.make_cfg(budget=gk["budget"], training=True, terminate_cnt=gk["terminate_cnt"], rollout="hash"))
    g.play()
    recs = g.records()
    assert len(recs) == len(gk["moves"]) and g.reroot_misses == 0
    for r, m, q, z in zip(recs, gk["moves"], gk["q"], gk["z"]):
        assert r["root_n"] == m["root_n"] and float(r["root_w"]) == m["root_w"]
        assert r["actions"] == [c["action"] for c in m["children"]]
        assert r["visits"] == [c["n"] for c in m["children"]]
        # the record carries q as float32; the reference's int / int quotient is recovered exactly from root_n / root_w
        assert float(r["q"]) == pytest.approx(q, abs=0, rel=2e-7) and r["z"] == z


@pytest.mark.parametrize("kind", ["hash", "uniform_material"])
def test_selfplay_records_golden(kind):
    f = np.load(os.path.join(GOLDEN, "selfplay_%s.npz" % kind))
    meta = json.loads(str(f["meta"]))
    g = O.Game(O.make_cfg(budget=meta["budget"], training=True, terminate_cnt=meta["terminate_cnt"]), kind)
    g.play()
    recs = g.records()
    assert len(recs) == len(f["q"])
    assert g.reroot_misses == 0
    for i, r in enumerate(recs):
        state, probs = record_planes(r)
        pl = [codec.plane_to_bits(state[j]) for j in (0, 1, 2, 3, 6, 7, 8, 9, 10, 11, 12, 13)]
        assert pl == [int(v) for v in f["planes"][i]]
        assert int(state[4, 0, 0]) == f["player"][i]
        assert r["plane5"] == f["plane5"][i]
        assert [int(v) for v in state[14, 0, 0:3]] == [int(v) for v in f["action"][i]]
        assert probs.reshape(512).tobytes() == f["probs"][i].tobytes()
        assert float(r["q"]) == f["q"][i] and r["z"] == f["z"][i]
    n_moves = len(f["root_n"])
    assert [r["root_n"] for r in recs[:n_moves]] == list(f["root_n"])
    assert [float(r["root_w"]) for r in recs[:n_moves]] == list(f["root_w"])


def test_tournament_golden():
    t = json.load(open(os.path.join(GOLDEN, "tournament.json")))
    names = {0: None, 1: "player1_wins", 2: "player2_wins", 3: "draw"}
    for game_num, p1, p2, outcome, move_count in t["outcomes"]:
        g = O.Game(O.make_cfg(budget=t["budget"], training=False), t["nets"]["data/model/" + p1],
                   t["nets"]["data/model/" + p2])
        g.play()
        assert names[g.outcome] == outcome and g.move_count == move_count


def test_pow_half_is_libm_pow():
    """node.n ** 0.5 (MCTS.py:110) is libm pow, which differs from sqrt for some integers;
    the oracle uses pow and the engine ships a host-built pow table."""
    import math
    assert sum(1 for n in range(1, 70000) if n ** 0.5 != math.sqrt(n)) > 0
