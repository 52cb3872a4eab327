"""CPU check of the product's thread-level rule header (csrc/ck_rules.cuh) compiled for the
host, against the oracle.  This is a logic check for a box without a GPU; the GPU parity tests
proper are in test_gpu_*.py and go through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import GOLDEN, ROOT, codec
from oracle import oracle as O

_SRC = os.path.join(ROOT, "tests", "host_rules", "host_rules.cpp")
_OUT = os.path.join(ROOT, "tests", "host_rules", "_build", "libhostrules.so")


@pytest.fixture(scope="module")
def hr():
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    hdr = os.path.join(ROOT, "checkers-mcts_b200", "csrc", "ck_rules.cuh")
    if not os.path.exists(_OUT) or os.path.getmtime(_OUT) < max(os.path.getmtime(_SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", _OUT, _SRC])
    L = C.CDLL(_OUT)
    L.ckh_movegen.argtypes = [C.POINTER(O.Pos), C.POINTER(O.Pos), C.POINTER(C.c_uint32),
                              C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ckh_movegen.restype = C.c_int
    L.ckh_movegen_fast.argtypes = [C.POINTER(O.Pos), C.POINTER(O.Pos)]
    L.ckh_movegen_fast.restype = C.c_int
    L.ckh_kth.argtypes = [C.POINTER(O.Pos), C.c_int, C.POINTER(O.Pos)]
    L.ckh_kth.restype = None
    L.ckh_status.argtypes = [C.POINTER(O.Pos), C.POINTER(C.c_int)]
    L.ckh_status.restype = C.c_int
    L.ckh_hash_playout.argtypes = [C.POINTER(O.Pos), C.c_uint32, C.POINTER(C.c_int)]
    L.ckh_hash_playout.restype = C.c_int
    return L


def _host_movegen(L, pos):
    ch = (O.Pos * O.MAX_CHILDREN)()
    mask = (C.c_uint32 * 8)()
    st, p5 = C.c_int(), C.c_int()
    p = O.Pos(*[int(v) for v in pos])
    n = L.ckh_movegen(C.byref(p), ch, mask, C.byref(st), C.byref(p5))
    one = O.Pos()
    for k in range(n):                                   # random access must agree with the generated list
        L.ckh_kth(C.byref(p), k, C.byref(one))
        assert one.tup() == ch[k].tup(), (pos, k)
    fast = (O.Pos * O.MAX_CHILDREN)()                     # the packed kernel's lean successor construction
    assert L.ckh_movegen_fast(C.byref(p), fast) == n
    assert [fast[i].tup() for i in range(n)] == [ch[i].tup() for i in range(n)], pos
    p5b = C.c_int()
    st2 = L.ckh_status(C.byref(p), C.byref(p5b))
    assert (st2, p5b.value) == (st.value, p5.value)
    return [ch[i].tup() for i in range(n)], list(mask), st.value, p5.value


def test_host_rules_golden(hr):
    g = np.load(os.path.join(GOLDEN, "movegen_cases.npz"))
    for i in range(len(g["pos"])):
        assert _host_movegen(hr, g["pos"][i]) == O.movegen(g["pos"][i])


def test_host_rules_random_walks(hr):
    rng = np.random.RandomState(7)
    n = 0
    for _ in range(60):
        pos = O.start_position()
        for _ply in range(400):
            a = _host_movegen(hr, pos)
            assert a == O.movegen(pos)
            n += 1
            kids, _, st, _ = a
            if st != codec.ONGOING:
                break
            pos = kids[rng.randint(len(kids))]
    assert n > 3000


def test_host_rules_synthetic(hr):
    """dense random boards (many kings, both sides to move, late plies for the draw rule)."""
    rng = np.random.RandomState(11)
    for _ in range(4000):
        n = rng.randint(2, 25)
        sq = rng.permutation(32)[:n]
        p1 = p2 = k = 0
        for j, s in enumerate(sq):
            side = j % 2 if rng.rand() < 0.8 else rng.randint(2)
            king = rng.rand() < 0.4
            x = s // 4
            if not king and ((side == 0 and x == 7) or (side == 1 and x == 0)):
                king = True
            if side == 0:
                p1 |= 1 << int(s)
            else:
                p2 |= 1 << int(s)
            if king:
                k |= 1 << int(s)
        meta = codec.make_meta(rng.randint(2), rng.randint(0, 90), 0, 0, rng.randint(0, 200))
        pos = (p1, p2, k, meta)
        assert _host_movegen(hr, pos) == O.movegen(pos)


def test_host_playouts_match_oracle(hr):
    """ck::play_out (the loop of K4 `rollout_kernel` and of `playout_eval_kernel`: one generation pass per ply,
    successor built from the legal-action planes) with the hashed choice against the oracle's playout, from
    positions all over random games -- outcome and length, including the draw rule's counters"""
    rng = np.random.RandomState(11)
    seen = set()
    for game in range(60):
        pos = O.start_position()
        for _ply in range(rng.randint(0, 120)):
            kids, _mask, status, _p5 = O.movegen(pos)
            if status != 0:
                break
            pos = kids[rng.randint(len(kids))]
        salt = game % 5
        plies = C.c_int()
        got = hr.ckh_hash_playout(C.byref(O.Pos(*[int(v) for v in pos])), salt, C.byref(plies))
        want = O.hash_playout(pos, salt)
        assert (got, plies.value) == want, (pos, salt)
        seen.add(got)
    assert seen >= {1, 2}                                  # both colours win somewhere in the sample
