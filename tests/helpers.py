"""Shared helpers for the test-suite (stub networks that exist on both sides)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "checkers-mcts_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from ckb200 import codec  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _mix32(h):
    h &= 0xFFFFFFFF
    h ^= h >> 16
    h = (h * 0x7FEB352D) & 0xFFFFFFFF
    h ^= h >> 15
    h = (h * 0x846CA68B) & 0xFFFFFFFF
    h ^= h >> 16
    return h


def hash_eval(pos, plane5):
    """Python twin of cko_eval_hash (oracle/ck_oracle.c) and the CUDA stub evaluator."""
    p1, p2, k, meta = (int(v) for v in pos)
    h = _mix32(p1 ^ 0x9E3779B9)
    h = _mix32(h ^ p2)
    h = _mix32(h ^ k)
    h = _mix32(h ^ (meta & 1) ^ ((plane5 << 8) & 0xFFFFFFFF))
    i = np.arange(512, dtype=np.uint64)
    g = (np.uint64(h) + i * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    g ^= g >> np.uint64(16)
    g = (g * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    g ^= g >> np.uint64(15)
    g = (g * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    g ^= g >> np.uint64(16)
    v = (((g >> np.uint64(8)) & np.uint64(0xFFFF)).astype(np.float32) + np.float32(1.0))
    pol = ((v * v) * v) * np.float32(2.0 ** -58)
    gv = _mix32(h ^ 0xDEADBEEF)
    val = np.float32(gv & 0xFFFFFF) * np.float32(2.0 ** -23) - np.float32(1.0)
    return pol.astype(np.float32), np.float32(val)


class KerasLikeStub(object):
    """Object with Keras' ``predict`` signature (Checkers.py:433) backed by a stub evaluator.

    kind: 'uniform_zero' (KAT-A), 'uniform_material' (KAT-B), 'hash'."""

    def __init__(self, kind):
        self.kind = kind
        self.calls = 0

    def predict(self, x):
        self.calls += 1
        x = np.asarray(x)[0]
        if self.kind in ("uniform_zero", "uniform_material"):
            pol = np.full((1, 512), 1 / 512, dtype=np.float32)
            if self.kind == "uniform_zero":
                v = np.float32(0)
            else:
                m = [x[..., i].sum() for i in range(4)]
                own, opp = (m[0] + 2 * m[1], m[2] + 2 * m[3])
                if x[0, 0, 4] == 1:
                    own, opp = opp, own
                v = np.float32((own - opp) / 32)
            return [pol, np.array([[v]], dtype=np.float32)]
        b = [codec.plane_to_bits(x[..., i]) for i in range(4)]
        player = int(x[0, 0, 4])
        plane5 = int(round(float(x[0, 0, 5]) * 80))
        pos = (b[0] | b[1], b[2] | b[3], b[1] | b[3], player)
        pol, v = hash_eval(pos, plane5)
        return [pol.reshape(1, 512), np.array([[v]], dtype=np.float32)]


def record_from_reference(entry):
    """reference record [state, probs, q, z] -> comparable dict."""
    state, probs, q, z = entry
    return dict(state=np.asarray(state, dtype=np.float64), probs=np.asarray(probs, dtype=np.float64),
                q=float(q), z=int(z))


def record_planes(rec):
    """oracle/engine record dict -> (state[15,8,8] f64, probs[8,8,8] f64) as the reference stores."""
    state = codec.decode_state(rec["pos"], rec["mask"], rec["plane5"])
    probs = np.zeros(512, dtype=np.float64)
    if rec["actions"]:
        v = np.asarray(rec["visits"], dtype=np.float64)
        probs[np.asarray(rec["actions"], dtype=np.int64)] = v
        probs = probs.reshape(8, 8, 8)
        probs /= np.sum(probs)
    return state, probs.reshape(8, 8, 8)


class hashed_playouts(object):
    """np.random.randint(0, n) inside MCTS.default_policy (MCTS.py:141) -> oracle.hash_choice(position, n): the
    reference code runs unmodified, only numpy's generator is replaced, and the replacement looks at the
    playout environment of the calling frame (``game_sim``) to hash the position the move is chosen from."""

    def __enter__(self):
        from oracle import oracle as O
        self.orig = np.random.randint

        def fake_randint(low, high=None, *a, **k):
            state = sys._getframe(1).f_locals['game_sim'].state
            bits = [codec.plane_to_bits(state[i]) for i in range(4)]
            pos = (bits[0] | bits[1], bits[2] | bits[3], bits[1] | bits[3], int(state[4, 0, 0]))
            return O.hash_choice(pos, high)

        np.random.randint = fake_randint
        return self

    def __exit__(self, *exc):
        np.random.randint = self.orig
