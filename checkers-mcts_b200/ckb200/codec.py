"""State codec: reference 15x8x8 float64 state  <->  compact bitboard position.

The reference's state layout is documented at Checkers.py:37-48: planes 0-3 pieces,
plane 4 side to move, plane 5 draw counter (n/80), planes 6-13 legal-action masks,
plane 14 holds the action triple (plane, x, y) of the move that produced the state
in ``state[14,0,0:3]`` (Checkers.py:138-143).

Compact position (``POS_DTYPE``, 16 bytes, mirrors ``ck_pos`` in include/ckb200.h):
``p1``/``p2`` = all pieces of each side, ``k`` = kings of either side, over the 32
playable squares ``s = 4*x + (y>>1)``; ``meta`` packs player / reversible-ply counter /
action id / ply index.
"""
import numpy as np

POS_DTYPE = np.dtype([("p1", "<u4"), ("p2", "<u4"), ("k", "<u4"), ("meta", "<u4")])

ONGOING, P1_WINS, P2_WINS, DRAW = 0, 1, 2, 3
OUTCOME_NAMES = {ONGOING: None, P1_WINS: "player1_wins", P2_WINS: "player2_wins", DRAW: "draw"}
OUTCOME_CODES = {v: k for k, v in OUTCOME_NAMES.items()}
MAX_CHILDREN = 48

# square tables
_SQ_X = np.array([s // 4 for s in range(32)])
_SQ_Y = np.array([2 * (s % 4) + (1 if (s // 4) % 2 == 0 else 0) for s in range(32)])
_BITS = (np.uint32(1) << np.arange(32, dtype=np.uint32))


def sq_index(x, y):
    return 4 * x + (y >> 1)


def make_meta(player, rev=0, action=0, has_action=0, ply=0):
    rev = min(int(rev), 127)
    ply = min(int(ply), 0x3FFF)
    return int(player) | (rev << 1) | (int(action) << 8) | (int(has_action) << 17) | (ply << 18)


def meta_player(m):
    return int(m) & 1


def meta_rev(m):
    return (int(m) >> 1) & 0x7F


def meta_action(m):
    return (int(m) >> 8) & 0x1FF


def meta_has_action(m):
    return (int(m) >> 17) & 1


def meta_ply(m):
    return (int(m) >> 18) & 0x3FFF


def action_id(plane, x, y):
    return (int(plane) - 6) * 64 + int(x) * 8 + int(y)


def action_triple(a):
    return (a >> 6) + 6, (a >> 3) & 7, a & 7


def plane_to_bits(plane):
    """8x8 0/1 plane -> 32-bit square set."""
    v = plane[_SQ_X, _SQ_Y] != 0
    return int(np.bitwise_or.reduce(_BITS[v])) if v.any() else 0


def bits_to_plane(bits, out=None):
    if out is None:
        out = np.zeros((8, 8), dtype=np.float64)
    sel = (np.uint32(bits) & _BITS) != 0
    out[_SQ_X[sel], _SQ_Y[sel]] = 1.0
    return out


def encode_state(state, rev=0, ply=0):
    """15x8x8 reference state -> compact tuple (p1, p2, k, meta).

    ``rev`` / ``ply`` are history-derived and must be supplied by the
    caller; the action triple is read from plane 14."""
    m1, k1, m2, k2 = (plane_to_bits(state[i]) for i in range(4))
    player = int(state[4, 0, 0])
    plane = int(state[14, 0, 0])
    has_action = 1 if 6 <= plane <= 13 else 0
    act = action_id(plane, state[14, 0, 1], state[14, 0, 2]) if has_action else 0
    return (m1 | k1, m2 | k2, k1 | k2, make_meta(player, rev, act, has_action, ply))


def decode_state(pos, mask=None, plane5=0):
    """compact position (+ mask planes 6..13 and plane-5 numerator) -> 15x8x8 float64."""
    p1, p2, k, meta = (int(v) for v in pos)
    st = np.zeros((15, 8, 8), dtype=np.float64)
    bits_to_plane(p1 & ~k, st[0])
    bits_to_plane(p1 & k, st[1])
    bits_to_plane(p2 & ~k, st[2])
    bits_to_plane(p2 & k, st[3])
    st[4] = float(meta_player(meta))
    if plane5:
        st[5] = plane5 / 80
    if mask is not None:
        for i in range(8):
            bits_to_plane(int(mask[i]), st[6 + i])
    if meta_has_action(meta):
        st[14, 0, 0], st[14, 0, 1], st[14, 0, 2] = action_triple(meta_action(meta))
    return st


def pos_array(positions):
    """list of (p1,p2,k,meta) -> structured array."""
    a = np.zeros(len(positions), dtype=POS_DTYPE)
    for i, p in enumerate(positions):
        a[i] = tuple(int(v) for v in p)
    return a


def start_position():
    p1 = 0
    p2 = 0
    for s in range(32):
        if _SQ_X[s] < 3:
            p1 |= 1 << s
        elif _SQ_X[s] > 4:
            p2 |= 1 << s
    return (p1, p2, 0, make_meta(0))


def nn_input_planes(pos, mask, plane5):
    """float32 [8,8,14] channels-last NN input (Checkers.py:431-432)."""
    st = decode_state(pos, mask, plane5)
    return np.moveaxis(st[:14], 0, -1).astype(np.float32)
