"""Conversion of engine records (ck_record, include/ckb200.h) to the reference's self-play record
format: ``[state float64[15,8,8], probs float64[8,8,8], q, z]`` (reference
training_pipeline.py:364-369, 406-411, 439-455), and the ``Keras_Generator`` batch layout
(:288-307).

Everything here works on whole record arrays (numpy bit-unpacking over all records at once): at
cfg2 a self-play iteration produces ~600 k records, and the per-record Python loop this replaces
took as long as the GPU needed to play the games.  ``save_reference_pickles`` fans the per-worker
files of ``generate_Checkers_data`` (:457-469) out over a process pool."""
import os
import pickle

import numpy as np

_SQ = np.arange(32)
_SQ_X = _SQ // 4
_SQ_Y = 2 * (_SQ % 4) + (1 - (_SQ_X % 2))
_FLAT = _SQ_X * 8 + _SQ_Y                      # index of playable square s in a flattened 8x8 plane


def _planes(bits, out=None):
    """uint32 square sets [...] -> float64 planes [..., 64] (flattened 8x8).  Square s = 4x + c sits at column 2c + 1 in
    even rows and 2c in odd rows, so the 32 unpacked bits land in the plane with two strided assignments."""
    bits = np.ascontiguousarray(bits, dtype="<u4")
    b = np.unpackbits(bits.view(np.uint8).reshape(bits.shape + (4,)), axis=-1, bitorder="little")      # [..., 32]
    if out is None:
        out = np.zeros(bits.shape + (64,), dtype=np.float64)
    o5 = out.reshape(bits.shape + (4, 2, 4, 2))          # (row pair, row parity, c, column parity)
    b4 = b.reshape(bits.shape + (4, 2, 4))
    o5[..., :, 0, :, 1] = b4[..., :, 0, :]
    o5[..., :, 1, :, 0] = b4[..., :, 1, :]
    return out


def states_of(records):
    """RECORD_DTYPE array [n] -> reference states float64 [n,15,8,8] (Checkers.py:37-48): planes 0-3 pieces, 4 side
    to move, 5 draw counter n/80, 6-13 legal-action planes, plane 14 the action triple at [0, 0:3]"""
    n = len(records)
    pos = records["pos"]
    if pos.dtype.names:
        p1, p2, k, meta = pos["p1"], pos["p2"], pos["k"], pos["meta"]
    else:                                                # plain uint32[4] columns
        p1, p2, k, meta = (np.ascontiguousarray(pos[:, i]) for i in range(4))
    st = np.zeros((n, 15, 64), dtype=np.float64)
    sets = np.empty((n, 12), dtype=np.uint32)            # planes 0-3 and 6-13 as square sets
    sets[:, 0], sets[:, 1], sets[:, 2], sets[:, 3] = p1 & ~k, p1 & k, p2 & ~k, p2 & k
    sets[:, 4:] = records["mask"]
    pl = _planes(sets)
    st[:, 0:4] = pl[:, 0:4]
    st[:, 6:14] = pl[:, 4:12]
    st[:, 4] = (meta & 1).astype(np.float64)[:, None]
    st[:, 5] = (records["plane5"].astype(np.float64) / 80)[:, None]      # n / 80 as the reference divides (Checkers.py:346-361)
    has = ((meta >> 17) & 1).astype(bool)
    act = ((meta >> 8) & 0x1FF).astype(np.int64)
    st[has, 14, 0] = (act[has] >> 6) + 6
    st[has, 14, 1] = (act[has] >> 3) & 7
    st[has, 14, 2] = act[has] & 7
    return st.reshape(n, 15, 8, 8)


def probs_of(records):
    """visit-count planes / their sum (_create_prob_planes, :421-437) float64 [n,8,8,8]; zero for terminal records"""
    n = len(records)
    nch = records["n_children"].astype(np.int64)
    cols = np.arange(records["visits"].shape[1])[None, :]
    live = cols < nch[:, None]
    probs = np.zeros((n, 512), dtype=np.float64)
    rows = np.broadcast_to(np.arange(n)[:, None], live.shape)[live]
    probs[rows, records["action"].astype(np.int64)[live]] = records["visits"].astype(np.float64)[live]
    tot = probs.sum(axis=1)                                              # integers: exact in any summation order
    nz = tot > 0
    probs[nz] /= tot[nz, None]
    return probs.reshape(n, 8, 8, 8)


def q_of(records, playouts=False):
    """list of the q entries with the reference's types: np.float32 for searched moves, the Python ints 0 / -1 for the
    terminal record (:407-408); with ``playouts`` (NEURAL_NET=False) rewards are Python ints and root.q their float64
    quotient (MCTS.py:389-394), rebuilt from the exact root_w / root_n and flipped to the root player's view (:365-368)"""
    nch = records["n_children"]
    q32 = records["q"]
    out = []
    if playouts:
        rw, rn = records["root_w"].astype(np.float64), records["root_n"].astype(np.int64)
        for i in range(len(records)):
            if not nch[i]:
                out.append(int(q32[i]))
                continue
            q = float(rw[i]) / int(rn[i]) if rn[i] else 0
            out.append(-q if q * float(q32[i]) < 0 else q)
        return out
    for i in range(len(records)):
        out.append(q32[i] if nch[i] else int(q32[i]))
    return out


def to_reference_list(records, playouts=False):
    """RECORD_DTYPE array -> list of [state, probs, q, z] exactly as the reference pickles it"""
    if len(records) == 0:
        return []
    states, probs, q = states_of(records), probs_of(records), q_of(records, playouts)
    z = records["z"].tolist()
    return [[states[i], probs[i], q[i], z[i]] for i in range(len(records))]


def to_reference(rec, playouts=False):
    """one RECORD_DTYPE element -> [state, probs, q, z]"""
    return to_reference_list(np.asarray(rec).reshape(1), playouts)[0]


def training_batch(records):
    """(states[:, :14] channels-last, [visit-probs(512), (q+z)/2]) as Keras_Generator.__getitem__ -- one pass"""
    x = np.moveaxis(states_of(records)[:, :14], 1, -1).astype(np.float32)
    probs = probs_of(records).reshape(len(records), 512)
    v = (records["q"].astype(np.float64) + records["z"].astype(np.float64)) / 2
    return x, [probs, v]


# ---- packed wire format (ck_record_hdr + child words, include/ckb200.h) ------------------------------------------------
def _mask_of_actions(rows, actions, n):
    """legal-action planes 6..13 as square sets [n,8] from (record row, action id) pairs: plane = a >> 6, square of
    (x, y) = ((a >> 3) & 7, a & 7) is s = 4x + (y >> 1)"""
    mask = np.zeros((n, 8), dtype=np.uint32)
    a = actions.astype(np.int64)
    bit = (np.uint32(1) << (4 * ((a >> 3) & 7) + ((a & 7) >> 1)).astype(np.uint32)).astype(np.uint32)
    np.bitwise_or.at(mask, (rows, a >> 6), bit)
    return mask


def pack(records):
    """RECORD_DTYPE array -> (RECORD_HDR_DTYPE array, uint32 words): numpy twin of the device-side pack kernel"""
    from .lib_types import RECORD_HDR_DTYPE
    n = len(records)
    hdr = np.zeros(n, dtype=RECORD_HDR_DTYPE)
    for f in ("pos", "q", "root_w", "root_n", "game", "ply", "chosen", "n_children", "plane5", "z"):
        hdr[f] = records[f]
    nch = records["n_children"].astype(np.int64)
    mf = (nch == 0) & (records["mask"] != 0).any(axis=1)
    hdr["flags"] = mf.astype(np.uint8)
    nw = nch + 8 * mf
    offs = np.concatenate([[0], np.cumsum(nw)])
    words = np.zeros(int(offs[-1]), dtype=np.uint32)
    cols = np.arange(records["visits"].shape[1])[None, :]
    live = cols < nch[:, None]
    dst = (offs[:-1, None] + cols)[live]
    words[dst] = (records["action"].astype(np.uint32)[live] << np.uint32(23)) | (records["visits"][live] & np.uint32(0x7FFFFF))
    for i in np.nonzero(mf)[0]:
        words[offs[i]:offs[i] + 8] = records["mask"][i]
    return hdr, words


def unpack(hdr, words):
    """(headers, child words) -> RECORD_DTYPE array, bit-identical to what ck_records_fetch returns"""
    from .lib_types import RECORD_DTYPE
    n = len(hdr)
    out = np.zeros(n, dtype=RECORD_DTYPE)
    for f in ("pos", "q", "root_w", "root_n", "game", "ply", "chosen", "n_children", "plane5", "z"):
        out[f] = hdr[f]
    nch = hdr["n_children"].astype(np.int64)
    mf = hdr["flags"].astype(bool) & (nch == 0)
    nw = nch + 8 * mf
    offs = np.concatenate([[0], np.cumsum(nw)])
    if int(offs[-1]) != len(words):
        raise ValueError("packed records: %d child words expected, %d present" % (int(offs[-1]), len(words)))
    cols = np.arange(out["visits"].shape[1])[None, :]
    live = cols < nch[:, None]
    w = words[(offs[:-1, None] + cols)[live]]
    rows = np.broadcast_to(np.arange(n)[:, None], live.shape)[live]
    cc = np.broadcast_to(cols, live.shape)[live]
    out["action"][rows, cc] = (w >> np.uint32(23)).astype(np.uint16)
    out["visits"][rows, cc] = w & np.uint32(0x7FFFFF)
    out["mask"] = _mask_of_actions(rows, w >> np.uint32(23), n)
    for i in np.nonzero(mf)[0]:
        out["mask"][i] = words[offs[i]:offs[i] + 8]
    return out


# ---- per-worker pickle files ------------------------------------------------------------------------
def _save_one(args):
    records, filename, playouts = args
    memory = to_reference_list(records, playouts)
    with open(filename, 'wb') as file:
        pickle.dump(memory, file)
    return filename


_POOL = {"ex": None, "n": 0}


def _pool(n):
    """process pool for the file writers; workers import numpy and this module only (no CUDA, no torch)"""
    import concurrent.futures as cf
    import multiprocessing as mp
    if _POOL["ex"] is None or _POOL["n"] < n:
        if _POOL["ex"] is not None:
            _POOL["ex"].shutdown()
        _POOL["ex"] = cf.ProcessPoolExecutor(max_workers=n, mp_context=mp.get_context("spawn"))
        _POOL["n"] = n
    return _POOL["ex"]


def shutdown_pool():
    if _POOL["ex"] is not None:
        _POOL["ex"].shutdown()
        _POOL["ex"], _POOL["n"] = None, 0


def save_reference_pickles(jobs, playouts=False, workers=None):
    """jobs: list of (RECORD_DTYPE array, filename).  Converts each array to the reference's list format and
    pickles it, one process per file up to ``workers`` (default: the host's cores); -> filenames in job order.
    Small jobs are done inline (a pool start costs more than converting a few thousand records)."""
    jobs = [(np.ascontiguousarray(r), fn, playouts) for r, fn in jobs]
    total = sum(len(j[0]) for j in jobs)
    if workers is None:
        try:
            workers = len(os.sched_getaffinity(0))
        except Exception:
            workers = os.cpu_count() or 1
    workers = max(1, min(workers, len(jobs), 32))
    if workers == 1 or total < 20000:
        return [_save_one(j) for j in jobs]
    return list(_pool(workers).map(_save_one, jobs))
