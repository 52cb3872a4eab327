"""numpy dtypes of the C ABI structs (include/ckb200.h) that do not need the shared library --
importable on a machine where libckb200.so is not built."""
import numpy as np

MAX_CHILDREN = 48
POS_DTYPE = np.dtype([("p1", "<u4"), ("p2", "<u4"), ("k", "<u4"), ("meta", "<u4")])
LEAF_DTYPE = np.dtype([("p1", "<u4"), ("p2", "<u4"), ("k", "<u4"), ("info", "<u4"), ("mask", "<u4", (8,))])
RECORD_DTYPE = np.dtype([
    ("pos", POS_DTYPE), ("mask", "<u4", (8,)), ("plane5", "<i4"), ("n_children", "<i4"),
    ("action", "<u2", (MAX_CHILDREN,)), ("visits", "<u4", (MAX_CHILDREN,)), ("q", "<f4"), ("z", "<i4"),
    ("root_n", "<u4"), ("root_w", "<f4"), ("chosen", "<i4"), ("game", "<i4"), ("ply", "<i4")])
RECORD_HDR_DTYPE = np.dtype([          # ck_record_hdr: packed record header, followed by one uint32 per child in the word stream
    ("pos", POS_DTYPE), ("q", "<f4"), ("root_w", "<f4"), ("root_n", "<u4"), ("game", "<i4"), ("ply", "<u2"), ("chosen", "<i2"),
    ("n_children", "u1"), ("plane5", "u1"), ("z", "i1"), ("flags", "u1")])
GAME_DTYPE = np.dtype([
    ("game", "<i4"), ("outcome", "<i4"), ("move_count", "<i4"), ("terminated", "<i4"),
    ("n_records", "<i4"), ("reroot_misses", "<i4"), ("p1_net", "<i4"), ("reserved", "<i4"),
    ("sims", "<u8"), ("nn_evals", "<u8")])


def records_from_dicts(dicts, game=0):
    """list of record dicts (pos, mask, plane5, actions, visits, q, z, root_n, root_w, chosen) -> RECORD_DTYPE"""
    out = np.zeros(len(dicts), dtype=RECORD_DTYPE)
    for i, d in enumerate(dicts):
        n = len(d["actions"])
        out[i]["pos"] = tuple(int(v) for v in d["pos"])
        out[i]["mask"] = d["mask"]
        out[i]["plane5"], out[i]["n_children"] = d["plane5"], n
        out[i]["action"][:n] = d["actions"]
        out[i]["visits"][:n] = d["visits"]
        out[i]["q"], out[i]["z"] = d["q"], d["z"]
        out[i]["root_n"], out[i]["root_w"], out[i]["chosen"] = d["root_n"], d["root_w"], d["chosen"]
        out[i]["game"], out[i]["ply"] = game, i
    return out
