"""Conversion of engine records (ck_record, include/ckb200.h) to the reference's self-play record
format: ``[state float64[15,8,8], probs float64[8,8,8], q, z]`` (reference
training_pipeline.py:364-369, 406-411, 439-455), and the ``Keras_Generator`` batch layout
(:288-307)."""
import numpy as np

from . import codec


def to_reference(rec, playouts=False):
    """one RECORD_DTYPE element -> [state, probs, q, z] exactly as the reference pickles it.
    ``playouts``: the record comes from a NEURAL_NET=False search, where rewards are Python ints and
    ``root.q`` is their float64 quotient (MCTS.py:389-394) -- rebuilt from the exact root_w / root_n."""
    pos = tuple(int(v) for v in rec["pos"])
    state = codec.decode_state(pos, [int(v) for v in rec["mask"]], int(rec["plane5"]))
    n = int(rec["n_children"])
    probs = np.zeros(512, dtype=np.float64)
    if n:
        # _create_prob_planes (:421-437): visit counts on the action squares, divided by their sum
        probs[rec["action"][:n].astype(np.int64)] = rec["visits"][:n].astype(np.float64)
        probs = probs.reshape(8, 8, 8)
        probs /= np.sum(probs)
    probs = probs.reshape(8, 8, 8)
    q = np.float32(rec["q"]) if n else int(rec["q"])      # terminal records carry the Python ints 0 / -1 (:407-408)
    if n and playouts:
        q = float(rec["root_w"]) / int(rec["root_n"]) if int(rec["root_n"]) else 0
        if q * float(rec["q"]) < 0:                       # recorded from the root player's point of view (:365-368)
            q = -q
    return [state, probs, q, int(rec["z"])]


def to_reference_list(records, playouts=False):
    return [to_reference(r, playouts) for r in records]


def training_batch(records):
    """(states[:, :14] channels-last, [visit-probs(512), (q+z)/2]) as Keras_Generator.__getitem__"""
    x = np.stack([np.moveaxis(to_reference(r)[0][:14], 0, -1) for r in records]).astype(np.float32)
    probs = np.stack([to_reference(r)[1].reshape(512) for r in records])
    v = np.array([(float(r["q"]) + int(r["z"])) / 2 for r in records])
    return x, [probs, v]
