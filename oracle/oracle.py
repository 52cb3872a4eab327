"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the C oracle (oracle/ck_oracle.c).

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libckoracle.so")

MAX_CHILDREN = 48
POS_DTYPE = np.dtype([("p1", "<u4"), ("p2", "<u4"), ("k", "<u4"), ("meta", "<u4")])


class Pos(C.Structure):
    _fields_ = [("p1", C.c_uint32), ("p2", C.c_uint32), ("k", C.c_uint32), ("meta", C.c_uint32)]

    def tup(self):
        return (self.p1, self.p2, self.k, self.meta)


class Cfg(C.Structure):
    _fields_ = [("uct_c", C.c_double), ("budget", C.c_int32), ("training", C.c_int32),
                ("alpha", C.c_double), ("epsilon", C.c_double),
                ("tau", C.c_double), ("tau_decay", C.c_double),
                ("tau_decay_delay", C.c_int32), ("terminate_cnt", C.c_int32),
                ("seed", C.c_uint64), ("rollout", C.c_int32), ("reserved", C.c_int32)]


class Record(C.Structure):
    _fields_ = [("pos", Pos), ("mask", C.c_uint32 * 8), ("plane5", C.c_int32),
                ("n_children", C.c_int32), ("action", C.c_uint16 * MAX_CHILDREN),
                ("visits", C.c_uint32 * MAX_CHILDREN), ("q", C.c_float), ("z", C.c_int32),
                ("root_n", C.c_uint32), ("root_w", C.c_float), ("chosen", C.c_int32)]


EVAL_FN = C.CFUNCTYPE(None, C.POINTER(Pos), C.POINTER(C.c_uint32), C.c_int,
                      C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p)


def build(force=False):
    src = os.path.join(_HERE, "ck_oracle.c")
    hdr = os.path.join(_HERE, "ck_oracle.h")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _SO
    subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    PP = C.POINTER(Pos)
    U32P = C.POINTER(C.c_uint32)
    FP = C.POINTER(C.c_float)
    IP = C.POINTER(C.c_int)
    L.cko_movegen.argtypes = [PP, PP, U32P, IP, IP]
    L.cko_movegen.restype = C.c_int
    L.cko_start_position.argtypes = [PP]
    L.cko_perft.argtypes = [PP, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.cko_perft.restype = C.c_uint64
    L.cko_mask_renorm.argtypes = [FP, U32P, FP]
    L.cko_tree_new.argtypes = [PP, C.c_int, C.POINTER(Cfg), C.c_void_p, C.c_void_p]
    L.cko_tree_new.restype = C.c_void_p
    L.cko_tree_free.argtypes = [C.c_void_p]
    L.cko_tree_search.argtypes = [C.c_void_p, C.c_int]
    L.cko_tree_root_children.argtypes = [C.c_void_p, PP, U32P, FP, FP, C.POINTER(C.c_int32)]
    L.cko_tree_root_children.restype = C.c_int
    L.cko_tree_root_stats.argtypes = [C.c_void_p, U32P, FP]
    L.cko_tree_best_child.argtypes = [C.c_void_p, C.c_int]
    L.cko_tree_best_child.restype = C.c_int
    L.cko_tree_node_count.argtypes = [C.c_void_p]
    L.cko_tree_node_count.restype = C.c_uint64
    L.cko_tree_nn_evals.argtypes = [C.c_void_p]
    L.cko_tree_nn_evals.restype = C.c_uint64
    L.cko_game_new.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cko_game_new.restype = C.c_void_p
    L.cko_game_free.argtypes = [C.c_void_p]
    L.cko_game_play_ply.argtypes = [C.c_void_p]
    L.cko_game_play_ply.restype = C.c_int
    L.cko_game_state.argtypes = [C.c_void_p, PP]
    for name in ("outcome", "move_count", "terminated", "num_records"):
        f = getattr(L, "cko_game_" + name)
        f.argtypes = [C.c_void_p]
        f.restype = C.c_int
    L.cko_game_get_records.argtypes = [C.c_void_p, C.POINTER(Record)]
    for name in ("total_sims", "nn_evals", "reroot_misses"):
        f = getattr(L, "cko_game_" + name)
        f.argtypes = [C.c_void_p]
        f.restype = C.c_uint64
    L.cko_random_playout.argtypes = [PP, C.POINTER(C.c_uint64), IP, C.c_int]
    L.cko_random_playout.restype = C.c_int
    L.cko_hash_playout.argtypes = [PP, C.c_uint32, IP]
    L.cko_hash_playout.restype = C.c_int
    L.cko_hash_choice.argtypes = [PP, C.c_uint32]
    L.cko_hash_choice.restype = C.c_uint32
    _lib = L
    return L


def _pos(t):
    return Pos(*[int(v) for v in t])


def start_position():
    p = Pos()
    lib().cko_start_position(C.byref(p))
    return p.tup()


def movegen(pos):
    """-> (children [(p1,p2,k,meta)], mask[8], status, plane5)"""
    ch = (Pos * MAX_CHILDREN)()
    mask = (C.c_uint32 * 8)()
    st = C.c_int()
    p5 = C.c_int()
    n = lib().cko_movegen(C.byref(_pos(pos)), ch, mask, C.byref(st), C.byref(p5))
    return [ch[i].tup() for i in range(n)], list(mask), st.value, p5.value


def perft(pos, depth):
    hops = C.c_uint64(0)
    cont = C.c_uint64(0)
    n = lib().cko_perft(C.byref(_pos(pos)), depth, C.byref(hops), C.byref(cont))
    return n, hops.value, cont.value


def mask_renorm(policy, mask):
    pin = np.ascontiguousarray(policy, dtype=np.float32).reshape(512)
    pout = np.empty(512, dtype=np.float32)
    m = (C.c_uint32 * 8)(*[int(v) for v in mask])
    lib().cko_mask_renorm(pin.ctypes.data_as(C.POINTER(C.c_float)), m,
                          pout.ctypes.data_as(C.POINTER(C.c_float)))
    return pout


def make_cfg(uct_c=4.0, budget=400, training=False, alpha=1.0, epsilon=0.0, tau=0.0,
             tau_decay=0.0, tau_decay_delay=0, terminate_cnt=0, seed=1, rollout=None):
    """rollout: None (NEURAL_NET=True), 'random' or 'hash' (NEURAL_NET=False: UCT + one playout per simulation)"""
    return Cfg(float(uct_c), int(budget), int(bool(training)), float(alpha), float(epsilon),
               float(tau), float(tau_decay), int(tau_decay_delay), int(terminate_cnt), int(seed),
               {None: 0, "random": 1, "hash": 2}[rollout], 0)


BUILTIN_EVALS = ("uniform_zero", "uniform_material", "hash", "hash_salted")


def _resolve_eval(ev):
    """-> (c function pointer as void*, keepalive)"""
    if isinstance(ev, str):
        f = getattr(lib(), "cko_eval_" + ev)
        return C.cast(f, C.c_void_p), None
    cb = EVAL_FN(ev)
    return C.cast(cb, C.c_void_p), cb


def python_eval(fn):
    """Wrap ``fn(pos_tuple, mask_list, plane5) -> (policy512 float32, value)`` as a callback."""
    def _cb(pos_p, mask_p, plane5, pol_p, val_p, _ctx):
        pos = pos_p.contents.tup()
        mask = [mask_p[i] for i in range(8)]
        pol, val = fn(pos, mask, plane5)
        dst = np.ctypeslib.as_array(pol_p, shape=(512,))
        dst[:] = np.asarray(pol, dtype=np.float32).reshape(512)
        val_p[0] = float(np.float32(val))
    return _cb


class Tree(object):
    def __init__(self, root, cfg, evaluator="uniform_zero", parent_player=-1):
        self._fn, self._keep = _resolve_eval(evaluator)
        self.cfg = cfg
        self._h = lib().cko_tree_new(C.byref(_pos(root)), parent_player, C.byref(cfg), self._fn, None)

    def search(self, sims):
        lib().cko_tree_search(self._h, sims)

    def root_stats(self):
        n = C.c_uint32()
        w = C.c_float()
        lib().cko_tree_root_stats(self._h, C.byref(n), C.byref(w))
        return n.value, np.float32(w.value)

    def root_children(self):
        pos = (Pos * MAX_CHILDREN)()
        n = (C.c_uint32 * MAX_CHILDREN)()
        w = (C.c_float * MAX_CHILDREN)()
        p = (C.c_float * MAX_CHILDREN)()
        t = (C.c_int32 * MAX_CHILDREN)()
        b = lib().cko_tree_root_children(self._h, pos, n, w, p, t)
        return [dict(pos=pos[i].tup(), n=n[i], w=np.float32(w[i]), p=np.float32(p[i]), terminal=t[i])
                for i in range(b)]

    def best_child(self, move_count=0):
        return lib().cko_tree_best_child(self._h, move_count)

    def node_count(self):
        return lib().cko_tree_node_count(self._h)

    def nn_evals(self):
        return lib().cko_tree_nn_evals(self._h)

    def close(self):
        if self._h:
            lib().cko_tree_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Game(object):
    def __init__(self, cfg, eval_p1="uniform_zero", eval_p2=None, salt=None):
        """salt: uint32 handed to the evaluators as ctx (only 'hash_salted' reads it)."""
        self._f1, self._k1 = _resolve_eval(eval_p1)
        if eval_p2 is None:
            self._f2, self._k2 = None, None
        else:
            self._f2, self._k2 = _resolve_eval(eval_p2)
        self.cfg = cfg
        self._salt = C.c_uint32(0 if salt is None else int(salt))
        ctx = C.cast(C.pointer(self._salt), C.c_void_p)
        self._h = lib().cko_game_new(C.byref(cfg), self._f1, ctx, self._f2, ctx)

    def play_ply(self):
        return bool(lib().cko_game_play_ply(self._h))

    def play(self, max_plies=None):
        k = 0
        while self.play_ply():
            k += 1
            if max_plies is not None and k >= max_plies:
                break
        return self

    def state(self):
        p = Pos()
        lib().cko_game_state(self._h, C.byref(p))
        return p.tup()

    outcome = property(lambda s: lib().cko_game_outcome(s._h))
    move_count = property(lambda s: lib().cko_game_move_count(s._h))
    terminated = property(lambda s: bool(lib().cko_game_terminated(s._h)))
    total_sims = property(lambda s: lib().cko_game_total_sims(s._h))
    nn_evals = property(lambda s: lib().cko_game_nn_evals(s._h))
    reroot_misses = property(lambda s: lib().cko_game_reroot_misses(s._h))

    def records(self):
        n = lib().cko_game_num_records(self._h)
        buf = (Record * max(n, 1))()
        lib().cko_game_get_records(self._h, buf)
        out = []
        for i in range(n):
            r = buf[i]
            out.append(dict(pos=r.pos.tup(), mask=list(r.mask), plane5=r.plane5,
                            actions=[r.action[j] for j in range(r.n_children)],
                            visits=[r.visits[j] for j in range(r.n_children)],
                            q=np.float32(r.q), z=r.z, root_n=r.root_n, root_w=np.float32(r.root_w),
                            chosen=r.chosen))
        return out

    def close(self):
        if self._h:
            lib().cko_game_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def random_playout(pos, seed, max_plies=0):
    st = C.c_uint64(seed)
    plies = C.c_int()
    out = lib().cko_random_playout(C.byref(_pos(pos)), C.byref(st), C.byref(plies), max_plies)
    return out, plies.value


def hash_playout(pos, salt=0):
    """-> (outcome, plies) of the deterministic playout (index = hash_choice(position, n_legal))"""
    plies = C.c_int()
    out = lib().cko_hash_playout(C.byref(_pos(pos)), int(salt), C.byref(plies))
    return out, plies.value


def hash_choice(pos, n):
    return int(lib().cko_hash_choice(C.byref(_pos(pos)), int(n)))
