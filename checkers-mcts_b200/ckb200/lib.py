"""ctypes binding of libckb200.so (include/ckb200.h).

There is no CPU fallback: if the CUDA library is missing the import of this module raises,
and every compute call fails loudly when no CUDA device is usable.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CKB200_LIB") or os.path.join(_HERE, "libckb200.so")     # CKB200_LIB: an experimental build for A/B runs

MAX_CHILDREN = 48
POLICY_SIZE = 512
NET_PARAM_COUNT = 1321774
EVAL_NET, EVAL_UNIFORM_ZERO, EVAL_UNIFORM_MATERIAL, EVAL_HASH = 0, 1, 2, 3
EVAL_HASH_SALTED = 4
EVAL_ROLLOUT, EVAL_ROLLOUT_HASH = 5, 6      # NEURAL_NET=False: UCT + one playout per simulation
EVAL_KINDS = {"net": EVAL_NET, "uniform_zero": EVAL_UNIFORM_ZERO, "uniform_material": EVAL_UNIFORM_MATERIAL,
              "hash": EVAL_HASH, "hash_salted": EVAL_HASH_SALTED, "rollout": EVAL_ROLLOUT,
              "rollout_hash": EVAL_ROLLOUT_HASH}
NET_IMPL_TC, NET_IMPL_SIMT = 0, 1
ERR_NET_RANGE = 8

from .lib_types import GAME_DTYPE, LEAF_DTYPE, POS_DTYPE, RECORD_DTYPE, RECORD_HDR_DTYPE  # noqa: E402,F401


class EngineCfg(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("n_slots", C.c_int32), ("pool_cap", C.c_int32), ("max_plies", C.c_int32),
        ("budget", C.c_int32), ("training", C.c_int32), ("tau_decay_delay", C.c_int32), ("terminate_cnt", C.c_int32),
        ("uct_c", C.c_double), ("alpha", C.c_double), ("epsilon", C.c_double), ("tau", C.c_double),
        ("tau_decay", C.c_double), ("seed", C.c_uint64),
        ("evaluator", C.c_int32), ("evaluator_p2", C.c_int32), ("arena", C.c_int32), ("keep_records", C.c_int32),
        ("reference_tau_quirk", C.c_int32), ("game_id_base", C.c_int32), ("game_id_stride", C.c_int32),
        ("max_terminal_sims_per_step", C.c_int32), ("compact_always", C.c_int32), ("eval_cache_entries", C.c_int32),
        ("max_chain_per_step", C.c_int32), ("stagger_budget", C.c_int32), ("stagger_plies", C.c_int32), ("reserved0", C.c_int32)]


class RunStats(C.Structure):
    _fields_ = [("sims", C.c_uint64), ("nn_evals", C.c_uint64), ("steps", C.c_uint64),
                ("games_finished", C.c_uint64), ("moves", C.c_uint64), ("nodes_created", C.c_uint64),
                ("compactions", C.c_uint64), ("gpu_ms", C.c_double), ("eval_ms", C.c_double),
                ("tower_ms", C.c_double), ("kernel_launches", C.c_uint64), ("cache_hits", C.c_uint64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libckb200 error %d: %s" % (code, msg))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libckb200.so is not built (%s). Run `python checkers-mcts_b200/build.py` (needs nvcc, sm_100a); "
            "there is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    L.ck_last_error.restype = C.c_char_p
    L.ck_abi_version.restype = C.c_int
    L.ck_device_count.restype = C.c_int
    L.ck_movegen.argtypes = [C.c_int, vp, i64, i32, vp, vp, vp, vp, vp]
    L.ck_movegen_device.argtypes = [vp, i64, i32, vp, vp, vp, vp, vp, vp]
    L.ck_rollout.argtypes = [C.c_int, vp, i64, u64, i32, vp, vp]
    L.ck_rollout_device.argtypes = [vp, i64, u64, i32, vp, vp, vp]
    L.ck_net_create.argtypes = [C.c_int]
    L.ck_net_create.restype = vp
    L.ck_net_destroy.argtypes = [vp]
    L.ck_net_destroy.restype = None
    L.ck_net_set_impl.argtypes = [vp, C.c_int]
    L.ck_net_set_weights.argtypes = [vp, vp, i64]
    L.ck_net_set_weights_device.argtypes = [vp, vp, i64]
    L.ck_net_forward.argtypes = [vp, vp, i64, vp, vp]
    L.ck_net_forward_planes.argtypes = [vp, vp, i64, vp, vp]
    L.ck_net_forward_logits.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    L.ck_net_range_status.argtypes = [vp]
    L.ck_net_last_features.argtypes = [vp, i64, vp, vp]
    L.ck_movegen_csr.argtypes = [C.c_int, vp, i64, vp, i64, vp, vp, vp, vp]
    L.ck_movegen_csr_device.argtypes = [vp, i64, vp, i64, vp, vp, vp, vp, vp]
    L.ck_net_forward_device.argtypes = [vp, vp, i64, vp, vp, vp]
    L.ck_mask_renorm.argtypes = [C.c_int, vp, vp, i64, vp]
    L.ck_engine_create.argtypes = [C.POINTER(EngineCfg)]
    L.ck_engine_create.restype = vp
    L.ck_engine_destroy.argtypes = [vp]
    L.ck_engine_destroy.restype = None
    L.ck_engine_set_net.argtypes = [vp, C.c_int, vp]
    L.ck_engine_begin.argtypes = [vp, i64]
    L.ck_engine_run.argtypes = [vp, i64, C.POINTER(RunStats)]
    L.ck_selfplay_run.argtypes = [vp, i64, C.POINTER(RunStats)]
    L.ck_arena_run.argtypes = [vp, i64, C.POINTER(RunStats)]
    L.ck_games_finished.argtypes = [vp]
    L.ck_games_finished.restype = i64
    L.ck_games_fetch.argtypes = [vp, vp, i64]
    L.ck_records_count.argtypes = [vp]
    L.ck_records_count.restype = i64
    L.ck_records_fetch.argtypes = [vp, vp, i64]
    L.ck_records_fetch_new.argtypes = [vp, vp, i64, vp, vp]
    L.ck_records_pack_device.argtypes = [vp, vp, i64, vp, i64, vp, vp]
    L.ck_records_fetch_packed.argtypes = [vp, vp, i64, vp, i64, vp, vp]
    L.ck_records_unpack.argtypes = [vp, i64, vp, i64, vp]
    L.ck_engine_set_profile.argtypes = [vp, C.c_int]
    L.ck_engine_set_budget.argtypes = [vp, i32]
    L.ck_tree_set_root.argtypes = [vp, vp, i32]
    L.ck_tree_search.argtypes = [vp, i32]
    L.ck_tree_root.argtypes = [vp, vp, vp, vp]
    L.ck_tree_root_children.argtypes = [vp, vp, vp, vp, vp, vp]
    L.ck_tree_children.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.ck_tree_reroot.argtypes = [vp, i32]
    L.ck_tree_best_child.argtypes = [vp, i32, vp]
    L.ck_tree_advance.argtypes = [vp, i32]
    L.ck_tree_node_count.argtypes = [vp]
    L.ck_tree_node_count.restype = i64
    L.ck_engine_pool_cap.argtypes = [vp]
    L.ck_engine_pool_cap.restype = i64
    L.ck_tree_epoch.argtypes = [vp]
    L.ck_tree_epoch.restype = i64
    return L


_lib = _load()


def raw():
    return _lib


def check(rc):
    if rc != 0:
        raise CkError(rc, _lib.ck_last_error().decode("utf-8", "replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def device_count():
    return _lib.ck_device_count()


def require_device():
    if device_count() < 1:
        raise CkError(1, "no usable CUDA device; libckb200 has no CPU fallback")


# ---- K1 / K4 -----------------------------------------------------------------------------
def movegen(pos, device=0, max_children=MAX_CHILDREN, want_children=True):
    """pos: POS_DTYPE array -> dict(children [n,max_children] POS_DTYPE, counts, masks [n,8], status, plane5)."""
    pos = np.ascontiguousarray(pos, dtype=POS_DTYPE)
    n = len(pos)
    children = np.zeros((n, max_children), dtype=POS_DTYPE) if want_children else None
    counts = np.zeros(n, dtype=np.int32)
    masks = np.zeros((n, 8), dtype=np.uint32)
    status = np.zeros(n, dtype=np.uint8)
    plane5 = np.zeros(n, dtype=np.uint8)
    check(_lib.ck_movegen(device, _ptr(pos), n, max_children, _ptr(children), _ptr(counts), _ptr(masks),
                          _ptr(status), _ptr(plane5)))
    return dict(children=children, counts=counts, masks=masks, status=status, plane5=plane5)


def movegen_csr(pos, device=0):
    """Packed successors: dict(children [total] POS_DTYPE, offsets [n+1] uint32, masks, status, plane5); the
    successors of position i are children[offsets[i]:offsets[i+1]] in the reference's list order."""
    pos = np.ascontiguousarray(pos, dtype=POS_DTYPE)
    n = len(pos)
    offsets = np.zeros(n + 1, dtype=np.uint32)
    check(_lib.ck_movegen_csr(device, _ptr(pos), n, None, 0, _ptr(offsets), None, None, None))   # sizes only
    total = int(offsets[n])
    children = np.zeros(max(total, 1), dtype=POS_DTYPE)
    masks = np.zeros((n, 8), dtype=np.uint32)
    status = np.zeros(n, dtype=np.uint8)
    plane5 = np.zeros(n, dtype=np.uint8)
    check(_lib.ck_movegen_csr(device, _ptr(pos), n, _ptr(children), total, _ptr(offsets), _ptr(masks), _ptr(status),
                              _ptr(plane5)))
    return dict(children=children[:total], offsets=offsets, masks=masks, status=status, plane5=plane5)


def rollout(pos, seed, device=0, max_plies=0):
    pos = np.ascontiguousarray(pos, dtype=POS_DTYPE)
    n = len(pos)
    outcome = np.zeros(n, dtype=np.uint8)
    plies = np.zeros(n, dtype=np.int32)
    check(_lib.ck_rollout(device, _ptr(pos), n, int(seed), int(max_plies), _ptr(outcome), _ptr(plies)))
    return outcome, plies


def mask_renorm(policy, masks, device=0):
    policy = np.ascontiguousarray(policy, dtype=np.float32).reshape(-1, POLICY_SIZE)
    masks = np.ascontiguousarray(masks, dtype=np.uint32).reshape(-1, 8)
    out = np.empty_like(policy)
    check(_lib.ck_mask_renorm(device, _ptr(policy), _ptr(masks), len(policy), _ptr(out)))
    return out


# ---- K3 ----------------------------------------------------------------------------------
class Net(object):
    """Device-resident policy/value network (create_nn, training_pipeline.py:44-120)."""

    def __init__(self, device=0, impl=None):
        self.device = device
        self._h = _lib.ck_net_create(device)
        if not self._h:
            raise CkError(1, _lib.ck_last_error().decode())
        if impl is not None:
            self.set_impl(impl)

    def set_impl(self, impl):
        check(_lib.ck_net_set_impl(self._h, {"tc": NET_IMPL_TC, "simt": NET_IMPL_SIMT}.get(impl, impl)))

    def set_weights(self, blob):
        """blob: float32 numpy array (host) of NET_PARAM_COUNT values, Keras order/layouts."""
        blob = np.ascontiguousarray(blob, dtype=np.float32).reshape(-1)
        check(_lib.ck_net_set_weights(self._h, _ptr(blob), blob.size))

    def set_weights_device(self, data_ptr, count):
        """device pointer of a float32 blob (e.g. torch_tensor.data_ptr())."""
        check(_lib.ck_net_set_weights_device(self._h, C.c_void_p(int(data_ptr)), int(count)))

    def forward(self, leaves):
        leaves = np.ascontiguousarray(leaves, dtype=LEAF_DTYPE)
        n = len(leaves)
        policy = np.empty((n, POLICY_SIZE), dtype=np.float32)
        value = np.empty(n, dtype=np.float32)
        check(_lib.ck_net_forward(self._h, _ptr(leaves), n, _ptr(policy), _ptr(value)))
        return policy, value

    def forward_logits(self, leaves):
        """-> (policy, value, logits [n,512] before the softmax, value before the tanh): the quantities the
        1e-5 accuracy contract is stated on"""
        leaves = np.ascontiguousarray(leaves, dtype=LEAF_DTYPE)
        n = len(leaves)
        policy = np.empty((n, POLICY_SIZE), dtype=np.float32)
        value = np.empty(n, dtype=np.float32)
        logits = np.empty((n, POLICY_SIZE), dtype=np.float32)
        vpre = np.empty(n, dtype=np.float32)
        check(_lib.ck_net_forward_logits(self._h, _ptr(leaves), n, _ptr(policy), _ptr(value), _ptr(logits), _ptr(vpre)))
        return policy, value, logits, vpre

    def last_features(self, n):
        """tower outputs of the last tensor-core forward of n positions: (pflat [n,512], vconv [n,64])"""
        pflat = np.empty((n, POLICY_SIZE), dtype=np.float32)
        vconv = np.empty((n, 64), dtype=np.float32)
        check(_lib.ck_net_last_features(self._h, n, _ptr(pflat), _ptr(vconv)))
        return pflat, vconv

    def range_status(self):
        """raises CkError(CK_ERR_NET_RANGE) if an activation left the tensor-core path's fp16 range"""
        check(_lib.ck_net_range_status(self._h))

    def predict(self, x):
        """Keras-like predict (Checkers.py:433): x [n,8,8,14] -> [policy [n,512], value [n,1]]."""
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, 8, 8, 14)
        n = len(x)
        policy = np.empty((n, POLICY_SIZE), dtype=np.float32)
        value = np.empty(n, dtype=np.float32)
        check(_lib.ck_net_forward_planes(self._h, _ptr(x), n, _ptr(policy), _ptr(value)))
        return [policy, value.reshape(n, 1)]

    def close(self):
        if self._h:
            _lib.ck_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- engine --------------------------------------------------------------------------------
def records_unpack(hdr, words):
    """packed records (RECORD_HDR_DTYPE array, uint32 child words) -> RECORD_DTYPE array; host-side C twin of
    ckb200.records.unpack (no device needed), multi-threaded"""
    hdr = np.ascontiguousarray(hdr, dtype=RECORD_HDR_DTYPE)
    words = np.ascontiguousarray(words, dtype=np.uint32)
    out = np.empty(len(hdr), dtype=RECORD_DTYPE)
    rc = _lib.ck_records_unpack(_ptr(hdr) if len(hdr) else None, len(hdr), _ptr(words) if len(words) else None, len(words),
                                _ptr(out) if len(out) else None)
    if rc != 0:
        raise ValueError(_lib.ck_last_error().decode())
    return out


def make_cfg(n_slots, budget, device=0, uct_c=4.0, training=False, alpha=1.0, epsilon=0.0, tau=0.0,
             tau_decay=0.0, tau_decay_delay=0, terminate_cnt=0, seed=1, evaluator="net", evaluator_p2=None,
             arena=False, keep_records=True, pool_cap=0, max_plies=0, reference_tau_quirk=False,
             game_id_base=0, game_id_stride=1, max_terminal_sims_per_step=0, compact_always=False,
             eval_cache_entries=0, max_chain_per_step=0, stagger_budget=0, stagger_plies=0):
    ev = EVAL_KINDS[evaluator] if isinstance(evaluator, str) else int(evaluator)
    ev2 = -1 if evaluator_p2 is None else (EVAL_KINDS[evaluator_p2] if isinstance(evaluator_p2, str) else int(evaluator_p2))
    return EngineCfg(device=device, n_slots=n_slots, pool_cap=pool_cap, max_plies=max_plies, budget=budget,
                     training=int(bool(training)), tau_decay_delay=tau_decay_delay, terminate_cnt=terminate_cnt,
                     uct_c=uct_c, alpha=alpha, epsilon=epsilon, tau=tau, tau_decay=tau_decay, seed=seed,
                     evaluator=ev, evaluator_p2=ev2, arena=int(bool(arena)), keep_records=int(bool(keep_records)),
                     reference_tau_quirk=int(bool(reference_tau_quirk)), game_id_base=game_id_base,
                     game_id_stride=game_id_stride, max_terminal_sims_per_step=max_terminal_sims_per_step,
                     compact_always=int(bool(compact_always)), eval_cache_entries=int(eval_cache_entries),
                     max_chain_per_step=int(max_chain_per_step), stagger_budget=int(stagger_budget),
                     stagger_plies=int(stagger_plies), reserved0=0)


class Engine(object):
    def __init__(self, cfg):
        self.cfg = cfg
        self._nets = [None, None]
        self._h = _lib.ck_engine_create(C.byref(cfg))
        if not self._h:
            raise CkError(1, _lib.ck_last_error().decode())

    def set_net(self, which, net):
        self._nets[which] = net          # keep alive
        check(_lib.ck_engine_set_net(self._h, which, net._h))

    def set_budget(self, budget):
        check(_lib.ck_engine_set_budget(self._h, int(budget)))
        self.cfg.budget = int(budget)

    def set_profile(self, on):
        check(_lib.ck_engine_set_profile(self._h, int(bool(on))))

    def begin(self, n_games):
        check(_lib.ck_engine_begin(self._h, int(n_games)))

    def run(self, n_steps=0):
        st = RunStats()
        check(_lib.ck_engine_run(self._h, int(n_steps), C.byref(st)))
        return st.as_dict()

    def selfplay(self, n_games):
        st = RunStats()
        check(_lib.ck_selfplay_run(self._h, int(n_games), C.byref(st)))
        return st.as_dict()

    def arena(self, n_games):
        st = RunStats()
        check(_lib.ck_arena_run(self._h, int(n_games), C.byref(st)))
        return st.as_dict()

    def games_finished(self):
        return int(_lib.ck_games_finished(self._h))

    def games(self):
        n = self.games_finished()
        out = np.zeros(max(n, 1), dtype=GAME_DTYPE)
        check(_lib.ck_games_fetch(self._h, _ptr(out), len(out)))
        return out[:n]

    def records(self):
        n = int(_lib.ck_records_count(self._h))
        out = np.zeros(max(n, 1), dtype=RECORD_DTYPE)
        if n:
            check(_lib.ck_records_fetch(self._h, _ptr(out), len(out)))
        return out[:n]

    def records_packed_sizes(self):
        """(n_records, n_words) of the packed form of all finished games' records"""
        n, w = C.c_int64(), C.c_int64()
        check(_lib.ck_records_pack_device(self._h, None, 0, None, 0, C.byref(n), C.byref(w)))
        return n.value, w.value

    def records_pack_device(self, hdr_ptr, hdr_cap, words_ptr, word_cap):
        """pack into caller-provided device buffers (raw pointers, e.g. torch tensors' data_ptr()); -> (n_records, n_words)"""
        n, w = C.c_int64(), C.c_int64()
        check(_lib.ck_records_pack_device(self._h, C.c_void_p(int(hdr_ptr)), int(hdr_cap), C.c_void_p(int(words_ptr)), int(word_cap),
                                          C.byref(n), C.byref(w)))
        return n.value, w.value

    def records_packed(self):
        """all finished games' records in packed form on the host: (RECORD_HDR_DTYPE array, uint32 child words);
        ckb200.records.unpack turns them into RECORD_DTYPE"""
        n, w = self.records_packed_sizes()
        hdr = np.zeros(max(n, 1), dtype=RECORD_HDR_DTYPE)
        words = np.zeros(max(w, 1), dtype=np.uint32)
        if n:
            a, b = C.c_int64(), C.c_int64()
            check(_lib.ck_records_fetch_packed(self._h, _ptr(hdr), len(hdr), _ptr(words), len(words), C.byref(a), C.byref(b)))
        return hdr[:n], words[:w]

    def records_new(self, buf):
        """records of games finished since the last call, written into ``buf`` (RECORD_DTYPE array);
        -> (n_records, n_games)"""
        n = C.c_int64()
        g = C.c_int64()
        check(_lib.ck_records_fetch_new(self._h, _ptr(buf), len(buf), C.byref(n), C.byref(g)))
        return n.value, g.value

    # single-search API (slot 0)
    def tree_set_root(self, pos, parent_player=-1):
        p = np.array([tuple(int(v) for v in pos)], dtype=POS_DTYPE)
        check(_lib.ck_tree_set_root(self._h, _ptr(p), int(parent_player)))

    def tree_search(self, sims):
        check(_lib.ck_tree_search(self._h, int(sims)))

    def tree_root(self):
        n = C.c_uint32()
        w = C.c_float()
        b = C.c_int32()
        check(_lib.ck_tree_root(self._h, C.byref(n), C.byref(w), C.byref(b)))
        return n.value, np.float32(w.value), b.value

    def tree_children(self, node=-1):
        """children of ``node`` (-1: root) -> list of dicts with idx, pos, n, w, p, terminal (CK_* status)"""
        idx = np.zeros(MAX_CHILDREN, dtype=np.int32)
        pos = np.zeros(MAX_CHILDREN, dtype=POS_DTYPE)
        n = np.zeros(MAX_CHILDREN, dtype=np.uint32)
        w = np.zeros(MAX_CHILDREN, dtype=np.float32)
        p = np.zeros(MAX_CHILDREN, dtype=np.float32)
        st = np.zeros(MAX_CHILDREN, dtype=np.int32)
        b = C.c_int32()
        check(_lib.ck_tree_children(self._h, int(node), _ptr(idx), _ptr(pos), _ptr(n), _ptr(w), _ptr(p), _ptr(st), C.byref(b)))
        if self.cfg.evaluator in (EVAL_ROLLOUT, EVAL_ROLLOUT_HASH):
            p[:] = 0.0       # no priors without a network (MCTS_Node.p stays 0); the engine keeps its own count in that field
        return [dict(idx=int(idx[i]), pos=tuple(int(v) for v in pos[i]), n=int(n[i]), w=np.float32(w[i]),
                     p=np.float32(p[i]), terminal=int(st[i])) for i in range(b.value)]

    def tree_root_children(self):
        return self.tree_children(-1)

    def tree_reroot(self, node):
        check(_lib.ck_tree_reroot(self._h, int(node)))

    def tree_best_child(self, move_count=0):
        idx = C.c_int32()
        check(_lib.ck_tree_best_child(self._h, int(move_count), C.byref(idx)))
        return idx.value

    def tree_advance(self, child_index):
        check(_lib.ck_tree_advance(self._h, int(child_index)))

    def tree_node_count(self):
        return int(_lib.ck_tree_node_count(self._h))

    def pool_cap(self):
        return int(_lib.ck_engine_pool_cap(self._h))

    def tree_epoch(self):
        return int(_lib.ck_tree_epoch(self._h))

    def close(self):
        if self._h:
            _lib.ck_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
