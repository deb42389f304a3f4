"""ctypes binding of libpprb200.so (C ABI in include/pprb200.h).

The library is the product: if it is missing this module raises -- there is no Python or CPU
fallback for any query.  Loading the library needs no GPU (the exported symbols can be inspected and
the host-only R-MAT generator used on a CPU box); every compute entry point needs one.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# PPRB200_LIB: tuning runs load an A/B build of the same library (csrc/Makefile `variant`); never a different implementation
LIB_PATH = os.environ.get("PPRB200_LIB") or os.path.join(_PKG, "libpprb200.so")

PPR_OK = 0
PPR_E_INVALID, PPR_E_CUDA, PPR_E_NOMEM, PPR_E_RANGE, PPR_E_CAPACITY, PPR_E_STATE, PPR_E_CALLBACK = -1, -2, -3, -4, -5, -6, -7
RESTART_AT_WALK_START = 0
RESTART_AT_SOURCE = 1
FX_ONE = 1 << 62
FX_INV = 1.0 / float(1 << 62)

# every symbol include/pprb200.h (product ABI) and include/pprb200_debug.h (measurement / test hooks) declare;
# checked by tests/test_abi.py against the header texts
EXPORTS = [
    "ppr_init", "ppr_shutdown", "ppr_set_tuning", "ppr_set_bwd_batch", "ppr_set_batch_streams", "ppr_device_sync", "ppr_last_error", "ppr_device_info", "ppr_set_balance", "ppr_host_alloc",
    "ppr_host_free", "ppr_graph_create", "ppr_graph_destroy", "ppr_graph_node_count", "ppr_graph_rel_count",
    "ppr_graph_export_csc", "ppr_forward_push", "ppr_monte_carlo", "ppr_fora", "ppr_fora_topk", "ppr_topk",
    "ppr_power_iteration", "ppr_fora_batch", "ppr_backward_search", "ppr_backward_all_pairs", "ppr_pairs_count",
    "ppr_pairs_export", "ppr_pairs_destroy", "ppr_fora_shard", "ppr_monte_carlo_shard", "ppr_fora_topk_shard", "ppr_result_devptr",
    "ppr_result_export", "ppr_walk_terminals", "ppr_version",
    "ppr_comm_unique_id", "ppr_comm_init", "ppr_comm_destroy", "ppr_comm_info", "ppr_comm_max_f64", "ppr_fora_sharded", "ppr_monte_carlo_sharded",
    "ppr_fora_topk_sharded", "ppr_backward_all_pairs_sharded", "ppr_metrics_keep", "ppr_metrics_compute",
]
DEBUG_EXPORTS = ["ppr_timer_record", "ppr_timer_elapsed_ms", "ppr_flush_l2", "ppr_set_walk_variant", "ppr_set_pull", "ppr_last_raw", "ppr_push_trace_begin",
                 "ppr_push_trace_step", "ppr_debug_timeline"]
GEN_EXPORTS = ["ppr_rmat_generate"]   # include/pprgen.h, libpprgen.so (host only)
GEN_PATH = os.path.join(_PKG, "libpprgen.so")
PPR_E_COMM = -8
BALANCE_PER_SHARD = 2


# ppr_reduce_fn: int32 (*)(void* user, uint64 dev_ptr, int64 n)
REDUCE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_uint64, C.c_int64)


class Stats(C.Structure):
    _fields_ = [
        ("pushes", C.c_int64), ("edges", C.c_int64), ("iters", C.c_int64), ("rounds", C.c_int64),
        ("walks", C.c_int64), ("walks_total", C.c_int64), ("steps", C.c_int64), ("walk_sources", C.c_int64),
        ("rmax_fx", C.c_uint64), ("rsum_fx", C.c_uint64), ("omega", C.c_double),
        ("ms_push", C.c_float), ("ms_walk", C.c_float), ("ms_select", C.c_float), ("ms_total", C.c_float),
        ("launches", C.c_int32), ("ms_comm", C.c_float),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Conf(C.Structure):
    """ppr_conf: delta / pfail / min_delta handed to the algorithm constructors (<= 0: Algo_Conf default)."""
    _fields_ = [("delta", C.c_double), ("pfail", C.c_double), ("min_delta", C.c_double)]


def _conf(conf):
    if conf is None:
        return None
    if isinstance(conf, Conf):
        return C.byref(conf)
    return C.byref(Conf(conf.get("delta", 0.0) or 0.0, conf.get("pfail", 0.0) or 0.0, conf.get("min_delta", 0.0) or 0.0))


class PPRError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("pprb200 error %d: %s" % (code, msg))
        self.code = code


_lib = None
_gen = None


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def load():
    """dlopen libpprb200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libpprb200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C %s`. "
            "There is no CPU fallback." % (LIB_PATH, os.path.join(_PKG, "csrc")))
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    SP = C.POINTER(Stats)
    sig = {
        "ppr_init": (i32, [i32, C.POINTER(vp)]),
        "ppr_shutdown": (i32, [vp]),
        "ppr_last_error": (C.c_char_p, [vp]),
        "ppr_device_info": (i32, [vp, C.POINTER(i32), C.POINTER(i64), C.c_char_p, i32]),
        "ppr_set_balance": (i32, [vp, f64, i32]),
        "ppr_set_tuning": (i32, [vp, i64, i32]),
        "ppr_set_bwd_batch": (i32, [vp, i32]),
        "ppr_set_batch_streams": (i32, [vp, i32]),
        "ppr_timer_record": (i32, [vp, i32]),
        "ppr_timer_elapsed_ms": (i32, [vp, i32, i32, C.POINTER(C.c_float)]),
        "ppr_flush_l2": (i32, [vp]),
        "ppr_device_sync": (i32, [vp]),
        "ppr_host_alloc": (i32, [vp, i64, C.POINTER(vp)]),
        "ppr_host_free": (i32, [vp, vp]),
        "ppr_graph_create": (i32, [vp, i64, i64, vp, vp, C.POINTER(vp)]),
        "ppr_graph_destroy": (i32, [vp]),
        "ppr_graph_node_count": (i64, [vp]),
        "ppr_graph_rel_count": (i64, [vp]),
        "ppr_graph_export_csc": (i32, [vp, vp, vp]),
        "ppr_forward_push": (i32, [vp, i32, f64, f64, vp, vp, C.POINTER(f64), SP]),
        "ppr_monte_carlo": (i32, [vp, i32, f64, f64, u64, vp, vp, SP]),
        "ppr_fora": (i32, [vp, i32, f64, f64, u64, i32, vp, vp, SP]),
        "ppr_fora_topk": (i32, [vp, i32, f64, f64, i32, u64, i32, vp, vp, vp, i32, C.POINTER(i32), SP]),
        "ppr_topk": (i32, [vp, vp, i64, i32, vp, vp, i32, C.POINTER(i32), C.POINTER(f64)]),
        "ppr_power_iteration": (i32, [vp, i32, f64, i32, vp]),
        "ppr_fora_batch": (i32, [vp, vp, i32, f64, f64, u64, i32, vp, vp, SP]),
        "ppr_backward_search": (i32, [vp, i32, f64, f64, vp, SP]),
        "ppr_backward_all_pairs": (i32, [vp, f64, f64, i32, i32, i32, C.POINTER(vp), SP]),
        "ppr_pairs_count": (i64, [vp]),
        "ppr_pairs_export": (i32, [vp, vp, vp, vp, vp]),
        "ppr_pairs_destroy": (i32, [vp]),
        "ppr_fora_shard": (i32, [vp, i32, f64, f64, u64, i32, vp, i32, i32, SP]),
        "ppr_monte_carlo_shard": (i32, [vp, i32, f64, f64, u64, vp, i32, i32, SP]),
        "ppr_fora_topk_shard": (i32, [vp, i32, f64, f64, i32, u64, i32, vp, i32, i32, REDUCE_FN, vp, vp, vp, i32, C.POINTER(i32), SP]),
        "ppr_result_devptr": (i32, [vp, C.POINTER(u64), C.POINTER(i64)]),
        "ppr_result_export": (i32, [vp, f64, vp]),
        "ppr_last_raw": (i32, [vp, vp, vp]),
        "ppr_push_trace_begin": (i32, [vp, i32, f64, f64]),
        "ppr_push_trace_step": (i32, [vp, vp, i64, C.POINTER(i64)]),
        "ppr_walk_terminals": (i32, [vp, i32, i32, u64, i64, u64, u64, f64, i32, vp]),
        "ppr_debug_timeline": (i32, [vp, vp, i64, C.POINTER(i64)]),
        "ppr_version": (C.c_char_p, []),
        "ppr_set_walk_variant": (i32, [vp, i32]),
        "ppr_set_pull": (i32, [vp, f64, f64]),
        "ppr_metrics_keep": (i32, [vp, i32]),
        "ppr_metrics_compute": (i32, [vp, i32, C.POINTER(f64), C.POINTER(f64), C.POINTER(f64)]),
        "ppr_comm_unique_id": (i32, [vp]),
        "ppr_comm_init": (i32, [vp, vp, i32, i32]),
        "ppr_comm_destroy": (i32, [vp]),
        "ppr_comm_info": (i32, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "ppr_comm_max_f64": (i32, [vp, vp, i32]),
        "ppr_fora_sharded": (i32, [vp, i32, f64, f64, u64, i32, vp, vp, SP]),
        "ppr_monte_carlo_sharded": (i32, [vp, i32, f64, f64, u64, vp, vp, SP]),
        "ppr_fora_topk_sharded": (i32, [vp, i32, f64, f64, i32, u64, i32, vp, vp, vp, i32, C.POINTER(i32), SP]),
        "ppr_backward_all_pairs_sharded": (i32, [vp, f64, f64, i32, C.POINTER(vp), SP]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def load_gen():
    """dlopen libpprgen.so, the host-only synthetic-input generator (never maps libpprb200.so)."""
    global _gen
    if _gen is None:
        if not os.path.exists(GEN_PATH):
            raise ImportError("libpprgen.so is not built (%s). Run `make -C %s`." % (GEN_PATH, os.path.join(_PKG, "csrc")))
        G = C.CDLL(GEN_PATH)
        G.ppr_rmat_generate.restype = C.c_int32
        G.ppr_rmat_generate.argtypes = [C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        _gen = G
    return _gen


def rmat(n, m, a=0.57, b=0.19, c=0.19, seed=42, permute=True, symmetric=False):
    """Seeded R-MAT graph as canonical CSR (host only, no GPU needed)."""
    G = load_gen()
    row_ptr = np.zeros(n + 1, np.int64)
    col_idx = np.zeros(max(m, 1), np.int32)
    rc = G.ppr_rmat_generate(n, m, a, b, c, seed, int(permute), int(symmetric), _ptr(row_ptr), _ptr(col_idx))
    if rc != PPR_OK:
        raise PPRError(rc, "ppr_rmat_generate failed")
    return row_ptr, col_idx[:m]


class Context:
    """One per GPU.  Mirrors the driver's database lifecycle (PPR.createDb / shutDown, J/PPR.java:52-68)."""

    def __init__(self, device=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.ppr_init(device, C.byref(h))
        if rc != PPR_OK:
            raise PPRError(rc, (self.L.ppr_last_error(None) or b"").decode())
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != PPR_OK:
            raise PPRError(rc, (self.L.ppr_last_error(self.h) or b"").decode())

    def set_balance(self, walk_cost=-1.0, max_rounds=64):
        self.check(self.L.ppr_set_balance(self.h, walk_cost, max_rounds))

    def device_info(self):
        sm, mem, name = C.c_int32(), C.c_int64(), C.create_string_buffer(256)
        self.check(self.L.ppr_device_info(self.h, C.byref(sm), C.byref(mem), name, 256))
        return {"sm_count": sm.value, "total_mem": mem.value, "name": name.value.decode()}

    def set_tuning(self, scan_threshold_edges=-1, hot_min_indeg=-1):
        self.check(self.L.ppr_set_tuning(self.h, scan_threshold_edges, hot_min_indeg))

    def set_batch_streams(self, streams=0):
        self.check(self.L.ppr_set_batch_streams(self.h, streams))

    def set_bwd_batch(self, targets_per_batch=0):
        self.check(self.L.ppr_set_bwd_batch(self.h, targets_per_batch))

    def set_walk_variant(self, hint=-1):
        self.check(self.L.ppr_set_walk_variant(self.h, hint))

    def set_pull(self, enter_share=-1.0, exit_share=0.0):
        """Direction switch of the forward push (pull-direction passes for fat iterations); enter_share < 0 = off, 0/0 = forced."""
        self.check(self.L.ppr_set_pull(self.h, enter_share, exit_share))

    # ---- multi-GPU (NCCL inside the library)
    @staticmethod
    def comm_unique_id():
        buf = (C.c_char * 128)()
        rc = load().ppr_comm_unique_id(buf)
        if rc != PPR_OK:
            raise PPRError(rc, (load().ppr_last_error(None) or b"").decode())
        return bytes(buf)

    def comm_init(self, uid, rank, world):
        self.check(self.L.ppr_comm_init(self.h, C.c_char_p(uid), rank, world))
        self.rank, self.world = rank, world

    def comm_destroy(self):
        self.check(self.L.ppr_comm_destroy(self.h))

    def comm_info(self):
        r, w, v = C.c_int32(), C.c_int32(), C.c_int32()
        self.check(self.L.ppr_comm_info(self.h, C.byref(r), C.byref(w), C.byref(v)))
        return {"rank": r.value, "world": w.value, "nccl_version": v.value}

    def comm_max(self, values):
        a = np.ascontiguousarray(values, np.float64).copy()
        self.check(self.L.ppr_comm_max_f64(self.h, _ptr(a), a.shape[0]))
        return a

    def timer_record(self, slot):
        self.check(self.L.ppr_timer_record(self.h, slot))

    def timer_elapsed_ms(self, a, b):
        ms = C.c_float()
        self.check(self.L.ppr_timer_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def flush_l2(self):
        self.check(self.L.ppr_flush_l2(self.h))

    def sync(self):
        self.check(self.L.ppr_device_sync(self.h))

    def pinned(self, shape, dtype):
        """numpy array backed by pinned host memory (freed with the context)."""
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        p = C.c_void_p()
        self.check(self.L.ppr_host_alloc(self.h, max(1, count * dtype.itemsize), C.byref(p)))
        buf = (C.c_char * max(1, count * dtype.itemsize)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
        self._pins = getattr(self, "_pins", [])
        self._pins.append(p)
        return arr

    def topk(self, vec, k):
        vec = np.ascontiguousarray(vec, np.float64)
        n = vec.shape[0]
        ids, vals = np.zeros(n, np.int32), np.zeros(n, np.float64)
        cnt, kth = C.c_int32(), C.c_double()
        self.check(self.L.ppr_topk(self.h, _ptr(vec), n, k, _ptr(ids), _ptr(vals), n, C.byref(cnt), C.byref(kth)))
        return ids[:cnt.value].copy(), vals[:cnt.value].copy(), kth.value

    def close(self):
        if self.h:
            for p in getattr(self, "_pins", []):
                self.L.ppr_host_free(self.h, p)
            self._pins = []
            self.L.ppr_shutdown(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Graph:
    """Device-resident CSR + CSC of the topology (replaces the HeavyGraph `adjM`, J/PPR.java:141-147)."""

    def __init__(self, ctx, row_ptr, col_idx):
        self.ctx = ctx
        self.L = ctx.L
        row_ptr = np.ascontiguousarray(row_ptr, np.int64)
        col_idx = np.ascontiguousarray(col_idx, np.int32)
        self.n = int(row_ptr.shape[0] - 1)
        self.m = int(col_idx.shape[0])
        self.out_degree = np.diff(row_ptr)
        h = C.c_void_p()
        ctx.check(self.L.ppr_graph_create(ctx.h, self.n, self.m, _ptr(row_ptr), _ptr(col_idx), C.byref(h)))
        self.h = h

    # ---- Graph accessors used by the reference
    def nodeCount(self):
        return int(self.L.ppr_graph_node_count(self.h))

    def relCount(self):
        return int(self.L.ppr_graph_rel_count(self.h))

    def export_csc(self):
        cp, ri = np.zeros(self.n + 1, np.int64), np.zeros(max(self.m, 1), np.int32)
        self.ctx.check(self.L.ppr_graph_export_csc(self.h, _ptr(cp), _ptr(ri)))
        return cp, ri[:self.m]

    # ---- queries
    def forward_push(self, src, alpha, rmax, want_residue=True):
        reserve = np.zeros(self.n)
        residue = np.zeros(self.n) if want_residue else None
        rsum, st = C.c_double(), Stats()
        self.ctx.check(self.L.ppr_forward_push(self.h, src, alpha, rmax, _ptr(reserve), _ptr(residue), C.byref(rsum), C.byref(st)))
        return reserve, residue, rsum.value, st.as_dict()

    def monte_carlo(self, src, alpha, eps, seed, out=None, conf=None):
        out = np.zeros(self.n) if out is None else out
        st = Stats()
        self.ctx.check(self.L.ppr_monte_carlo(self.h, src, alpha, eps, seed, _conf(conf), _ptr(out), C.byref(st)))
        return out, st.as_dict()

    def fora(self, src, alpha, eps, seed, flags=0, out=None, conf=None):
        out = np.zeros(self.n) if out is None else out
        st = Stats()
        self.ctx.check(self.L.ppr_fora(self.h, src, alpha, eps, seed, flags, _conf(conf), _ptr(out), C.byref(st)))
        return out, st.as_dict()

    def fora_shard(self, src, alpha, eps, seed, shard, nshards, flags=0, conf=None):
        st = Stats()
        self.ctx.check(self.L.ppr_fora_shard(self.h, src, alpha, eps, seed, flags, _conf(conf), shard, nshards, C.byref(st)))
        return st.as_dict()

    def monte_carlo_shard(self, src, alpha, eps, seed, shard, nshards, conf=None):
        st = Stats()
        self.ctx.check(self.L.ppr_monte_carlo_shard(self.h, src, alpha, eps, seed, _conf(conf), shard, nshards, C.byref(st)))
        return st.as_dict()

    # ---- collective forms (every rank calls with the same arguments; ppr_comm_init first)
    def fora_sharded(self, src, alpha, eps, seed, flags=0, out=None, conf=None, resident=False):
        out = None if resident else (np.zeros(self.n) if out is None else out)
        st = Stats()
        self.ctx.check(self.L.ppr_fora_sharded(self.h, src, alpha, eps, seed, flags, _conf(conf), _ptr(out), C.byref(st)))
        return out, st.as_dict()

    def monte_carlo_sharded(self, src, alpha, eps, seed, out=None, conf=None, resident=False):
        out = None if resident else (np.zeros(self.n) if out is None else out)
        st = Stats()
        self.ctx.check(self.L.ppr_monte_carlo_sharded(self.h, src, alpha, eps, seed, _conf(conf), _ptr(out), C.byref(st)))
        return out, st.as_dict()

    def fora_topk_sharded(self, src, alpha, eps, k, seed, flags=0, conf=None, cap=None):
        cap = int(min(self.n, max(4 * k, k + 4096))) if cap is None else int(cap)
        while True:
            ids, vals = np.empty(cap, np.int32), np.empty(cap)
            cnt, st = C.c_int32(), Stats()
            rc = self.L.ppr_fora_topk_sharded(self.h, src, alpha, eps, k, seed, flags, _conf(conf), _ptr(ids), _ptr(vals), cap, C.byref(cnt), C.byref(st))
            if rc == PPR_E_CAPACITY and cnt.value > cap:
                cap = cnt.value
                continue
            self.ctx.check(rc)
            return ids[:cnt.value].copy(), vals[:cnt.value].copy(), st.as_dict()

    def backward_all_pairs_sharded(self, alpha, threshold, k=-1, raw=False):
        p, st = C.c_void_p(), Stats()
        self.ctx.check(self.L.ppr_backward_all_pairs_sharded(self.h, alpha, threshold, k, C.byref(p), C.byref(st)))
        return self._pairs_out(p, raw) + (st.as_dict(),)

    def _pairs_out(self, p, raw):
        try:
            cnt = int(self.L.ppr_pairs_count(p))
            ptr, tgt = np.zeros(self.n + 1, np.int64), np.zeros(max(cnt, 1), np.int32)
            val = np.zeros(max(cnt, 1)) if not raw else None
            vfx = np.zeros(max(cnt, 1), np.uint64) if raw else None
            self.ctx.check(self.L.ppr_pairs_export(p, _ptr(ptr), _ptr(tgt), _ptr(val), _ptr(vfx)))
        finally:
            self.L.ppr_pairs_destroy(p)
        return ptr, tgt[:cnt], (vfx[:cnt] if raw else val[:cnt])

    def fora_batch(self, srcs, alpha, eps, seed, flags=0, out=None, want_stats=True, conf=None, resident=False):
        srcs = np.ascontiguousarray(srcs, np.int32)
        nq = srcs.shape[0]
        if not resident:
            out = np.zeros((nq, self.n)) if out is None else out
        st = (Stats * max(nq, 1))()
        self.ctx.check(self.L.ppr_fora_batch(self.h, _ptr(srcs), nq, alpha, eps, seed, flags, _conf(conf), _ptr(out), st if want_stats else None))
        return out, [st[i].as_dict() for i in range(nq)] if want_stats else None

    def fora_topk(self, src, alpha, eps, k, seed, flags=0, conf=None, cap=None):
        # buffers for k + ties (cap: first-call capacity, for tests); a result with more ties than that reports the needed size (PPR_E_CAPACITY) and is
        # fetched by a second call with room for everything
        cap = int(min(self.n, max(4 * k, k + 4096))) if cap is None else int(cap)
        while True:
            ids, vals = np.empty(cap, np.int32), np.empty(cap)
            cnt, st = C.c_int32(), Stats()
            rc = self.L.ppr_fora_topk(self.h, src, alpha, eps, k, seed, flags, _conf(conf), _ptr(ids), _ptr(vals), cap, C.byref(cnt), C.byref(st))
            if rc == PPR_E_CAPACITY and cnt.value > cap:
                cap = cnt.value
                continue
            self.ctx.check(rc)
            return ids[:cnt.value].copy(), vals[:cnt.value].copy(), st.as_dict()

    def fora_topk_shard(self, src, alpha, eps, k, seed, shard, nshards, reduce=None, flags=0, conf=None, cap=None):
        """ppr_fora_topk_shard; `reduce(dev_ptr, n)` sums the round's raw estimate vector over the shards in place.
        Output buffers as in fora_topk (every shard sees the same count, so all of them repeat or none does)."""
        cap = int(min(self.n, max(4 * k, k + 4096))) if cap is None else int(cap)
        failure = []

        def hook(_user, dev_ptr, n):
            try:
                reduce(int(dev_ptr), int(n))
                return 0
            except BaseException as e:   # never let an exception cross the C frame
                failure.append(e)
                return 1

        cb = REDUCE_FN(hook) if reduce is not None else C.cast(None, REDUCE_FN)
        while True:
            ids, vals = np.empty(cap, np.int32), np.empty(cap)
            cnt, st = C.c_int32(), Stats()
            rc = self.L.ppr_fora_topk_shard(self.h, src, alpha, eps, k, seed, flags, _conf(conf), shard, nshards, cb, None,
                                            _ptr(ids), _ptr(vals), cap, C.byref(cnt), C.byref(st))
            if failure:
                raise failure[0]
            if rc == PPR_E_CAPACITY and cnt.value > cap:
                cap = cnt.value
                continue
            self.ctx.check(rc)
            return ids[:cnt.value].copy(), vals[:cnt.value].copy(), st.as_dict()

    def power_iteration(self, src, alpha, iters=100):
        out = np.zeros(self.n)
        self.ctx.check(self.L.ppr_power_iteration(self.h, src, alpha, iters, _ptr(out)))
        return out

    def backward_search(self, target, alpha, rmax):
        out, st = np.zeros(self.n), Stats()
        self.ctx.check(self.L.ppr_backward_search(self.h, target, alpha, rmax, _ptr(out), C.byref(st)))
        return out, st.as_dict()

    def backward_all_pairs(self, alpha, threshold, k=-1, t_lo=0, t_hi=None, raw=False):
        t_hi = self.n if t_hi is None else t_hi
        p, st = C.c_void_p(), Stats()
        self.ctx.check(self.L.ppr_backward_all_pairs(self.h, alpha, threshold, k, t_lo, t_hi, C.byref(p), C.byref(st)))
        return self._pairs_out(p, raw) + (st.as_dict(),)

    # ---- Gen_Util.computeError on the device (J/Gen_Util.java:259-326)
    def metrics_keep(self, slot):
        """Snapshot the last query's dense result in HBM: slot 0 = ground truth (power iteration), 1 = estimate."""
        self.ctx.check(self.L.ppr_metrics_keep(self.h, slot))

    def metrics_compute(self, k):
        me, pr, nd = C.c_double(), C.c_double(), C.c_double()
        self.ctx.check(self.L.ppr_metrics_compute(self.h, k, C.byref(me), C.byref(pr), C.byref(nd)))
        return {"max_err": me.value, "precision": pr.value, "ndcg": nd.value}

    # ---- raw state / replay hooks (parity tests, multi-GPU plumbing)
    def last_raw(self, want_residue=True):
        reserve = np.zeros(self.n, np.uint64)
        residue = np.zeros(self.n, np.uint64) if want_residue else None
        self.ctx.check(self.L.ppr_last_raw(self.h, _ptr(reserve), _ptr(residue)))
        return reserve, residue

    def result_devptr(self):
        p, n = C.c_uint64(), C.c_int64()
        self.ctx.check(self.L.ppr_result_devptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def result_export(self, divisor=0.0, out=None):
        out = np.zeros(self.n) if out is None else out
        self.ctx.check(self.L.ppr_result_export(self.h, divisor, _ptr(out)))
        return out

    def push_trace(self, src, alpha, rmax, max_iters=100000):
        """Frontier sets of the synchronous push, one sorted id array per iteration."""
        self.ctx.check(self.L.ppr_push_trace_begin(self.h, src, alpha, rmax))
        buf = np.zeros(self.n, np.int32)
        out = []
        for _ in range(max_iters):
            cnt = C.c_int64()
            self.ctx.check(self.L.ppr_push_trace_step(self.h, _ptr(buf), self.n, C.byref(cnt)))
            if cnt.value == 0:
                break
            out.append(np.sort(buf[:cnt.value]))
        return out

    def debug_timeline(self):
        buf = np.zeros((4096, 6), np.uint64)
        cnt = C.c_int64()
        self.ctx.check(self.L.ppr_debug_timeline(self.h, _ptr(buf), 4096, C.byref(cnt)))
        return buf[:cnt.value]

    def walk_terminals(self, start, restart, j0, count, seed, tag, alpha, no_zero_hop):
        out = np.zeros(max(count, 1), np.int32)
        self.ctx.check(self.L.ppr_walk_terminals(self.h, start, restart, j0, count, seed, tag, alpha, int(no_zero_hop), _ptr(out)))
        return out[:count]

    def close(self):
        if self.h and self.ctx.h:
            self.L.ppr_graph_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
