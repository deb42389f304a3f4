"""ctypes front-end of the CPU oracle (oracle/ppr_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs.  The product package never imports this module.  Parity is unpinned by the reference (it has
no tests and cannot run here) -- see the header of ppr_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORC_REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)   # int (*)(void* user, uint64* est, int64 n)
_LIB = None

FX_ONE = 1 << 62
FX_INV = 1.0 / float(1 << 62)

_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libppr_oracle.so")
    src = os.path.join(_HERE, "ppr_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libppr_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        G = [C.c_int64, C.c_int64, _i64p, _i32p]
        L.orc_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.orc_conf.argtypes = [C.c_int, C.c_int64, C.c_int64, _f64p]
        L.orc_fora_params.argtypes = [C.c_int64, C.c_int64, C.c_double, C.c_double, _f64p]
        L.orc_power_method.argtypes = G + [C.c_int32, C.c_double, C.c_int, C.c_double, _f64p]
        L.orc_power_method.restype = C.c_int
        L.orc_walk.argtypes = G + [C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, C.c_uint64, C.c_double, C.c_int, _i64p]
        L.orc_walk.restype = C.c_int32
        L.orc_a_forward_push.argtypes = G + [C.c_int32, C.c_double, C.c_double, _f64p, _f64p, _f64p, _i64p]
        L.orc_a_monte_carlo.argtypes = G + [C.c_int32, C.c_double, C.c_double, C.c_uint64, _f64p, _i64p]
        L.orc_a_monte_carlo.restype = C.c_int64
        L.orc_a_fora.argtypes = G + [C.c_int32, C.c_double, C.c_double, C.c_uint64, C.c_double, C.c_int, C.c_int, _f64p, _i64p, _f64p]
        L.orc_a_fora_topk.argtypes = G + [C.c_int32, C.c_double, C.c_double, C.c_int64, C.c_uint64, C.c_int, _f64p, _i64p]
        L.orc_a_backward_search.argtypes = G + [C.c_int32, C.c_double, C.c_double, _f64p, _i64p]
        L.orc_a_base_all_pairs.argtypes = G + [C.c_double, C.c_double, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _i64p]
        L.orc_a_base_all_pairs.restype = C.c_int64
        L.orc_kth_ppr.argtypes = [_f64p, C.c_int64, C.c_int64, _f64p]
        L.orc_kth_ppr.restype = C.c_int
        L.orc_topk_select_f64.argtypes = [_f64p, C.c_int64, C.c_int64, _i32p, C.c_int64]
        L.orc_topk_select_f64.restype = C.c_int64
        L.orc_topk_select_u64.argtypes = [_u64p, C.c_int64, C.c_int64, _i32p, C.c_int64]
        L.orc_topk_select_u64.restype = C.c_int64
        L.orc_topk_metrics.argtypes = [_f64p, C.c_int64, C.c_int64, _i32p, C.c_int64, _f64p]
        L.orc_max_err.argtypes = [_f64p, _f64p, C.c_int64]
        L.orc_max_err.restype = C.c_double
        L.orc_b_forward_push.argtypes = G + [C.c_int32, C.c_double, C.c_double, _u64p, _u64p, _i64p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, _i64p]
        L.orc_b_monte_carlo.argtypes = G + [C.c_int32, C.c_double, C.c_double, C.c_uint64, C.c_int64, C.c_int64, _u64p, _f64p, _i64p]
        L.orc_b_monte_carlo.restype = C.c_int64
        L.orc_b_fora.argtypes = G + [C.c_int32, C.c_double, C.c_double, C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_int64, C.c_int64, _u64p, _i64p]
        L.orc_b_fora_topk.argtypes = G + [C.c_int32, C.c_double, C.c_double, C.c_int64, C.c_uint64, C.c_int, _u64p, _i64p]
        L.orc_b_fora_topk_shard.argtypes = G + [C.c_int32, C.c_double, C.c_double, C.c_int64, C.c_uint64, C.c_int, C.c_int64, C.c_int64, ORC_REDUCE_FN, C.c_void_p, _u64p, _i64p]
        L.orc_b_backward_search.argtypes = G + [C.c_int32, C.c_double, C.c_double, _u64p, _u64p, _i64p]
        L.orc_b_base_all_pairs.argtypes = G + [C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _i64p]
        L.orc_b_base_all_pairs.restype = C.c_int64
    return _LIB


class OGraph:
    """CSR view handed to every oracle call (row_ptr int64[n+1], col_idx int32[m])."""

    def __init__(self, row_ptr, col_idx):
        self.rp = np.ascontiguousarray(row_ptr, dtype=np.int64)
        self.ci = np.ascontiguousarray(col_idx, dtype=np.int32)
        self.n = int(self.rp.shape[0] - 1)
        self.m = int(self.ci.shape[0])

    def args(self):
        return (self.n, self.m, self.rp, self.ci)


def philox(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().orc_philox4x32_10(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
    return out


def conf(n, k=None):
    out = np.zeros(3)
    lib().orc_conf(0 if k is None else 1, n, 0 if k is None else k, out)
    return {"delta": out[0], "pfail": out[1], "min_delta": out[2]}


def fora_params(n, m, alpha, eps):
    out = np.zeros(2)
    lib().orc_fora_params(n, m, alpha, eps, out)
    return {"rmax": out[0], "omega": out[1]}


def power_method(g, s, alpha, iters=100, tol=0.0):
    out = np.zeros(g.n)
    lib().orc_power_method(*g.args(), s, alpha, iters, tol, out)
    return out


def walk(g, start, restart, j, seed, tag, alpha, no_zero_hop):
    steps = np.zeros(1, np.int64)
    t = lib().orc_walk(*g.args(), start, restart, j, seed, tag, alpha, 1 if no_zero_hop else 0, steps)
    return int(t), int(steps[0])


# ---------------------------------------------------------------- mode A (reference order, double)
def a_forward_push(g, s, alpha, rmax):
    reserve, residue, rsum, st = np.zeros(g.n), np.zeros(g.n), np.zeros(1), np.zeros(2, np.int64)
    lib().orc_a_forward_push(*g.args(), s, alpha, rmax, reserve, residue, rsum, st)
    return reserve, residue, float(rsum[0]), {"pushes": int(st[0]), "edges": int(st[1])}


def a_monte_carlo(g, s, alpha, eps, seed):
    ppr, steps = np.zeros(g.n), np.zeros(1, np.int64)
    w = lib().orc_a_monte_carlo(*g.args(), s, alpha, eps, seed, ppr, steps)
    return ppr, {"walks": int(w), "steps": int(steps[0])}


def a_fora(g, s, alpha, eps, seed, walk_cost=32.0, max_rounds=64, restart_at_source=False):
    out, st, ds = np.zeros(g.n), np.zeros(5, np.int64), np.zeros(2)
    lib().orc_a_fora(*g.args(), s, alpha, eps, seed, walk_cost, max_rounds, int(restart_at_source), out, st, ds)
    return out, dict(zip(["pushes", "edges", "rounds", "walks", "steps"], map(int, st)), rmax_final=ds[0], rsum=ds[1])


def a_fora_topk(g, s, alpha, eps, k, seed, restart_at_source=False):
    out, st = np.zeros(g.n), np.zeros(5, np.int64)
    lib().orc_a_fora_topk(*g.args(), s, alpha, eps, k, seed, int(restart_at_source), out, st)
    return out, dict(zip(["pushes", "edges", "rounds", "walks", "steps"], map(int, st)))


def a_backward_search(g, t, alpha, rmax):
    out, st = np.zeros(g.n), np.zeros(2, np.int64)
    lib().orc_a_backward_search(*g.args(), t, alpha, rmax, out, st)
    return out, {"pushes": int(st[0]), "edges": int(st[1])}


def a_base_all_pairs(g, alpha, threshold, k=-1, t_lo=0, t_hi=None):
    t_hi = g.n if t_hi is None else t_hi
    st = np.zeros(2, np.int64)
    total = lib().orc_a_base_all_pairs(*g.args(), alpha, threshold, k, t_lo, t_hi, None, None, None, 0, st)
    ptr, tgt, val = np.zeros(g.n + 1, np.int64), np.zeros(max(total, 1), np.int32), np.zeros(max(total, 1))
    lib().orc_a_base_all_pairs(*g.args(), alpha, threshold, k, t_lo, t_hi, ptr.ctypes.data, tgt.ctypes.data,
                               val.ctypes.data, total, st)
    return ptr, tgt[:total], val[:total], {"pushes": int(st[0]), "edges": int(st[1])}


# ---------------------------------------------------------------- selection and metrics
def kth_ppr(vals, k):
    out = np.zeros(1)
    ok = lib().orc_kth_ppr(np.ascontiguousarray(vals, np.float64), len(vals), k, out)
    return float(out[0]) if ok else None


def topk_select(vals, k):
    vals = np.ascontiguousarray(vals)
    ids = np.zeros(len(vals), np.int32)
    if vals.dtype == np.uint64:
        c = lib().orc_topk_select_u64(vals, len(vals), k, ids, len(ids))
    else:
        c = lib().orc_topk_select_f64(vals.astype(np.float64), len(vals), k, ids, len(ids))
    return ids[:c].copy()


def topk_metrics(gnd, k, algo_ids):
    out = np.zeros(2)
    algo_ids = np.ascontiguousarray(algo_ids, np.int32)
    lib().orc_topk_metrics(np.ascontiguousarray(gnd, np.float64), len(gnd), k, algo_ids, len(algo_ids), out)
    return {"precision": out[0], "ndcg": out[1]}


def max_err(gnd, est):
    return lib().orc_max_err(np.ascontiguousarray(gnd, np.float64), np.ascontiguousarray(est, np.float64), len(gnd))


# ---------------------------------------------------------------- mode B (synchronous, fixed point)
_B_STATS = ["pushes", "edges", "iters", "rounds", "walks", "steps", "walk_sources", "rmax_fx_final", "rsum_fx"]


def b_forward_push(g, s, alpha, rmax, trace=False):
    reserve, residue = np.zeros(g.n, np.uint64), np.zeros(g.n, np.uint64)
    st, ts = np.zeros(3, np.int64), np.zeros(2, np.int64)
    lib().orc_b_forward_push(*g.args(), s, alpha, rmax, reserve, residue, st, None, 0, None, 0, ts)
    out = {"pushes": int(st[0]), "edges": int(st[1]), "iters": int(st[2])}
    if trace:
        ip, vs = np.zeros(ts[0] + 1, np.int64), np.zeros(max(int(ts[1]), 1), np.int32)
        lib().orc_b_forward_push(*g.args(), s, alpha, rmax, reserve, residue, st, ip.ctypes.data, int(ts[0]),
                                 vs.ctypes.data, int(ts[1]), ts)
        out["frontiers"] = [vs[ip[i]:ip[i + 1]].copy() for i in range(int(ts[0]))]
    return reserve, residue, out


def b_monte_carlo(g, s, alpha, eps, seed, shard=0, nshards=1):
    counts, om, steps = np.zeros(g.n, np.uint64), np.zeros(1), np.zeros(1, np.int64)
    w = lib().orc_b_monte_carlo(*g.args(), s, alpha, eps, seed, shard, nshards, counts, om, steps)
    return counts, {"walks": int(w), "omega": float(om[0]), "steps": int(steps[0])}


def b_fora(g, s, alpha, eps, seed, walk_cost=6.0, max_rounds=64, restart_at_source=False, shard=0, nshards=1):
    out, st = np.zeros(g.n, np.uint64), np.zeros(9, np.int64)
    lib().orc_b_fora(*g.args(), s, alpha, eps, seed, walk_cost, max_rounds, int(restart_at_source), shard, nshards, out, st)
    return out, dict(zip(_B_STATS, map(int, st)))


def b_fora_topk(g, s, alpha, eps, k, seed, restart_at_source=False, shard=0, nshards=1, reduce=None):
    """FORA top-k estimate vector (fixed point).  Sharded form: `reduce(est)` must sum the uint64 numpy vector `est`
    over the shards IN PLACE; it is called once per delta-round (the spec of ppr_fora_topk_shard)."""
    out, st = np.zeros(g.n, np.uint64), np.zeros(9, np.int64)
    if nshards == 1 and reduce is None:
        lib().orc_b_fora_topk(*g.args(), s, alpha, eps, k, seed, int(restart_at_source), out, st)
        return out, dict(zip(_B_STATS, map(int, st)))

    def hook(_user, ptr, n):
        reduce(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint64)), shape=(n,)))
        return 0

    cb = ORC_REDUCE_FN(hook)
    lib().orc_b_fora_topk_shard(*g.args(), s, alpha, eps, k, seed, int(restart_at_source), shard, nshards, cb, None, out, st)
    return out, dict(zip(_B_STATS, map(int, st)))


def b_backward_search(g, t, alpha, rmax):
    reserve, residue, st = np.zeros(g.n, np.uint64), np.zeros(g.n, np.uint64), np.zeros(3, np.int64)
    lib().orc_b_backward_search(*g.args(), t, alpha, rmax, reserve, residue, st)
    return reserve, residue, {"pushes": int(st[0]), "edges": int(st[1]), "iters": int(st[2])}


def b_base_all_pairs(g, alpha, threshold, t_lo=0, t_hi=None):
    t_hi = g.n if t_hi is None else t_hi
    st = np.zeros(3, np.int64)
    total = lib().orc_b_base_all_pairs(*g.args(), alpha, threshold, t_lo, t_hi, None, None, None, 0, st)
    ptr, tgt, val = np.zeros(g.n + 1, np.int64), np.zeros(max(total, 1), np.int32), np.zeros(max(total, 1), np.uint64)
    lib().orc_b_base_all_pairs(*g.args(), alpha, threshold, t_lo, t_hi, ptr.ctypes.data, tgt.ctypes.data,
                               val.ctypes.data, total, st)
    return ptr, tgt[:total], val[:total], {"pushes": int(st[0]), "edges": int(st[1]), "iters": int(st[2])}
