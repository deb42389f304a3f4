// select.cuh -- result export, tie-inclusive top-k (MSB radix select) and the power iteration.
//
// Top-k replaces Algo_Util.kth_ppr (randomised quick-select on boxed Doubles, J/Algo_Util.java:32-73)
// and the retrieveTopK filters (J/Fora_Topk.java:186-199, J/Forward_Push.java:413-429): find the k-th
// largest positive value, keep everything >= it.  Values are the raw fixed-point uint64, whose integer
// order is the numeric order, so an 8-bit-digit MSB radix select needs at most 8 histogram passes over
// the (L2-resident) vector and no sort.  The <= k+ties survivors are ordered on the host.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
namespace cg = cooperative_groups;

#define SEL_THREADS 256

// raw fixed point / counts -> double.  divisor == 0: value * 2^-62; else value / divisor
// (Monte-Carlo: count / omega, J/Monte_Carlo.java:156-157).
__global__ void k_export_f64(const u64* __restrict__ src, double* __restrict__ dst, i64 n, double divisor) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 v = gtid; v < n; v += gsize) {
        double x = (double)src[v];
        dst[v] = divisor == 0.0 ? __dmul_rn(x, FX_INV) : __ddiv_rn(x, divisor);
    }
}

__global__ void k_import_f64(const double* __restrict__ src, u64* __restrict__ dst, i64 n) {
    // order-preserving image of non-negative doubles (IEEE bit pattern); negatives/NaN map to 0
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 v = gtid; v < n; v += gsize) {
        double x = src[v];
        dst[v] = x > 0.0 ? (u64)__double_as_longlong(x) : 0ull;
    }
}

struct SelArgs {
    const u64* vals;
    i64 n;
    i64 k;
    u32* hist;       // [9][256]: slot 0 = positive count, slots 1..8 = per-pass digit histograms
    i32* sel_ids;
    u64* sel_vals;
    i64 sel_cap;
    PushCtl* ctl;
};

__global__ void __launch_bounds__(SEL_THREADS, 4) k_topk_select(SelArgs A) {
    cg::grid_group grid = cg::this_grid();
    __shared__ u32 sh[256];
    __shared__ u64 s_prefix;
    __shared__ i64 s_krem;
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    PushCtl* ctl = A.ctl;

    for (i64 i = gtid; i < 9 * 256; i += gsize) A.hist[i] = 0;
    if (gtid == 0) { ctl->sel_count = 0; ctl->sel_pos = 0; ctl->sel_kth = 0; }
    grid.sync();

    // positives
    {
        u32 c = 0;
        for (i64 v = gtid; v < A.n; v += gsize) c += A.vals[v] != 0;
        c = (u32)warp_sum_u64(c);
        if (lane_id() == 0 && c) atomicAdd(&A.hist[0], c);
    }
    grid.sync();
    const u32 positives = *(volatile u32*)&A.hist[0];
    u64 kth = 0;
    if (A.k > 0 && (i64)positives >= A.k) {
        u64 prefix = 0;
        i64 krem = A.k;
        for (int pass = 0; pass < 8; pass++) {
            const int shift = 56 - 8 * pass;
            u32* gh = A.hist + 256 * (pass + 1);
            sh[threadIdx.x] = 0;
            __syncthreads();
            const u64 himask = pass == 0 ? 0ull : (~0ull << (shift + 8));
            for (i64 v = gtid; v < A.n; v += gsize) {
                u64 x = A.vals[v];
                if (x && (x & himask) == prefix) atomicAdd(&sh[(x >> shift) & 0xff], 1u);
            }
            __syncthreads();
            if (sh[threadIdx.x]) atomicAdd(&gh[threadIdx.x], sh[threadIdx.x]);
            grid.sync();
            // every CTA picks the same bucket from the global histogram
            sh[threadIdx.x] = *(volatile u32*)&gh[threadIdx.x];
            __syncthreads();
            if (threadIdx.x == 0) {
                i64 acc = 0;
                int b = 255;
                for (; b > 0; b--) { if (acc + (i64)sh[b] >= krem) break; acc += sh[b]; }
                s_prefix = prefix | ((u64)b << shift);
                s_krem = krem - acc;
            }
            __syncthreads();
            prefix = s_prefix;
            krem = s_krem;
            __syncthreads();
        }
        kth = prefix;
    }
    if (gtid == 0) { ctl->sel_kth = kth; ctl->sel_pos = positives; }
    // survivors: positive and >= kth (kth == 0 keeps every positive entry: fewer than k present,
    // the "kth_ppr returned null" branch of J/Fora_Topk.java:188-191)
    for (i64 base = 0; base < A.n; base += gsize) {
        i64 v = base + gtid;
        u64 x = v < A.n ? A.vals[v] : 0;
        bool keep = x != 0 && x >= kth;
        u32 mask = __ballot_sync(0xffffffffu, keep);
        if (mask) {
            u32 leader = __ffs(mask) - 1, b = 0;
            if (lane_id() == leader) b = atomicAdd(&ctl->sel_count, (u32)__popc(mask));
            b = __shfl_sync(0xffffffffu, b, leader);
            u32 slot = b + __popc(mask & lanemask_lt());
            if (keep && (i64)slot < A.sel_cap) { A.sel_ids[slot] = (i32)v; A.sel_vals[slot] = x; }
        }
    }
}

// ------------------------------------------------------------------------------------------ power iteration
// Power_Method.computeWholeGraphPPR, J/Power_Method.java:57-100: pi += alpha r; r' = (1-alpha) P^T r with
// dangling mass returned to the source.  Pull formulation over the CSC: no atomics, fixed summation
// order (ascending source inside a column), so the result is reproducible run to run.
__global__ void k_pm_scale(const double* __restrict__ r, double* __restrict__ q, double* __restrict__ reserve,
                           const u32* __restrict__ out_deg, i64 n, double alpha, double* __restrict__ blk_dang) {
    __shared__ double sm[256];
    double dang = 0.0;
    const i64 per = (n + gridDim.x - 1) / gridDim.x;
    const i64 lo = min(n, (i64)blockIdx.x * per), hi = min(n, lo + per);
    for (i64 v = lo + threadIdx.x; v < hi; v += blockDim.x) {
        double rc = r[v];
        double qq = 0.0;
        if (rc > 0) {
            reserve[v] = __dadd_rn(reserve[v], __dmul_rn(rc, alpha));
            double remain = __dmul_rn(rc, 1.0 - alpha);
            u32 d = out_deg[v];
            if (d == 0) dang = __dadd_rn(dang, remain);
            else qq = __ddiv_rn(remain, (double)d);
        }
        q[v] = qq;
    }
    sm[threadIdx.x] = dang;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sm[threadIdx.x] = __dadd_rn(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) blk_dang[blockIdx.x] = sm[0];
}
__global__ void k_pm_pull(const double* __restrict__ q, double* __restrict__ r_new, const u64* __restrict__ in_vtx,
                          const i32* __restrict__ row_idx, i64 n, i32 src, const double* __restrict__ blk_dang, int nblk) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 u = gtid; u < n; u += gsize) {
        u64 pk = in_vtx[u];
        u64 b = vtx_begin(pk);
        u32 d = vtx_deg(pk);
        double s = 0.0;
        for (u32 e = 0; e < d; e++) s = __dadd_rn(s, q[row_idx[b + e]]);
        if (u == (i64)src) {
            double dg = 0.0;
            for (int k = 0; k < nblk; k++) dg = __dadd_rn(dg, blk_dang[k]);
            s = __dadd_rn(s, dg);
        }
        r_new[u] = s;
    }
}
