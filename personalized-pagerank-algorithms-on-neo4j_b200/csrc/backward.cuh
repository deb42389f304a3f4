// backward.cuh -- target-batched reverse push over the CSC (All-Pair-Backward-Search) for sm_100a.
//
// Replaces Backward_Search.backward_search_whole_graph (J/Backward_Search.java:38-100) run once per
// target by Base_Whole_Graph.preprocessing (J/Base_Whole_Graph.java:76-92).  The n searches are
// independent, so a batch of T targets is advanced together: dense per-target state r[T][n],
// reserve[T][n] (fixed point), one frontier of (target slot, vertex) items, one synchronous iteration
// for all T searches per pair of phases.  Frontier discovery is the same exactly-once crossing test as
// the forward push, with the reference's strict absolute threshold r > rmax (J/Backward_Search.java:89).
//   phase A: x = r[t][v]; r[t][v] = 0; reserve[t][v] += mulhi(x, alpha); y = x - that
//   phase B: every in-neighbour u of v: r[t][u] += mulhi(y, invdeg[u]),  invdeg[u] = (2^64-1)/d_out(u)
//            (J/Backward_Search.java:85: (1-alpha) r(v) / |N_out(u)|)
// After convergence the same kernel compacts {(v, t, reserve) : reserve >= threshold} in (t, v) order
// (J/Base_Whole_Graph.java:80-88).  A target without in-edges yields {t: 1} (J/Backward_Search.java:46-49).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "push.cuh"
namespace cg = cooperative_groups;

struct BwdArgs {
    i64 n;
    const u64* in_vtx;
    const i32* row_idx;
    const u64* invdeg;
    u64* r;        // [T][n]
    u64* reserve;  // [T][n]
    u64* front[2]; // items (slot << 32 | v)
    i64 front_cap;
    u64* s_pack; u64* s_y; u32* s_t;
    u64* l_pack; u64* l_y; u32* l_t;
    uint2* chunks; i64 chunk_cap;
    u64* blk_a;
    PushCtl* ctl;
    i32 t0;        // first target of the batch
    i32 T;         // targets in the batch
    u64 alpha_q64;
    u64 rmax_fx;   // largest FX value <= rmax
    u64 thr_fx;    // smallest FX value >= threshold
    // output triples
    u32* out_v; u32* out_t; u64* out_val; i64 out_cap;
};

__device__ __forceinline__ void warp_append_u64(u64* list, u32* counter, i64 cap, u32* err, bool pred, u64 value) {
    u32 mask = __ballot_sync(0xffffffffu, pred);
    if (mask) {
        u32 leader = __ffs(mask) - 1, base = 0;
        if (lane_id() == leader) base = atomicAdd(counter, (u32)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        u32 slot = base + __popc(mask & lanemask_lt());
        if (pred) { if ((i64)slot < cap) list[slot] = value; else *err = 1; }
    }
}

__device__ __forceinline__ bool relax_in_edge(const BwdArgs& A, u64* rt, u32 u, u64 y) {
    u64 inc = mulhi64(y, __ldg(&A.invdeg[u]));
    u64 old = atomicAdd((ull*)&rt[u], (ull)inc);
    return old <= A.rmax_fx && old + inc > A.rmax_fx;
}

__global__ void __launch_bounds__(PUSH_THREADS, 4) k_bwd_batch(BwdArgs A) {
    cg::grid_group grid = cg::this_grid();
    __shared__ u64 sm_excl[(PUSH_THREADS / 32) * 32];
    __shared__ u64 sm_beg[(PUSH_THREADS / 32) * 32];
    __shared__ u64 sm_y[(PUSH_THREADS / 32) * 32];
    __shared__ u32 sm_t[(PUSH_THREADS / 32) * 32];
    PushCtl* ctl = A.ctl;
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    const i64 gwarp = gtid >> 5, nwarps = gsize >> 5;
    const u32 lane = lane_id();
    const int wib = threadIdx.x >> 5;
    const i64 n = A.n;

    // seed: state was zeroed by the host (memset); one item per target with in-edges
    if (gtid == 0) {
        ctl->cnt[0] = ctl->cnt[1] = 0; ctl->ns[0] = ctl->ns[1] = 0; ctl->nl[0] = ctl->nl[1] = 0;
        ctl->nchunks[0] = ctl->nchunks[1] = 0; ctl->pushes = 0; ctl->edges = 0; ctl->iters = 0; ctl->err = 0; ctl->err_a = 0;
        ctl->pairs = 0; ctl->pairs_needed = 0;
    }
    grid.sync();
    for (i64 base = 0; base < A.T; base += gsize) {
        i64 s = base + gtid;
        bool pred = false;
        u64 item = 0;
        if (s < A.T) {
            u32 t = (u32)(A.t0 + s);
            if (vtx_deg(A.in_vtx[t]) == 0) A.reserve[s * n + t] = FX_ONE;
            else { A.r[s * n + t] = FX_ONE; pred = true; item = ((u64)s << 32) | t; }
        }
        warp_append_u64(A.front[0], &ctl->cnt[0], A.front_cap, &ctl->err, pred, item);
    }
    grid.sync();

    int par = 0;
    for (;;) {
        const u32 cnt = *(volatile u32*)&ctl->cnt[par];
        if (cnt == 0 || *(volatile u32*)&ctl->err) break;
        // ---------------- phase A
        if (gtid == 0) {
            ctl->cnt[par ^ 1] = 0; ctl->ns[par ^ 1] = 0; ctl->nl[par ^ 1] = 0; ctl->nchunks[par ^ 1] = 0;
            ctl->pushes += cnt; ctl->iters += 1;
        }
        u64 edges_local = 0;
        for (i64 base = 0; base < cnt; base += gsize) {
            i64 i = base + gtid;
            bool is_small = false, is_large = false;
            u64 pk = 0, y = 0; u32 slot_t = 0;
            if (i < cnt) {
                u64 item = A.front[par][i];
                slot_t = (u32)(item >> 32);
                u32 v = (u32)item;
                u64* rt = A.r + (i64)slot_t * n;
                u64 x = rt[v];
                rt[v] = 0;
                u64 a = mulhi64(x, A.alpha_q64);
                A.reserve[(i64)slot_t * n + v] += a;
                y = x - a;
                pk = A.in_vtx[v];
                u32 d = vtx_deg(pk);
                edges_local += d;
                is_small = d > 0 && d <= PUSH_SMALL_MAX;
                is_large = d > PUSH_SMALL_MAX;
            }
            u32 mask = __ballot_sync(0xffffffffu, is_small);
            if (mask) {
                u32 leader = __ffs(mask) - 1, b = 0;
                if (lane == leader) b = atomicAdd(&ctl->ns[par], (u32)__popc(mask));
                b = __shfl_sync(0xffffffffu, b, leader);
                if (is_small) { u32 s = b + __popc(mask & lanemask_lt()); A.s_pack[s] = pk; A.s_y[s] = y; A.s_t[s] = slot_t; }
            }
            if (is_large) {
                u32 s = atomicAdd(&ctl->nl[par], 1u);
                A.l_pack[s] = pk; A.l_y[s] = y; A.l_t[s] = slot_t;
                u32 d = vtx_deg(pk);
                u32 nch = (d + PUSH_CHUNK - 1) / PUSH_CHUNK;
                u32 cb = atomicAdd(&ctl->nchunks[par], nch);
                if ((i64)cb + nch <= A.chunk_cap) { for (u32 c = 0; c < nch; c++) A.chunks[cb + c] = make_uint2(s, c); }
                else ctl->err_a = 1;
            }
        }
        edges_local = warp_sum_u64(edges_local);
        if (lane == 0 && edges_local) atomicAdd((ull*)&ctl->edges, (ull)edges_local);
        grid.sync();
        if (*(volatile u32*)&ctl->err_a) break;
        // ---------------- phase B
        const u32 ns = *(volatile u32*)&ctl->ns[par];
        const u32 nch = *(volatile u32*)&ctl->nchunks[par];
        u64* next = A.front[par ^ 1];
        u32* next_cnt = &ctl->cnt[par ^ 1];
        u64* w_excl = sm_excl + wib * 32; u64* w_beg = sm_beg + wib * 32; u64* w_y = sm_y + wib * 32; u32* w_t = sm_t + wib * 32;
        for (i64 wbase = gwarp * 32; wbase < ns; wbase += nwarps * 32) {
            i64 item = wbase + lane;
            u64 pk = 0, y = 0; u32 st = 0;
            if (item < ns) { pk = A.s_pack[item]; y = A.s_y[item]; st = A.s_t[item]; }
            u32 d = vtx_deg(pk), incl = d;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            u32 excl = incl - d;
            u32 total = __shfl_sync(0xffffffffu, incl, 31);
            w_excl[lane] = excl; w_beg[lane] = vtx_begin(pk) - excl; w_y[lane] = y; w_t[lane] = st;
            __syncwarp();
            for (u32 eb = 0; eb < total; eb += 32) {
                u32 e = eb + lane;
                bool cross = false;
                u64 nitem = 0;
                if (e < total) {
                    int o = 0;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) { int p = o + s; if ((u32)w_excl[p] <= e) o = p; }
                    u32 u = (u32)__ldg(&A.row_idx[w_beg[o] + e]);
                    u32 t = w_t[o];
                    cross = relax_in_edge(A, A.r + (i64)t * n, u, w_y[o]);
                    nitem = ((u64)t << 32) | u;
                }
                warp_append_u64(next, next_cnt, A.front_cap, &ctl->err, cross, nitem);
            }
            __syncwarp();
        }
        for (i64 c = gwarp; c < nch; c += nwarps) {
            uint2 desc = A.chunks[c];
            u64 pk = A.l_pack[desc.x];
            u64 y = A.l_y[desc.x];
            u32 t = A.l_t[desc.x];
            u32 d = vtx_deg(pk), off = desc.y * PUSH_CHUNK;
            u32 len = min((u32)PUSH_CHUNK, d - off);
            const i32* row = A.row_idx + vtx_begin(pk) + off;
            u64* rt = A.r + (i64)t * n;
            for (u32 eb = 0; eb < len; eb += 32) {
                u32 e = eb + lane;
                bool cross = false;
                u64 nitem = 0;
                if (e < len) {
                    u32 u = (u32)__ldg(&row[e]);
                    cross = relax_in_edge(A, rt, u, y);
                    nitem = ((u64)t << 32) | u;
                }
                warp_append_u64(next, next_cnt, A.front_cap, &ctl->err, cross, nitem);
            }
        }
        grid.sync();
        par ^= 1;
    }
    if (*(volatile u32*)&ctl->err || *(volatile u32*)&ctl->err_a) return;

    // ---------------- ordered compaction of {reserve >= thr_fx} over the [T][n] matrix
    const i64 total = (i64)A.T * n;
    const i64 per = (total + gridDim.x - 1) / gridDim.x;
    const i64 lo = min(total, (i64)blockIdx.x * per), hi = min(total, lo + per);
    u64 c = 0;
    for (i64 i = lo + threadIdx.x; i < hi; i += blockDim.x) { u64 x = A.reserve[i]; c += (x != 0 && x >= A.thr_fx); }
    c = block_sum_u64(c, sm_excl);
    if (threadIdx.x == 0) A.blk_a[blockIdx.x] = c;
    grid.sync();
    u64 off = 0;
    for (u32 b = threadIdx.x; b < blockIdx.x; b += blockDim.x) off += A.blk_a[b];
    off = block_sum_u64(off, sm_excl);
    for (i64 tile = lo; tile < hi; tile += blockDim.x) {
        i64 i = tile + threadIdx.x;
        u64 x = i < hi ? A.reserve[i] : 0;
        u64 has = (x != 0 && x >= A.thr_fx) ? 1 : 0;
        u64 e1, e2, t1, t2;
        block_excl_scan2(has, 0, sm_excl, e1, e2, t1, t2);
        if (has) {
            u64 slot = off + e1;
            if ((i64)slot < A.out_cap) { A.out_v[slot] = (u32)(i % n); A.out_t[slot] = (u32)(A.t0 + i / n); A.out_val[slot] = x; }
        }
        off += t1;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        ctl->pairs_needed = off;
        ctl->pairs = min(off, (u64)A.out_cap);
    }
}
