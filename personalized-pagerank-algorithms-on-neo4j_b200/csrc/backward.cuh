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
    u32* wcnt;     // per-warp counts of the output compaction
    u64* dbg;      // per-iteration timeline (same layout as the forward push)
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

// four predicated items per lane, one counter round trip per warp
__device__ __forceinline__ void warp_append4_u64(u64* list, u32* counter, i64 cap, u32* err, const bool (&pred)[4], const u64 (&value)[4]) {
    u32 m[4], tot = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { m[k] = __ballot_sync(0xffffffffu, pred[k]); tot += __popc(m[k]); }
    if (tot) {
        u32 base = 0;
        if (lane_id() == 0) base = atomicAdd(counter, tot);
        base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (pred[k]) { u32 slot = base + __popc(m[k] & lanemask_lt()); if ((i64)slot < cap) list[slot] = value[k]; else *err = 1; }
            base += __popc(m[k]);
        }
    }
}

__device__ __forceinline__ bool relax_in_edge(const BwdArgs& A, u64* rt, u32 u, u64 y) {
    u64 inc = mulhi64(y, __ldg(&A.invdeg[u]));
    u64 old = atomicAdd((ull*)&rt[u], (ull)inc);
    return old <= A.rmax_fx && old + inc > A.rmax_fx;
}

__global__ void __launch_bounds__(PUSH_THREADS, PUSH_CTAS_PER_SM) k_bwd_batch(BwdArgs A) {
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
        ctl->cnt[0] = ctl->cnt[1] = 0; ctl->stage[0] = ctl->stage[1] = 0;
        ctl->pushes = 0; ctl->edges = 0; ctl->iters = 0; ctl->err = 0; ctl->err_a = 0;
        ctl->pairs = 0; ctl->pairs_needed = 0; ctl->dbg_count = 0;
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
        // ---------------- phase A: warp-granular tiles of 128 items, one packed staging atomic per tile
        u64 tt0 = 0, tt1 = 0;
        if (gtid == 0) tt0 = globaltimer_ns();
        if (gtid == 0) {
            ctl->cnt[par ^ 1] = 0; ctl->stage[par ^ 1] = 0;
            ctl->pushes += cnt; ctl->iters += 1;
        }
        u64 edges_local = 0;
        for (i64 tile = gwarp * 128; tile < cnt; tile += nwarps * 128) {
            u64 pk[4], y[4], item[4], x[4];
            u32 cls[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { i64 i = tile + k * 32 + lane; item[k] = i < cnt ? A.front[par][i] : ~0ull; }
#pragma unroll
            for (int k = 0; k < 4; k++) x[k] = item[k] != ~0ull ? A.r[(i64)(u32)(item[k] >> 32) * n + (u32)item[k]] : 0;   // loads first,
#pragma unroll
            for (int k = 0; k < 4; k++) pk[k] = item[k] != ~0ull ? __ldg(&A.in_vtx[(u32)item[k]]) : 0;                      // stores afterwards
            u32 my_nch = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                y[k] = 0; cls[k] = 0;
                if (item[k] != ~0ull) {
                    const i64 idx = (i64)(u32)(item[k] >> 32) * n + (u32)item[k];
                    A.r[idx] = 0;
                    u64 a = mulhi64(x[k], A.alpha_q64);
                    A.reserve[idx] += a;
                    y[k] = x[k] - a;
                    u32 d = vtx_deg(pk[k]);
                    edges_local += d;
                    cls[k] = d == 0 ? 0u : (d <= PUSH_SMALL_MAX ? 1u : 2u);
                    if (cls[k] == 2) my_nch += (d + PUSH_CHUNK - 1) / PUSH_CHUNK;
                }
            }
            u32 ms[4], small_total = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) { ms[k] = __ballot_sync(0xffffffffu, cls[k] == 1); small_total += __popc(ms[k]); }
            u32 chunk_incl = 0, chunk_total = 0;
            if (__any_sync(0xffffffffu, my_nch != 0)) {
                chunk_incl = my_nch;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, chunk_incl, o); if (lane >= o) chunk_incl += t; }
                chunk_total = __shfl_sync(0xffffffffu, chunk_incl, 31);
            }
            if (small_total | chunk_total) {
                u64 base = 0;
                if (lane == 0) base = atomicAdd((ull*)&ctl->stage[par], ((ull)chunk_total << 32) | (ull)small_total);
                base = __shfl_sync(0xffffffffu, base, 0);
                u32 off = (u32)base;
                u32 cb = (u32)(base >> 32) + chunk_incl - my_nch;
                if ((i64)off + small_total > A.front_cap || (i64)(base >> 32) + chunk_total > A.chunk_cap) {
                    if (lane == 0) ctl->err_a = 1;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        if (cls[k] == 1) { u32 s = off + __popc(ms[k] & lanemask_lt()); A.s_pack[s] = pk[k]; A.s_y[s] = y[k]; A.s_t[s] = (u32)(item[k] >> 32); }
                        off += __popc(ms[k]);
                        if (cls[k] == 2) {
                            u32 nch = (vtx_deg(pk[k]) + PUSH_CHUNK - 1) / PUSH_CHUNK;
                            A.l_pack[cb] = pk[k]; A.l_y[cb] = y[k]; A.l_t[cb] = (u32)(item[k] >> 32);
                            for (u32 c = 0; c < nch; c++) A.chunks[cb + c] = make_uint2(cb, c);
                            cb += nch;
                        }
                    }
                }
            }
        }
        edges_local = warp_sum_u64(edges_local);
        if (lane == 0 && edges_local) atomicAdd((ull*)&ctl->edges, (ull)edges_local);
        grid.sync();
        if (*(volatile u32*)&ctl->err_a) break;
        if (gtid == 0) tt1 = globaltimer_ns();
        // ---------------- phase B
        const u64 staged = *(volatile u64*)&ctl->stage[par];
        const u32 ns = (u32)staged, nch = (u32)(staged >> 32);
        u64* next = A.front[par ^ 1];
        u32* next_cnt = &ctl->cnt[par ^ 1];
        u64* w_excl = sm_excl + wib * 32; u64* w_beg = sm_beg + wib * 32; u64* w_y = sm_y + wib * 32; u32* w_t = sm_t + wib * 32;
        const i64 bw = min((i64)32, max((i64)1, ((i64)ns + nwarps - 1) / nwarps));
        for (i64 wbase = gwarp * bw; wbase < ns; wbase += nwarps * bw) {
            i64 item = wbase + lane;
            u64 pk = 0, y = 0; u32 st = 0;
            if (lane < bw && item < ns) { pk = A.s_pack[item]; y = A.s_y[item]; st = A.s_t[item]; }
            u32 d = vtx_deg(pk), incl = d;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            u32 excl = incl - d;
            u32 total = __shfl_sync(0xffffffffu, incl, 31);
            w_excl[lane] = excl; w_beg[lane] = vtx_begin(pk) - excl; w_y[lane] = y; w_t[lane] = st;
            __syncwarp();
            for (u32 eb = 0; eb < total; eb += 128) {
                u32 u[4], tt[4];
                u64 yy[4], nitem[4];
                bool cr[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    u32 e = eb + 32 * k + lane;
                    u[k] = 0xffffffffu; tt[k] = 0; yy[k] = 0;
                    if (e < total) {
                        int o = 0;
#pragma unroll
                        for (int s = 16; s > 0; s >>= 1) { int p = o + s; if ((u32)w_excl[p] <= e) o = p; }
                        u[k] = (u32)__ldg(&A.row_idx[w_beg[o] + e]);
                        tt[k] = w_t[o]; yy[k] = w_y[o];
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    cr[k] = u[k] != 0xffffffffu ? relax_in_edge(A, A.r + (i64)tt[k] * n, u[k], yy[k]) : false;
                    nitem[k] = ((u64)tt[k] << 32) | u[k];
                }
                warp_append4_u64(next, next_cnt, A.front_cap, &ctl->err, cr, nitem);
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
            for (u32 eb = 0; eb < len; eb += 128) {
                u32 u[4];
                u64 nitem[4];
                bool cr[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { u32 e = eb + 32 * k + lane; u[k] = e < len ? (u32)__ldg(&row[e]) : 0xffffffffu; }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    cr[k] = u[k] != 0xffffffffu ? relax_in_edge(A, rt, u[k], y) : false;
                    nitem[k] = ((u64)t << 32) | u[k];
                }
                warp_append4_u64(next, next_cnt, A.front_cap, &ctl->err, cr, nitem);
            }
        }
        grid.sync();
        if (gtid == 0) {
            u32 k = ctl->dbg_count;
            if (k < DBG_ITERS) { u64* d = A.dbg + (size_t)k * DBG_FIELDS; d[0] = tt0; d[1] = tt1; d[2] = globaltimer_ns(); d[3] = cnt; d[4] = ctl->edges; d[5] = 0; ctl->dbg_count = k + 1; }
        }
        par ^= 1;
    }
    if (*(volatile u32*)&ctl->err || *(volatile u32*)&ctl->err_a) return;

    // ---------------- ordered compaction of {reserve >= thr_fx} over the [T][n] matrix
    // Warp-granular: every warp owns a contiguous slice, counts its qualifying entries (pass 1), gets its
    // offset from the per-warp counts, and writes with ballots (pass 2).  No CTA barrier inside the loops
    // (a block-scan version spent ~7 ms per 1024-target batch in __syncthreads).
    const i64 total = (i64)A.T * n;
    const i64 per = (((total + nwarps - 1) / nwarps) + 31) & ~(i64)31;   // slice length, multiple of 32
    const i64 lo = min(total, gwarp * per), hi = min(total, lo + per);
    u32 c = 0;
    for (i64 i = lo + lane; i < hi; i += 32) { u64 x = A.reserve[i]; c += (x != 0 && x >= A.thr_fx); }
    c = (u32)warp_sum_u64(c);
    if (lane == 0) A.wcnt[gwarp] = c;
    grid.sync();
    u64 off = 0;
    for (i64 w = lane; w < gwarp; w += 32) off += A.wcnt[w];
    off = warp_sum_u64(off);
    for (i64 base = lo; base < hi; base += 32) {
        i64 i = base + lane;
        u64 x = i < hi ? A.reserve[i] : 0;
        bool has = x != 0 && x >= A.thr_fx;
        u32 mask = __ballot_sync(0xffffffffu, has);
        if (has) {
            u64 slot = off + __popc(mask & lanemask_lt());
            if ((i64)slot < A.out_cap) { A.out_v[slot] = (u32)(i % n); A.out_t[slot] = (u32)(A.t0 + i / n); A.out_val[slot] = x; }
        }
        off += __popc(mask);
    }
    if (gwarp == nwarps - 1 && lane == 0) {
        ctl->pairs_needed = off;
        ctl->pairs = min(off, (u64)A.out_cap);
    }
}
