// bwdcta.cuh -- All-Pair-Backward-Search: one CTA per target, targets handed out by a ticket counter (sm_100a).
//
// Replaces the loop of Base_Whole_Graph.preprocessing (J/Base_Whole_Graph.java:76-92: one
// Backward_Search.backward_search_whole_graph per target, J/Backward_Search.java:38-100) and the grouping of the
// stored pairs by source (J/Base_Whole_Graph.java:83-87,112-163).
//
// Why not the grid-synchronous batch kernel of backward.cuh (kept for the single-target ppr_backward_search): with T
// targets advancing in lock step every iteration pays two grid barriers, and the dense T x n residue matrix is far
// larger than the L2, so every relaxed edge was a returning atomic into HBM (1.3-1.5e10 edges/s, 1.4-4.9 % of the HBM
// roofline in round 1).  The n searches are independent, so here a CTA owns one search from seed to convergence:
//   * iterations are synchronised with __syncthreads (a few us for the small frontiers that dominate at large
//     thresholds), hundreds of searches are in flight, nothing waits for the slowest target of a batch;
//   * the state of a search is sparse: dense rows r[n], reserve[n] exist per CTA *slot*, but only the entries a search
//     touches are ever read, written or cleaned (touched / pushed lists), so the L2 holds the live sectors of all slots;
//   * same synchronous schedule and the same 2^-62 fixed-point arithmetic as oracle mode B: iteration k pushes
//     {v : r[v] > rmax} (strict, absolute: J/Backward_Search.java:89) simultaneously; integer adds commute, so the
//     bits do not depend on which thread relaxes which edge.
// Pairs (v, t, reserve >= threshold) are appended to one device array in arrival order with key = v << B | t and then
// put into (source, target) order by an LSD radix sort written here (k_rs_*); row pointers come from the sorted keys.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "push.cuh"

struct BwdCtaArgs {
    i64 n;
    const u64* in_vtx; const i32* row_idx; const u64* invdeg;
    const u64* edge_inv;        // [m] invdeg[row_idx[e]] per in-edge (nullable): the reciprocal degree arrives with the edge instead of a dependent gather
    // per-slot state (slot = blockIdx.x), all rows n entries long
    u64* r; u64* reserve;       // dense rows, zero outside a search
    u32* front;                 // [slots][2][n] frontier lists
    u32* pushed;                // [slots][n] distinct vertices that were pushed (the only ones holding reserve)
    u32* touched;               // [slots][n] vertices whose residue left zero (duplicates possible; overflow -> dense clean)
    u64* st_pk; u64* st_y;      // [slots][n] staged (vertex record, y) of frontiers larger than one tile
    u32* ticket;                // next target offset
    const i32* order;           // targets of this launch in processing order (heaviest first), n_order of them; NULL: t_lo + ticket
    i32 n_order;
    i32 t_lo, t_hi;
    u64 alpha_q64, rmax_fx, thr_fx;
    int key_bits;               // B: key = v << B | t
    u64 scan_edges;             // an iteration whose frontier owns at least this many in-edges adds with plain reductions and finds the next frontier by a sweep of the row
    u64* out_key; u64* out_val; i64 out_cap;
    ull* counters;              // [0] pair cursor, [1] pushes, [2] edges, [3] iterations (summed over targets)
                                // timing != 0: [4] ticket+seed ns, [5] phase A ns, [6] phase B ns, [7] output+clean ns, [8] targets, [9] longest target ns (thread 0's clock)
    int timing;
    int dbg;                    // bring-up aid: 1 = tickets only, 2 = seed only, 3 = search without output; 0 = normal
};

// append to a per-CTA list whose cursor lives in shared memory; the first `mirror_cap` entries are mirrored in shared memory
__device__ __forceinline__ void cta_append(u32* list, u32* sm_cursor, bool pred, u32 value, u32* mirror, u32 mirror_cap, u32 cap, u32* sm_overflow) {
    const u32 mask = __ballot_sync(0xffffffffu, pred);
    if (mask) {
        const u32 leader = __ffs(mask) - 1;
        u32 base = 0;
        if (lane_id() == leader) base = atomicAdd(sm_cursor, (u32)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pred) {
            const u32 idx = base + __popc(mask & lanemask_lt());
            if (idx < cap) __stcg(&list[idx], value); else *sm_overflow = 1;
            if (mirror && idx < mirror_cap) mirror[idx] = value;
        }
    }
}

// four predicated values per lane, one shared-memory cursor update per warp
__device__ __forceinline__ void cta_append4(u32* list, u32* sm_cursor, const bool (&pred)[4], const u32 (&value)[4], u32* mirror, u32 mirror_cap, u32 cap, u32* sm_overflow) {
    u32 m[4], tot = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { m[k] = __ballot_sync(0xffffffffu, pred[k]); tot += __popc(m[k]); }
    if (tot) {
        u32 base = 0;
        if (lane_id() == 0) base = atomicAdd(sm_cursor, tot);
        base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (pred[k]) {
                const u32 idx = base + __popc(m[k] & lanemask_lt());
                if (idx < cap) __stcg(&list[idx], value[k]); else *sm_overflow = 1;
                if (mirror && idx < mirror_cap) mirror[idx] = value[k];
            }
            base += __popc(m[k]);
        }
    }
}

// Per-search cursors.  One copy per CTA in shared memory; a team (thread-block cluster) uses the copy of its rank-0 CTA through
// distributed shared memory (cluster.map_shared_rank), so appends from all CTAs of the team land in one place.
struct BwdCursors {
    u32 ticket;
    u32 ncount[2];      // frontier sizes by iteration parity
    u32 fedges[2];      // in-edges owned by the frontier (staged path), by iteration parity
    u32 npushed, ntouched, overflow;
    u32 outoff, total;  // output compaction
    ull outbase;
};

// TEAM == 1: one CTA per search.  TEAM > 1: a thread-block cluster of TEAM CTAs (TEAM SMs) per search -- for searches
// that are too large for one SM (the hubs at large thresholds hold the launch's critical path: 35 ms for one target of
// 70 ms at 1e-3; every search at small thresholds).  The CTAs of a team split phase A by frontier item and phase B by
// EDGE range (every CTA scans the same tile of up to BD frontier vertices, then relaxes its share of the concatenated
// in-edge list), share the cursors through DSMEM and separate phases with cluster.sync() (which also makes the other
// CTAs' global writes visible).  Fewer searches in flight also keeps their residue rows L2-resident.
template <int BD, int TEAM>
__global__ void __launch_bounds__(BD) k_bwd_cta(const __grid_constant__ BwdCtaArgs A) {
    __shared__ u64 m_excl[BD + 1];      // exclusive in-degree prefix of the tile's frontier vertices; ~0 beyond them (and as the sentinel)
    __shared__ u64 m_beg[BD];
    __shared__ u64 m_y[BD];
    __shared__ u32 mf[2][TEAM == 1 ? BD : 1];   // TEAM == 1: mirrors of the first BD entries of the two frontier lists
    __shared__ u64 scratch[64];
    __shared__ BwdCursors cur_local;
    const u32 tid = threadIdx.x;
    const i64 n = A.n;
    u32 rank = 0;
    BwdCursors* C = &cur_local;
    if (TEAM > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        rank = cluster.block_rank();
        C = cluster.map_shared_rank(&cur_local, 0);
    }
    auto team_sync = [&]() { if (TEAM > 1) cg::this_cluster().sync(); else __syncthreads(); };
    const size_t slot = blockIdx.x / TEAM;
    u64* R = A.r + slot * (size_t)n;
    u64* RES = A.reserve + slot * (size_t)n;
    u32* F[2] = {A.front + slot * 2 * (size_t)n, A.front + (slot * 2 + 1) * (size_t)n};
    u32* PUSHED = A.pushed + slot * (size_t)n;
    u32* TOUCHED = A.touched + slot * (size_t)n;
    u64* ST_PK = A.st_pk + slot * (size_t)n;
    u64* ST_Y = A.st_y + slot * (size_t)n;
    const u32 cap = (u32)n;
    const u32 tstride = (u32)TEAM * BD, tfirst = rank * BD + tid;   // item distribution over the team's threads

    for (;;) {
        team_sync();
        u64 tm0 = 0, tmA = 0, tmB = 0, tm1 = 0;
        const bool clock = A.timing && rank == 0 && tid == 0;
        if (clock) tm0 = globaltimer_ns();
        if (rank == 0 && tid == 0) C->ticket = atomicAdd(A.ticket, 1u);
        team_sync();
        const u32 ticket = C->ticket;
        if (A.order ? (i64)ticket >= (i64)A.n_order : (i64)A.t_lo + (i64)ticket >= (i64)A.t_hi) {
            team_sync();   // no CTA of a team may exit while another one still reads the cursors in its shared memory
            break;
        }
        const i64 t = A.order ? (i64)__ldg(&A.order[ticket]) : (i64)A.t_lo + (i64)ticket;
        if (A.dbg == 1) continue;
        const u64 pk_t = __ldg(&A.in_vtx[t]);
        if (vtx_deg(pk_t) == 0) {   // J/Backward_Search.java:46-49: a target without in-edges answers {t: 1}
            if (rank == 0 && tid == 0 && FX_ONE >= A.thr_fx) {
                const ull at = atomicAdd(&A.counters[0], 1ull);
                if ((i64)at < A.out_cap) { A.out_key[at] = ((u64)t << A.key_bits) | (u64)t; A.out_val[at] = FX_ONE; }
            }
            continue;
        }
        if (rank == 0 && tid == 0) {
            __stcg(reinterpret_cast<ull*>(&R[t]), (ull)FX_ONE);
            if (TEAM == 1) mf[0][0] = (u32)t;
            __stcg(&F[0][0], (u32)t); __stcg(&TOUCHED[0], (u32)t);
            C->ncount[0] = 1; C->ncount[1] = 0; C->fedges[0] = 0; C->fedges[1] = 0; C->ntouched = 1; C->npushed = 0; C->overflow = 0; C->total = 0; C->outoff = 0;
        }
        int par = 0;
        u64 acc_pushes = 0, acc_edges = 0, acc_iters = 0;   // rank 0, thread 0
        team_sync();
        if (A.dbg == 2) { if (rank == 0 && tid == 0) __stcg(reinterpret_cast<ull*>(&R[t]), 0ull); continue; }
        if (clock) { tm1 = globaltimer_ns(); tm0 = tm1 - tm0; }
        for (;;) {
            const u32 cnt = C->ncount[par];
            if (cnt == 0) break;
            if (rank == 0 && tid == 0) { C->ncount[par ^ 1] = 0; C->fedges[par ^ 1] = 0; }   // last read one iteration ago, behind two team barriers
            u32* next = F[par ^ 1];
            u32* next_m = TEAM == 1 ? mf[par ^ 1] : nullptr;
            // ---- phase A (J/Backward_Search.java:58-68): capture and zero the residue, move the alpha part to the reserve
            const bool single = TEAM == 1 && cnt <= (u32)BD;
            u64 pk = 0, y = 0;
            if (single) {
                bool fresh = false;
                u32 v = 0;
                if (tid < cnt) {
                    v = mf[par][tid];
                    const u64 x = __ldcg(&R[v]);       // residues are updated by L2 atomics: never through L1
                    R[v] = 0;
                    const u64 a = mulhi64(x, A.alpha_q64);
                    const u64 old = RES[v];
                    RES[v] = old + a;
                    fresh = old == 0;
                    y = x - a;
                    pk = __ldg(&A.in_vtx[v]);
                }
                cta_append(PUSHED, &C->npushed, fresh, v, nullptr, 0u, cap, &C->overflow);
            } else {
                u64 my_d = 0;
                for (u32 base = 0; base < cnt; base += tstride) {
                    const u32 i = base + tfirst;
                    bool fresh = false;
                    u32 v = 0;
                    if (i < cnt) {
                        v = __ldcg(&F[par][i]);
                        const u64 x = __ldcg(&R[v]);
                        __stcg(reinterpret_cast<ull*>(&R[v]), 0ull);
                        const u64 a = mulhi64(x, A.alpha_q64);
                        const u64 old = __ldcg(&RES[v]);
                        __stcg(reinterpret_cast<ull*>(&RES[v]), (ull)(old + a));
                        fresh = old == 0;
                        const u64 rec = __ldg(&A.in_vtx[v]);
                        my_d += vtx_deg(rec);
                        __stcg(reinterpret_cast<ull*>(&ST_PK[i]), (ull)rec);
                        __stcg(reinterpret_cast<ull*>(&ST_Y[i]), (ull)(x - a));
                    }
                    cta_append(PUSHED, &C->npushed, fresh, v, nullptr, 0u, cap, &C->overflow);
                }
                const u64 cta_d = block_sum_u64(my_d, scratch);
                if (tid == 0 && cta_d) atomicAdd(&C->fedges[par], (u32)min(cta_d, (u64)0x7fffffffull));
            }
            team_sync();   // every capture precedes every add of this iteration
            if (clock) { const u64 now = globaltimer_ns(); tmA += now - tm1; tm1 = now; }
            // ---- phase B (J/Backward_Search.java:77-94): tiles of BD frontier vertices; the team's CTAs take equal shares of a
            // tile's concatenated in-edge list, four consecutive edges per thread
            u64 it_edges = 0;
            u64 fed = single ? 0 : (u64)C->fedges[par];
            bool scan = false;
            for (u32 base = 0; base < cnt; base += BD) {
                if (!single) {
                    const u32 i = base + tid;
                    pk = i < cnt ? __ldcg(&ST_PK[i]) : 0;
                    y = i < cnt ? __ldcg(&ST_Y[i]) : 0;
                }
                const u32 d = vtx_deg(pk);
                u64 e_d, e_x, t_d, t_x;
                block_excl_scan2((u64)d, 0ull, scratch, e_d, e_x, t_d, t_x);
                m_excl[tid] = (base + tid < cnt) ? e_d : ~0ull;
                if (tid == 0) m_excl[BD] = ~0ull;
                m_beg[tid] = vtx_begin(pk) - e_d;   // edge address of concatenated position e is m_beg[owner] + e
                m_y[tid] = y;
                __syncthreads();
                it_edges += t_d;
                if (single) fed = t_d;
                scan = fed >= A.scan_edges;   // uniform over the team: dense iteration, see below
                const u64 e_lo = TEAM == 1 ? 0 : t_d * rank / TEAM, e_hi = TEAM == 1 ? t_d : t_d * (rank + 1) / TEAM;
                for (u64 eb = e_lo; eb < e_hi; eb += 4 * BD) {
                    // four consecutive edges per thread: one binary search for the first, the others advance linearly
                    u32 u[4];
                    u64 inc[4], old[4];
                    const u64 e0 = eb + 4ull * tid;
                    int o = 0;
                    if (e0 < e_hi) {
#pragma unroll
                        for (int s = BD / 2; s > 0; s >>= 1) { const int p = o + s; if (m_excl[p] <= e0) o = p; }
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const u64 e = e0 + k;
                        u[k] = 0xffffffffu; inc[k] = 0;
                        if (e < e_hi) {
                            while (m_excl[o + 1] <= e) o++;
                            u[k] = (u32)__ldg(&A.row_idx[m_beg[o] + e]);
                            inc[k] = m_y[o];
                            if (A.edge_inv) inc[k] = mulhi64(inc[k], __ldg(&A.edge_inv[m_beg[o] + e]));   // (1-alpha) r(v) / d_out(u)
                        }
                    }
                    if (!A.edge_inv) {
#pragma unroll
                        for (int k = 0; k < 4; k++) if (u[k] != 0xffffffffu) inc[k] = mulhi64(inc[k], __ldg(&A.invdeg[u[k]]));
                    }
                    if (scan) {   // dense iteration: fire-and-forget reductions, nothing waits for a returned value
#pragma unroll
                        for (int k = 0; k < 4; k++) if (u[k] != 0xffffffffu && inc[k]) red_add_u64(&R[u[k]], inc[k]);
                        continue;
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) old[k] = (u[k] != 0xffffffffu && inc[k]) ? atomicAdd((ull*)&R[u[k]], (ull)inc[k]) : 1ull;
                    bool first[4], cross[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const bool live = u[k] != 0xffffffffu && inc[k] != 0;
                        first[k] = live && old[k] == 0;
                        cross[k] = live && old[k] <= A.rmax_fx && old[k] + inc[k] > A.rmax_fx;   // strict r > rmax, exactly one add crosses
                    }
                    cta_append4(TOUCHED, &C->ntouched, first, u, nullptr, 0u, cap, &C->overflow);
                    cta_append4(next, &C->ncount[par ^ 1], cross, u, next_m, (u32)BD, cap, &C->overflow);
                }
                __syncthreads();
            }
            if (scan) {
                // Dense iteration (J/Backward_Search.java:84-94 with the test r > rmax applied after all adds): the residues above rmax
                // after the adds are exactly the vertices that crossed during the iteration (everything above rmax before it was in the
                // frontier and has been zeroed), so a sweep of the row finds the same next frontier as the per-add test.  The touched
                // list is not maintained: the row is cleaned densely when the search ends.
                if (tid == 0) C->overflow = 1;
                __threadfence();
                team_sync();   // every reduction of the iteration has been performed
                for (i64 base = 0; base < n; base += 4 * (i64)tstride) {
                    bool pr[4];
                    u32 vv[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const i64 v = base + (i64)k * tstride + tfirst;
                        vv[k] = (u32)v;
                        pr[k] = v < n && __ldcg(&R[v]) > A.rmax_fx;
                    }
                    cta_append4(next, &C->ncount[par ^ 1], pr, vv, next_m, (u32)BD, cap, &C->overflow);
                }
            }
            if (rank == 0 && tid == 0) { acc_pushes += cnt; acc_edges += it_edges; acc_iters += 1; }
            par ^= 1;
            team_sync();
            if (clock) { const u64 now = globaltimer_ns(); tmB += now - tm1; tm1 = now; }
        }
        // ---- output: pushed vertices with reserve >= threshold (J/Base_Whole_Graph.java:80-88); clean the slot
        if (A.dbg == 3) {
            const u32 np3 = C->npushed, nt3 = C->ntouched;
            for (u32 i = tfirst; i < np3; i += tstride) __stcg(reinterpret_cast<ull*>(&RES[__ldcg(&PUSHED[i])]), 0ull);
            for (u32 i = tfirst; i < nt3 && i < cap; i += tstride) __stcg(reinterpret_cast<ull*>(&R[__ldcg(&TOUCHED[i])]), 0ull);
            continue;
        }
        const u32 np = C->npushed;
        u32 mine = 0;
        for (u32 i = tfirst; i < np; i += tstride) { const u64 x = __ldcg(&RES[__ldcg(&PUSHED[i])]); mine += (x != 0 && x >= A.thr_fx) ? 1u : 0u; }
        const u64 cta_total = block_sum_u64((u64)mine, scratch);
        if (tid == 0 && cta_total) atomicAdd(&C->total, (u32)cta_total);
        team_sync();
        if (rank == 0 && tid == 0) {
            const u32 total = C->total;
            C->outbase = total ? atomicAdd(&A.counters[0], (ull)total) : 0ull;
            atomicAdd(&A.counters[1], (ull)acc_pushes); atomicAdd(&A.counters[2], (ull)acc_edges); atomicAdd(&A.counters[3], (ull)acc_iters);
        }
        team_sync();
        const ull obase = C->outbase;
        const bool fits = (i64)(obase + C->total) <= A.out_cap;
        for (u32 base = 0; base < np; base += tstride) {
            const u32 i = base + tfirst;
            u32 v = 0; u64 x = 0;
            if (i < np) { v = __ldcg(&PUSHED[i]); x = __ldcg(&RES[v]); __stcg(reinterpret_cast<ull*>(&RES[v]), 0ull); }
            const bool has = x != 0 && x >= A.thr_fx;
            const u32 mask = __ballot_sync(0xffffffffu, has);
            if (mask) {
                const u32 leader = __ffs(mask) - 1;
                u32 off = 0;
                if (lane_id() == leader) off = atomicAdd(&C->outoff, (u32)__popc(mask));
                off = __shfl_sync(0xffffffffu, off, leader);
                if (has && fits) {
                    const ull at = obase + off + __popc(mask & lanemask_lt());
                    A.out_key[at] = ((u64)v << A.key_bits) | (u64)t;
                    A.out_val[at] = x;
                }
            }
        }
        if (C->overflow) { for (i64 v = tfirst; v < n; v += tstride) __stcg(reinterpret_cast<ull*>(&R[v]), 0ull); }
        else { const u32 nt = C->ntouched; for (u32 i = tfirst; i < nt; i += tstride) __stcg(reinterpret_cast<ull*>(&R[__ldcg(&TOUCHED[i])]), 0ull); }
        if (A.timing) {
            team_sync();
            if (clock) {
                const u64 now = globaltimer_ns();
                atomicAdd(&A.counters[4], (ull)tm0); atomicAdd(&A.counters[5], (ull)tmA); atomicAdd(&A.counters[6], (ull)tmB);
                atomicAdd(&A.counters[7], (ull)(now - tm1)); atomicAdd(&A.counters[8], 1ull);
                atomicMax(&A.counters[9], (ull)(tm0 + tmA + tmB + (now - tm1)));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ LSD radix sort of (key, value) pairs
// 8-bit digits; three kernels per pass: per-block digit histograms, one exclusive scan over (digit, block), stable
// scatter (ranks inside a 256-item round from __match_any_sync + per-warp counts).
#define RS_THREADS 256
#define RS_ROUNDS 32
#define RS_TILE (RS_THREADS * RS_ROUNDS)

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const u64* __restrict__ keys, i64 count, int shift, u32* __restrict__ hist /*[256][nb]*/, u32 nb) {
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const i64 lo = (i64)blockIdx.x * RS_TILE, hi = min(count, lo + RS_TILE);
    for (i64 i = lo + threadIdx.x; i < hi; i += RS_THREADS) atomicAdd(&h[(u32)(keys[i] >> shift) & 255u], 1u);
    __syncthreads();
    hist[(size_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `len` u32 counters in place (one CTA of 1024 threads, 4 per thread per step, running carry)
__global__ void __launch_bounds__(1024) k_rs_scan(u32* __restrict__ a, i64 len) {
    __shared__ u32 wsum[32];
    __shared__ u32 carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (i64 base = 0; base < len; base += 4096) {
        const i64 i0 = base + (i64)threadIdx.x * 4;
        u32 x[4];
#pragma unroll
        for (int k = 0; k < 4; k++) x[k] = i0 + k < len ? a[i0 + k] : 0;
        const u32 mine = x[0] + x[1] + x[2] + x[3];
        u32 incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) wsum[w] = incl;
        __syncthreads();
        if (w == 0) {
            u32 s = wsum[lane], si = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const u32 t = __shfl_up_sync(0xffffffffu, si, o); if (lane >= o) si += t; }
            wsum[lane] = si - s;   // exclusive over warps
        }
        __syncthreads();
        u32 run = carry_s + wsum[w] + incl - mine;
#pragma unroll
        for (int k = 0; k < 4; k++) { if (i0 + k < len) a[i0 + k] = run; run += x[k]; }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = run;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const u64* __restrict__ keys, const u64* __restrict__ vals, i64 count, int shift,
                                                           const u32* __restrict__ offs /*[256][nb] exclusive*/, u32 nb,
                                                           u64* __restrict__ keys_out, u64* __restrict__ vals_out) {
    __shared__ u32 run[256];                          // next output position of every digit for this block
    __shared__ u32 wc[RS_THREADS / 32][256];          // per-warp digit counts of the current round
    const u32 tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    run[tid] = offs[(size_t)tid * nb + blockIdx.x];
    const i64 lo = (i64)blockIdx.x * RS_TILE, hi = min(count, lo + RS_TILE);
    for (i64 base = lo; base < hi; base += RS_THREADS) {
#pragma unroll
        for (int k = 0; k < RS_THREADS / 32; k++) wc[k][tid] = 0;
        __syncthreads();
        const i64 i = base + tid;
        const bool ok = i < hi;
        u64 key = 0, val = 0;
        if (ok) { key = keys[i]; val = vals[i]; }
        const u32 dg = ok ? ((u32)(key >> shift) & 255u) : 256u + lane;   // idle lanes never match each other or a digit
        const u32 peers = __match_any_sync(0xffffffffu, dg);
        const u32 rank = __popc(peers & lanemask_lt());
        if (ok && rank == 0) wc[w][dg] = __popc(peers);
        __syncthreads();
        if (ok) {
            u32 pos = run[dg] + rank;
            for (u32 k = 0; k < w; k++) pos += wc[k][dg];
            keys_out[pos] = key; vals_out[pos] = val;
        }
        __syncthreads();
        {
            u32 s = 0;
#pragma unroll
            for (int k = 0; k < RS_THREADS / 32; k++) s += wc[k][tid];
            run[tid] += s;
        }
        __syncthreads();
    }
}

// sorted keys -> targets and the row pointers of the index grouped by source (J/Base_Whole_Graph.java:83-87)
__global__ void k_rs_rows(const u64* __restrict__ keys, i64 count, int key_bits, i64 n, i32* __restrict__ tgt, i64* __restrict__ src_ptr) {
    const i64 gsize = (i64)gridDim.x * blockDim.x;
    const u64 mask = (1ull << key_bits) - 1;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gsize) {
        const u64 k = keys[i];
        const i64 v = (i64)(k >> key_bits);
        tgt[i] = (i32)(k & mask);
        const i64 vp = i ? (i64)(keys[i - 1] >> key_bits) : -1;
        for (i64 x = vp + 1; x <= v; x++) src_ptr[x] = i;
        if (i == count - 1) for (i64 x = v + 1; x <= n; x++) src_ptr[x] = count;
    }
}
