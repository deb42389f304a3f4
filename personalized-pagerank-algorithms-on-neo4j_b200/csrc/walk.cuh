// walk.cuh -- alpha-terminating random walks (Monte-Carlo and FORA's walk phase) for sm_100a.
//
// Replaces Monte_Carlo.random_walk / random_walk_no_zero_hop (J/Monte_Carlo.java:60-133) and the walk
// loops of J/Monte_Carlo.java:149-154, J/Fora_Whole_Graph.java:133-139, J/Fora_Topk.java:161-167.
//
// WALK_PER_LANE walks in flight per lane.  Randomness is counter based (Philox4x32-10, key = seed, counter = (start vertex,
// walk index, step)), so the terminal of walk (v, j) is a pure function of its arguments: the CPU
// replay in oracle/ reproduces every terminal, and the result does not depend on how walks are split
// over lanes, warps, launches or GPUs.  Contributions are integer (fixed-point increments or counts)
// and are added with red.global.add.u64.
//
// Scheduling: each warp owns a contiguous slice of the global walk index space and refills idle lanes
// every step (walk lengths are geometric; without refill a warp idles ~70 % of its lanes waiting for
// the longest of 32 walks).  Walk index -> (source, local index) uses the exclusive prefix ws_begin
// produced by walk_prep; because a warp's indices are consecutive, the lookup is a 33-entry window
// of that prefix held in shared memory that slides forward, not a global binary search per walk.
#pragma once
#include "common.cuh"

struct WalkArgs {
    const u64* out_vtx;
    const u32* row32;    // [n+1] 32-bit row pointers (m < 2^32), half the footprint of out_vtx: used by the beyond-L2 variant (HINT == 2)
    const i32* col_idx;
    const u32* ws_v;
    const u64* ws_begin;
    const u64* ws_inc;
    const u64* ws_rem;
    u64* target;       // vector receiving the contributions
    PushCtl* ctl;
    u64 shard, nshards;
    u32 k0, k1;        // Philox key
    u64 tstop;         // ceil(alpha * 2^53)
    i32 src;           // query source (restart target when restart_at_source)
    int restart_at_source;
    int no_zero_hop;
};

struct WalkState {
    u32 start, cur, restart, step;
    u64 j;
    bool first;  // next step is the mandatory first move of random_walk_no_zero_hop
};

// One step of a walk.  Returns true when the walk has stopped (w.cur is the terminal).
//   J/Monte_Carlo.java:75-90: draw U; U < alpha -> stop; else move to a uniform out-neighbour, or to the
//   restart vertex when the current vertex is dangling (no neighbour draw in that case).
//   J/Monte_Carlo.java:110-112: the first move of the no-zero-hop variant is unconditional.
template <bool HINT>
__device__ __forceinline__ bool walk_step(WalkState& w, const u64* __restrict__ out_vtx, const i32* __restrict__ col_idx,
                                          u32 k0, u32 k1, u64 tstop, u64& steps, u64 pol_keep, u64 pol_stream) {
    Philox4 x = philox4x32_10(w.start, (u32)w.j, (u32)(w.j >> 32), w.step, k0, k1);
    w.step++;
    if (!w.first) {
        u64 u53 = ((u64)x.x0 << 21) | (u64)(x.x1 >> 11);
        if (u53 < tstop) return true;
    }
    w.first = false;
    u64 pk = HINT ? ld_hint_u64(&out_vtx[w.cur], pol_keep) : __ldg(&out_vtx[w.cur]);
    u32 d = vtx_deg(pk);
    if (d) {
        u64 rnd = ((u64)x.x2 << 32) | (u64)x.x3;
        const i32* nb = &col_idx[vtx_begin(pk) + mulhi64(rnd, (u64)d)];
        w.cur = HINT ? ld_hint_u32(nb, pol_stream) : (u32)__ldg(nb);
    } else {
        w.cur = w.restart;
    }
    steps++;
    return false;
}

__device__ __forceinline__ u64 mul_div_u64(u64 a, u64 b, u64 c) {  // floor(a*b/c) for b <= c (the result fits in 64 bits), 128-bit product
    const u64 hi = __umul64hi(a, b), lo = a * b;
    if (hi == 0) return lo / c;
    // long division of (hi:lo) by c, c < 2^32 in every caller (shard and warp counts): two 64/32 steps
    u64 r = hi % c;                       // hi < c because a*b/c < 2^64
    u64 q_hi = ((r << 32) | (lo >> 32)) / c;
    r = ((r << 32) | (lo >> 32)) % c;
    u64 q_lo = ((r << 32) | (lo & 0xffffffffull)) / c;
    return (q_hi << 32) | q_lo;
}

// HINT: L2 eviction-priority hints (vertex records evict-last, neighbour sectors evict-first).  Worth 6 % of the
// walk time once the graph exceeds the L2 (scale 22+); costs 4 % in the issue-bound L2-resident case (scale 20).
// NW: walks in flight per lane.  A move is two dependent random loads (vertex record, neighbour), so a lane with one
// walk spends most of its time waiting; with NW walks the loads of each stage are issued together (predicated, no
// branches in between) and the latency is paid once per stage for all of them.
// HINT == 2: as 1, and the vertex record is read as two adjacent 32-bit row pointers from a table half the size of out_vtx
// (16 MB instead of 32 MB at R-MAT scale 22): the randomly accessed per-vertex data competes with the count / reserve vector and the
// streamed edge array for the L2, and every record miss costs a 64-byte DRAM fetch.
// HINT == 3 / 4 / 5 (experiments): as 2, with the neighbour sectors at normal priority / with an address-hashed 25 % / 15 % of them
// evict-last and the rest evict-first, so that a fixed part of the edge array stays in the L2 next to the row pointers.
// Measured at R-MAT scale 22 (profiles/r02_push_variants.txt, run r2x): FORA's short walk launches gain 2.5 % with 4 or 5, but the long
// Monte-Carlo launch loses 23 % (1.06e10 -> 8.1e9 walks/s): kept sectors + row pointers + the count vector oversubscribe the L2.  Default stays 2.
template <int HINT, int NW>
__global__ void __launch_bounds__(WALK_THREADS, WALK_CTAS_PER_SM) k_walk(WalkArgs A) {
    __shared__ u64 win[WALK_THREADS / 32][34];
    PushCtl* ctl = A.ctl;
    const u64 W = ctl->total_walks;
    const u32 nsrc = ctl->n_sources;
    if (W == 0 || nsrc == 0) return;
    const u64 lo = mul_div_u64(W, A.shard, A.nshards), hi = mul_div_u64(W, A.shard + 1, A.nshards);
    const u64 nw = ((u64)gridDim.x * blockDim.x) >> 5;
    const u64 gw = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const u64 span = hi - lo;
    const u64 wl = lo + mul_div_u64(span, gw, nw), wh = lo + mul_div_u64(span, gw + 1, nw);
    if (wl >= wh) return;
    const u32 lane = lane_id();
    u64* mywin = win[threadIdx.x >> 5];

    // source containing wl: largest s in [0, nsrc) with ws_begin[s] <= wl
    u32 scur;
    {
        u32 a = 0, b = nsrc;  // invariant: ws_begin[a] <= wl < ws_begin[b] (ws_begin[nsrc] = W > wl)
        while (b - a > 1) {
            u32 mid = a + ((b - a) >> 1);
            if (__ldg(&A.ws_begin[mid]) <= wl) a = mid; else b = mid;
        }
        scur = a;
    }
    const u64 pol_keep = l2_policy_evict_last();
    const u64 pol_stream = HINT == 4 ? l2_policy_keep_fraction_25() : HINT == 5 ? l2_policy_keep_fraction_15() : l2_policy_evict_first();
    u64 wnext = wl;
    // per-slot walk state (arrays are fully unrolled into registers)
    u32 w_start[NW], w_cur[NW], w_step[NW];
    u64 w_j[NW], inc[NW];
    bool active[NW], first[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) { w_start[k] = w_cur[k] = w_step[k] = 0; w_j[k] = 0; inc[k] = 0; active[k] = false; first[k] = false; }
    u64 steps = 0, done = 0;
    u64 src_acc = 0;  // contributions to the query source stay in a register: in Monte-Carlo 15 % of all walks stop on
                      // it, and same-address REDs serialise at ~0.7 ns each in L2

    for (;;) {
        u32 needy[NW], idle = 0;
#pragma unroll
        for (int k = 0; k < NW; k++) { needy[k] = __ballot_sync(0xffffffffu, !active[k]); idle += __popc(needy[k]); }
        // refill in batches: the bookkeeping below costs about as many instructions as a walk step, so wait until
        // WALK_REFILL_MIN slots per walk-in-flight are idle (or the warp has nothing left to do)
        if (wnext < wh && (idle >= NW * WALK_REFILL_MIN || idle == 32 * NW)) {
#pragma unroll
            for (int k = 0; k < NW; k++) {
                if (wnext >= wh || needy[k] == 0) continue;  // warp-uniform
                mywin[lane] = A.ws_begin[scur + lane];
                if (lane < 2) mywin[32 + lane] = A.ws_begin[scur + 32 + lane];
                __syncwarp();
                u32 rank = __popc(needy[k] & lanemask_lt());
                u64 idx = wnext + rank;
                if (!active[k] && idx < wh) {
                    int o = 0;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) { int p = o + s; if (mywin[p] <= idx) o = p; }
                    u32 si = scur + o;
                    w_start[k] = A.ws_v[si];
                    w_j[k] = idx - mywin[o];
                    inc[k] = A.ws_inc[si] + (w_j[k] == 0 ? A.ws_rem[si] : 0);
                    w_cur[k] = w_start[k];
                    w_step[k] = 0;
                    first[k] = A.no_zero_hop != 0;
                    active[k] = true;
                }
                u64 take = min((u64)__popc(needy[k]), wh - wnext);
                wnext += take;
                if (wnext < wh) {  // slide the window: source containing wnext (offset may be 32)
                    int o = 0;
#pragma unroll
                    for (int s = 32; s > 0; s >>= 1) { int p = o + s; if (p <= 32 && mywin[p] <= wnext) o = p; }
                    scur += (u32)o;
                }
                __syncwarp();
            }
        }
        bool any = false;
#pragma unroll
        for (int k = 0; k < NW; k++) any |= active[k];
        if (!__any_sync(0xffffffffu, any)) break;

        // stage 1: draws and stop tests (J/Monte_Carlo.java:75-90; :110-112 for the unconditional first move)
        u64 rnd[NW];
        bool go[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) {
            go[k] = false;
            rnd[k] = 0;
            if (active[k]) {
                Philox4 x = philox4x32_10(w_start[k], (u32)w_j[k], (u32)(w_j[k] >> 32), w_step[k], A.k0, A.k1);
                w_step[k]++;
                u64 u53 = ((u64)x.x0 << 21) | (u64)(x.x1 >> 11);
                go[k] = first[k] || u53 >= A.tstop;
                first[k] = false;
                rnd[k] = ((u64)x.x2 << 32) | (u64)x.x3;
            }
        }
        // stage 2: vertex records of all moving walks
        u64 rb[NW];   // first edge of the row
        u32 rd[NW];   // out-degree
#pragma unroll
        for (int k = 0; k < NW; k++) {
            rb[k] = 0; rd[k] = 0;
            if (go[k]) {
                if (HINT >= 2) {
                    const u32 b = ld_hint_u32(&A.row32[w_cur[k]], pol_keep), e = ld_hint_u32(&A.row32[w_cur[k] + 1], pol_keep);
                    rb[k] = b; rd[k] = e - b;
                } else {
                    const u64 pk = HINT ? ld_hint_u64(&A.out_vtx[w_cur[k]], pol_keep) : __ldg(&A.out_vtx[w_cur[k]]);
                    rb[k] = vtx_begin(pk); rd[k] = vtx_deg(pk);
                }
            }
        }
        // stage 3: neighbours (a dangling vertex sends the walk to its restart vertex without a neighbour draw)
        u32 nb[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) {
            const u32 d = rd[k];
            nb[k] = A.restart_at_source ? (u32)A.src : w_start[k];
            if (go[k] && d) {
                const i32* p = &A.col_idx[rb[k] + mulhi64(rnd[k], (u64)d)];
                nb[k] = (HINT && HINT != 3) ? ld_hint_u32(p, pol_stream) : (u32)__ldg(p);
            }
        }
        // stage 4: moves and terminals
#pragma unroll
        for (int k = 0; k < NW; k++) {
            if (go[k]) { w_cur[k] = nb[k]; steps++; }
            else if (active[k]) {
                if (w_cur[k] == (u32)A.src) src_acc += inc[k];
                else atomicAdd((ull*)&A.target[w_cur[k]], (ull)inc[k]);
                active[k] = false;
                done++;
            }
        }
    }
    steps = warp_sum_u64(steps);
    done = warp_sum_u64(done);
    src_acc = warp_sum_u64(src_acc);
    if (lane == 0) {
        if (src_acc) atomicAdd((ull*)&A.target[A.src], (ull)src_acc);
        atomicAdd((ull*)&ctl->steps, (ull)steps);
        atomicAdd((ull*)&ctl->walks_done, (ull)done);
    }
}

// Monte-Carlo setup: one source (the query vertex) with W = floor(omega) unit-weight walks
// (J/Monte_Carlo.java:145-154).  A dangling source yields W walks that all return it without a draw.
__global__ void k_mc_setup(u64* counts, i64 n, u32* ws_v, u64* ws_begin, u64* ws_inc, u64* ws_rem, PushCtl* ctl,
                           const u32* out_deg, i32 src, u64 W, u64 shard, u64 nshards) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 v = gtid; v < n; v += gsize) counts[v] = 0;
    if (gtid == 0) {
        ctl->steps = 0; ctl->walks_done = 0; ctl->pushes = 0; ctl->edges = 0; ctl->iters = 0; ctl->rounds = 0;
        ctl->rsum_fx = 0; ctl->N = (i64)W; ctl->err = 0; ctl->rmax_fx = 0; ctl->reserve_total = 0;
        bool dangling = out_deg[src] == 0;
        ctl->src_dangling = dangling ? 1u : 0u;
        ws_v[0] = (u32)src; ws_inc[0] = 1; ws_rem[0] = 0; ws_begin[0] = 0;
        for (int k = 1; k < 40; k++) ws_begin[k] = W;
        ctl->total_walks = dangling ? 0 : W;
        ctl->n_sources = dangling ? 0u : 1u;
    }
}
__global__ void k_mc_dangling(u64* counts, PushCtl* ctl, i32 src, u64 W, u64 shard, u64 nshards) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && ctl->src_dangling) {
        u64 lo = mul_div_u64(W, shard, nshards), hi = mul_div_u64(W, shard + 1, nshards);
        counts[src] = hi - lo;
        ctl->walks_done = hi - lo;
    }
}

// Terminals of consecutive walks, one per thread (test / replay API; same walk_step as k_walk).
__global__ void k_walk_terminals(const u64* out_vtx, const i32* col_idx, u32 start, u32 restart, u64 j0, i64 count,
                                 u32 k0, u32 k1, u64 tstop, int no_zero_hop, i32* out) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    if (vtx_deg(out_vtx[start]) == 0) { out[i] = (i32)start; return; }  // J/Monte_Carlo.java:70-72
    WalkState w; w.start = start; w.cur = start; w.restart = restart; w.step = 0; w.j = j0 + (u64)i; w.first = no_zero_hop != 0;
    u64 steps = 0;
    const u64 pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    while (!walk_step<false>(w, out_vtx, col_idx, k0, k1, tstop, steps, pol_keep, pol_stream)) {}
    out[i] = (i32)w.cur;
}
