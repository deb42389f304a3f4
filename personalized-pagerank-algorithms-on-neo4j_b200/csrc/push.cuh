// push.cuh -- frontier-synchronous forward push (Forward Push and FORA's push phase) for sm_100a.
//
// Replaces the FIFO loops of J/Forward_Push.java:85-141 (whole graph) and :165-248 (incremental,
// FORA top-k) and FORA's balance loop J/Fora_Whole_Graph.java:93-103.
//
// One synchronous iteration = two phases separated by a grid-wide barrier:
//   phase A (per frontier vertex): capture x = r[v]; reserve[v] += mulhi(x, alpha); r[v] = leftover;
//            stage (packed vertex record, per-edge amount q) by degree class;
//   phase B (per edge): old = atomicAdd(&r[u], q); the add that crosses thr(u) appends u to the next
//            frontier (exactly one add crosses, because after phase A every residue is below its
//            threshold and phase B only adds) -- no dense scan, no duplicates, no flags to clear.
// Work classes in phase B: rows of <= PUSH_SMALL_MAX edges are gathered 32 vertices per warp with a
// warp prefix sum (lanes walk the concatenated edge list, so neighbour loads stay dense even for
// degree-1 vertices); longer rows are cut into PUSH_CHUNK-edge chunks, one warp per chunk with
// coalesced 128-byte neighbour loads.  Dangling vertices reduce their mass per CTA and one thread adds
// it to r[source] (J/Forward_Push.java:101-115).
//
// Two ways to find the next frontier, chosen per iteration from the size of the current one:
//   LIST mode (sparse frontiers): phase B uses the returning atomic and the crossing test above;
//   SCAN mode (frontier edges >= scan_threshold, default n): phase B issues plain red.global.add.u64
//            (no return value, no degree load, no append) and the next phase A sweeps r[] densely,
//            capturing every vertex above threshold on the fly.  Both produce the same state bits.
//
// The whole query (all iterations, FORA's rmax-halving rounds and the walk-source scan) runs as ONE
// cooperative persistent kernel: frontier sizes never travel to the host.  The same device functions
// are also wrapped as plain kernels for the stepped/trace API used by the schedule tests.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
namespace cg = cooperative_groups;

struct PushArgs {
    // graph
    i64 n;
    const u64* out_vtx;
    const i32* col_idx;
    const u32* out_deg;
    const u32* push_idx;   // col_idx with hub targets replaced by HOT_BIT | slot (see hot_id)
    const u32* hot_id;     // slot -> vertex id of the n_hot highest in-degree vertices
    u32 n_hot;
    // state
    u64* r;
    u64* reserve;
    u64* est;
    u32* front[2];
    u64* s_pack; u64* s_q; u64* l_pack; u64* l_q;
    uint2* chunks;
    u32* ws_v; u64* ws_begin; u64* ws_inc; u64* ws_rem;
    u64* blk_a; u64* blk_b;
    PushCtl* ctl;
    u64* dbg;
    // query
    i32 src;
    u64 alpha_q64;
    double alpha;
    double omega;       // FORA omega (whole graph) or the round's omega (top-k)
    double walk_cost;
    double rmax0;       // initial rmax (double)
    int max_rounds;
    int mode;           // 0: plain push, one round | 1: FORA whole graph | 2: FORA top-k round
    int init;           // 1: reset state and seed r[src] = 1
    int zero_reserve_after_push;  // shard != 0 of a multi-GPU FORA: drop the push reserve before the walk prep
    int prep_adds_push_part;      // shard == 0
    i64 scan_threshold;           // frontier edges >= this: next frontier is found by a dense sweep (SCAN mode)
    int mini;                     // 1: frontiers of <= PUSH_MINI_MAX vertices are iterated by one CTA (push_mini)
    int midi_ctas;                // > 1: frontiers of <= midi_v vertices / midi_e edges are iterated by this many CTAs (push_small)
    u32 midi_v;
    u64 midi_e;
    // pull-direction iterations (pull.cuh): CSC with hub sources encoded, dense amount buffers, static work split
    const u64* in_vtx;
    const u32* pull_idx;          // row_idx with the PULL_SLOTS highest out-degree sources replaced by HOT_BIT | slot
    const u32* pull_hot_id; u32 n_pull_hot;
    const u32* long_v; u32 n_long;            // vertices with more than PULL_LONG in-edges
    const uint2* long_chunk; u32 n_long_chunk; // (vertex, chunk number)
    const u32* pull_task; u32 n_pull_task;     // boundaries (in 32-vertex groups) of runs with equal in-edge counts
    u64* q[2];
    i64 pull_enter, pull_exit;    // frontier edges: a dense iteration with >= pull_enter edges is followed by pull passes until the frontier has < pull_exit; pull_enter < 0: off
};

#define PMODE_PUSH 0
#define PMODE_FORA 1
#define PMODE_TOPK 2

// --------------------------------------------------------------------------------- small device utilities
__device__ __forceinline__ void warp_append_u32(u32* list, u32* counter, bool pred, u32 value) {
    u32 mask = __ballot_sync(0xffffffffu, pred);
    if (mask) {
        u32 leader = __ffs(mask) - 1;
        u32 base = 0;
        if (lane_id() == leader) base = atomicAdd(counter, (u32)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (pred) list[base + __popc(mask & lanemask_lt())] = value;
    }
}

// Four predicated values per lane, one counter round trip for the whole warp.  The out-degrees of the appended
// vertices (each clamped to 2^20, enough to compare against PUSH_MINI_EDGES) are summed into *edge_counter.
template <typename EdgeCounter>
__device__ __forceinline__ void warp_append4_u32(u32* list, u32* counter, EdgeCounter* edge_counter, const bool (&pred)[4],
                                                 const u32 (&value)[4], const u32 (&deg)[4], u32* mirror, u32 mirror_cap) {
    u32 m[4], tot = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { m[k] = __ballot_sync(0xffffffffu, pred[k]); tot += __popc(m[k]); }
    if (tot) {
        u32 dsum = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) if (pred[k]) dsum += min(deg[k], 1u << 20);
        dsum = __reduce_add_sync(0xffffffffu, dsum);
        u32 base = 0;
        if (lane_id() == 0) { base = atomicAdd(counter, tot); atomicAdd(edge_counter, (EdgeCounter)dsum); }
        base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (pred[k]) {
                u32 idx = base + __popc(m[k] & lanemask_lt());
                list[idx] = value[k];
                if (mirror && idx < mirror_cap) mirror[idx] = value[k];
            }
            base += __popc(m[k]);
        }
    }
}

__device__ __forceinline__ u64 globaltimer_ns() { u64 t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// --------------------------------------------------------------------------------- PTX helpers (sm_100a)
// Result-less adds are written as explicit `red`: with a fence anywhere in the kernel ptxas otherwise turns
// `atomicAdd` with an unused result into ATOMG ... RZ, which costs the dense phase B ~30 %.
__device__ __forceinline__ void red_add_u64(u64* p, u64 v) { asm volatile("red.global.add.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u32 ld_acquire_u32(const u32* p) { u32 v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release_u32(u32* p, u32 v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
// Barrier among the first K CTAs of the grid (the "small regime" below): one release-add and an acquire spin per
// CTA on a monotonic counter.  ~1.5 us for K = 32, against ~3-5 us for a barrier over the whole grid.
__device__ __forceinline__ void sub_sync(u32* bar, u32 K, u32& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += K;
        red_release_u32(bar, 1u);
        while (ld_acquire_u32(bar) < target) {}
    }
    __syncthreads();
}

// floor: below 1/alpha units mulhi64(x, alpha_q64) is 0 and a push would move nothing into the reserve
__device__ __forceinline__ u64 rmax_to_fx_dev(double rmax, u64 alpha_q64) {
    const u64 lo = ~0ull / alpha_q64 + 2;
    double x = __dmul_rn(rmax, FX_SCALE);
    if (!(x >= 1.0)) return lo;
    if (x >= 18446744073709551615.0) return ~0ull;
    u64 v = (u64)x;  // truncation
    return v > lo ? v : lo;
}

// --------------------------------------------------------------------------------- frontier seeding by dense scan
// {v : r[v] >= thr(v)}: start of a round (rmax just changed) -- the counterpart of Q / Q_next hand-over
// in J/Fora_Topk.java:145 and of the fresh queue in J/Forward_Push.java:81.
__device__ void push_scan(const PushArgs& A, int par, u64 rmax_fx, i64 gtid, i64 gsize) {
    PushCtl* ctl = A.ctl;
    for (i64 base = 0; base < A.n; base += gsize) {
        i64 v = base + gtid;
        bool pred = false;
        if (v < A.n) {
            u64 x = A.r[v];
            if (x) pred = x >= thr_fwd(rmax_fx, A.out_deg[v]);
        }
        warp_append_u32(A.front[par], &ctl->cnt[par], pred, (u32)v);
    }
}

// --------------------------------------------------------------------------------- phase A
// Per-vertex capture: x = r[v]; reserve[v] += mulhi(x, alpha); r[v] = leftover; stage (vertex record, per-edge amount)
// by degree class.  Two front ends:
//   push_phase_a_list : walks the frontier list (LIST mode), 128 consecutive items per warp, every load of a stage
//                       issued for all four items of a lane before anything is stored;
//   push_sweep_warp   : dense sweep over all vertices (SCAN mode), 128 consecutive vertices per warp; r[] and out_deg[] of
//                       the warp's next tile are requested before the current one is processed.
// The sweep moves ~170 MB per iteration at R-MAT scale 22 (r, degrees, vertex records and reserves of the takers in, r,
// reserve and staging out) and runs at about half of the DRAM bandwidth.  A version that streamed r[]/out_deg[] through
// shared memory with 1-D bulk copies (cp.async.bulk + mbarrier, 2048-vertex CTA tiles, two stages) was measured slower
// (75 vs 54 us per sweep at scale 22: two CTA barriers and the slot-reserving atomic sit in every tile's critical path)
// and was removed; profiles/r02_push_variants.txt.
struct PhaseATotals { u64 edges, resv, dang, pushed; };

__device__ __forceinline__ void phase_a_publish(PushCtl* ctl, int par, PhaseATotals t, u32* sm_cnt /*16 words, 8-byte aligned*/) {
    // CTA-level totals first (shared-memory atomics), then one global atomic per counter per CTA
    u64* sm_acc = reinterpret_cast<u64*>(sm_cnt + 8);
    const u32 lane = lane_id();
    __syncthreads();
    if (threadIdx.x < 4) sm_acc[threadIdx.x] = 0;
    __syncthreads();
    t.edges = warp_sum_u64(t.edges);
    t.resv = warp_sum_u64(t.resv);
    t.dang = warp_sum_u64(t.dang);
    t.pushed = warp_sum_u64(t.pushed);
    if (lane == 0) {
        if (t.pushed) atomicAdd((ull*)&sm_acc[0], (ull)t.pushed);
        if (t.edges) atomicAdd((ull*)&sm_acc[1], (ull)t.edges);
        if (t.resv) atomicAdd((ull*)&sm_acc[2], (ull)t.resv);
        if (t.dang) atomicAdd((ull*)&sm_acc[3], (ull)t.dang);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sm_acc[0]) atomicAdd(&ctl->fcnt[par], (u32)sm_acc[0]);
        if (sm_acc[1]) atomicAdd((ull*)&ctl->fedges[par], (ull)sm_acc[1]);
        if (sm_acc[2]) atomicAdd((ull*)&ctl->reserve_total, (ull)sm_acc[2]);
        if (sm_acc[3]) atomicAdd((ull*)&ctl->dang[par], (ull)sm_acc[3]);
    }
}

__device__ __forceinline__ void phase_a_recycle(PushCtl* ctl, int par) {  // the other parity's counters (last read before the preceding barrier)
    ctl->cnt[par ^ 1] = 0; ctl->stage[par ^ 1] = 0; ctl->dang[par ^ 1] = 0;
    ctl->fcnt[par ^ 1] = 0; ctl->fedges[par ^ 1] = 0; ctl->nedge[par ^ 1] = 0;
}

#ifndef PUSH_PA_ITEMS
#define PUSH_PA_ITEMS 4
#endif
constexpr int PA_ITEMS = PUSH_PA_ITEMS;   // vertices per lane and tile in phase A (list and dense sweep)

__device__ void push_phase_a_list(const PushArgs& A, int par, u64 rmax_fx, i64 gtid, i64 gsize, u32* sm_cnt) {
    PushCtl* ctl = A.ctl;
    const i64 cnt = (i64)*(volatile u32*)&ctl->cnt[par];
    if (gtid == 0) phase_a_recycle(ctl, par);
    PhaseATotals tot = {0, 0, 0, 0};
    const u32* F = A.front[par];
    const u32 lane = lane_id();
    const i64 gwarp = gtid >> 5, nwarps = gsize >> 5;
    for (i64 tile = gwarp * (32 * PA_ITEMS); tile < cnt; tile += nwarps * (32 * PA_ITEMS)) {
        u64 pk[PA_ITEMS], x[PA_ITEMS], rs[PA_ITEMS];
        u32 v[PA_ITEMS];
#pragma unroll
        for (int k = 0; k < PA_ITEMS; k++) { i64 i = tile + (i64)k * 32 + lane; v[k] = i < cnt ? F[i] : 0xffffffffu; }
#pragma unroll
        for (int k = 0; k < PA_ITEMS; k++) {   // every listed vertex is a taker: all three loads go out together
            const bool ok = v[k] != 0xffffffffu;
            x[k] = ok ? A.r[v[k]] : 0;
            pk[k] = ok ? A.out_vtx[v[k]] : 0;
            rs[k] = ok ? A.reserve[v[k]] : 0;
        }
        u64 q[PA_ITEMS];
        u32 cls[PA_ITEMS], my_nch = 0;  // cls: 0 none, 1 short row, 2 long row
#pragma unroll
        for (int k = 0; k < PA_ITEMS; k++) {
            q[k] = 0; cls[k] = 0;
            if (v[k] != 0xffffffffu) {
                const u32 d = vtx_deg(pk[k]);
                const u64 a = mulhi64(x[k], A.alpha_q64);
                const u64 y = x[k] - a;
                u64 left;
                if (d == 0) { left = 0; tot.dang += y; }
                else { q[k] = y / (u64)d; left = y - q[k] * (u64)d; }
                A.r[v[k]] = left;
                A.reserve[v[k]] = rs[k] + a;
                tot.resv += a; tot.edges += d; tot.pushed += 1;
                cls[k] = d == 0 ? 0u : (d <= PUSH_SMALL_MAX ? 1u : 2u);
                if (cls[k] == 2) my_nch += (d + PUSH_CHUNK - 1) / PUSH_CHUNK;
            }
        }
        u32 ms[PA_ITEMS], small_total = 0;
#pragma unroll
        for (int k = 0; k < PA_ITEMS; k++) { ms[k] = __ballot_sync(0xffffffffu, cls[k] == 1); small_total += __popc(ms[k]); }
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
#pragma unroll
            for (int k = 0; k < PA_ITEMS; k++) {
                if (cls[k] == 1) { u32 slot = off + __popc(ms[k] & lanemask_lt()); A.s_pack[slot] = pk[k]; A.s_q[slot] = q[k]; }
                off += __popc(ms[k]);
                if (cls[k] == 2) {  // a long row is identified by its first chunk slot
                    u32 nch = (vtx_deg(pk[k]) + PUSH_CHUNK - 1) / PUSH_CHUNK;
                    A.l_pack[cb] = pk[k]; A.l_q[cb] = q[k];
                    for (u32 c = 0; c < nch; c++) A.chunks[cb + c] = make_uint2(cb, c);
                    cb += nch;
                }
            }
        }
    }
    phase_a_publish(ctl, par, tot, sm_cnt);
}

// Warp-granular dense sweep: a warp takes 128 consecutive vertices at a time; r[] and out_deg[] of its NEXT tile are
// requested before the current tile is processed, vertex records and reserves of the current tile's takers are fetched
// together with the slot-reserving atomic, so a tile costs one dependent round trip instead of seven.
__device__ void push_sweep_warp(const PushArgs& A, int par, u64 rmax_fx, i64 gtid, i64 gsize, u32* sm_cnt) {
    PushCtl* ctl = A.ctl;
    if (gtid == 0) phase_a_recycle(ctl, par);
    PhaseATotals tot = {0, 0, 0, 0};
    const u32 lane = lane_id();
    const i64 n = A.n, gwarp = gtid >> 5, stride = (gsize >> 5) * (32 * PA_ITEMS);
    u64 xn[PA_ITEMS];
    u32 dn[PA_ITEMS];
    i64 tile = gwarp * (32 * PA_ITEMS);
#pragma unroll
    for (int k = 0; k < PA_ITEMS; k++) { const i64 v = tile + k * 32 + lane; xn[k] = v < n ? A.r[v] : 0; dn[k] = v < n ? A.out_deg[v] : 0; }
    for (; tile < n; tile += stride) {
        u64 x[PA_ITEMS], pk[PA_ITEMS], rs[PA_ITEMS];
        u32 d[PA_ITEMS], cls[PA_ITEMS], my_nch = 0;
        bool take[PA_ITEMS];
#pragma unroll
        for (int k = 0; k < PA_ITEMS; k++) { x[k] = xn[k]; d[k] = dn[k]; }
#pragma unroll
        for (int k = 0; k < PA_ITEMS; k++) {
            const i64 v = tile + k * 32 + lane;
            take[k] = x[k] != 0 && x[k] >= thr_fwd(rmax_fx, d[k]);
            pk[k] = (take[k] && d[k]) ? A.out_vtx[v] : 0;
            rs[k] = take[k] ? A.reserve[v] : 0;
            cls[k] = (!take[k] || d[k] == 0) ? 0u : (d[k] <= PUSH_SMALL_MAX ? 1u : 2u);
            if (cls[k] == 2) my_nch += (d[k] + PUSH_CHUNK - 1) / PUSH_CHUNK;
        }
        {   // next tile
            const i64 nt = tile + stride;
#pragma unroll
            for (int k = 0; k < PA_ITEMS; k++) { const i64 v = nt + k * 32 + lane; xn[k] = v < n ? A.r[v] : 0; dn[k] = v < n ? A.out_deg[v] : 0; }
        }
        u32 ms[PA_ITEMS], small_total = 0;
#pragma unroll
        for (int k = 0; k < PA_ITEMS; k++) { ms[k] = __ballot_sync(0xffffffffu, cls[k] == 1); small_total += __popc(ms[k]); }
        u32 chunk_incl = 0, chunk_total = 0;
        if (__any_sync(0xffffffffu, my_nch != 0)) {
            chunk_incl = my_nch;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { u32 c = __shfl_up_sync(0xffffffffu, chunk_incl, o); if (lane >= o) chunk_incl += c; }
            chunk_total = __shfl_sync(0xffffffffu, chunk_incl, 31);
        }
        u64 base = 0;
        if (lane == 0 && (small_total | chunk_total)) base = atomicAdd((ull*)&ctl->stage[par], ((ull)chunk_total << 32) | (ull)small_total);
        base = __shfl_sync(0xffffffffu, base, 0);
        u32 off = (u32)base;
        u32 cb = (u32)(base >> 32) + chunk_incl - my_nch;
#pragma unroll
        for (int k = 0; k < PA_ITEMS; k++) {
            if (take[k]) {
                const i64 v = tile + k * 32 + lane;
                const u64 a = mulhi64(x[k], A.alpha_q64);
                const u64 y = x[k] - a;
                u64 left, q = 0;
                if (d[k] == 0) { left = 0; tot.dang += y; }
                else { q = y / (u64)d[k]; left = y - q * (u64)d[k]; }
                A.r[v] = left;
                A.reserve[v] = rs[k] + a;
                tot.resv += a; tot.edges += d[k]; tot.pushed += 1;
                if (cls[k] == 1) { const u32 sl = off + __popc(ms[k] & lanemask_lt()); A.s_pack[sl] = pk[k]; A.s_q[sl] = q; }
                if (cls[k] == 2) {
                    const u32 nch = (d[k] + PUSH_CHUNK - 1) / PUSH_CHUNK;
                    A.l_pack[cb] = pk[k]; A.l_q[cb] = q;
                    for (u32 c = 0; c < nch; c++) A.chunks[cb + c] = make_uint2(cb, c);
                    cb += nch;
                }
            }
            off += __popc(ms[k]);
        }
    }
    phase_a_publish(ctl, par, tot, sm_cnt);
}

// Plain dense sweep (the round-1 form): r[] and the vertex record of every vertex that holds residue are loaded for
// SWEEP_ITEMS x 32 consecutive vertices per warp, nothing is prefetched across tiles.  Selected with -DSWEEP_PLAIN.
#ifndef SWEEP_ITEMS
#define SWEEP_ITEMS 4
#endif
__device__ void push_sweep_plain(const PushArgs& A, int par, u64 rmax_fx, i64 gtid, i64 gsize, u32* sm_cnt) {
    PushCtl* ctl = A.ctl;
    if (gtid == 0) phase_a_recycle(ctl, par);
    PhaseATotals tot = {0, 0, 0, 0};
    const u32 lane = lane_id();
    const i64 n = A.n, gwarp = gtid >> 5, nwarps = gsize >> 5;
    constexpr int IT = SWEEP_ITEMS;
    for (i64 tile = gwarp * (32 * IT); tile < n; tile += nwarps * (32 * IT)) {
        u64 pk[IT], q[IT], x[IT];
        u32 cls[IT];
#pragma unroll
        for (int k = 0; k < IT; k++) { const i64 v = tile + (i64)k * 32 + lane; x[k] = v < n ? A.r[v] : 0; }
#pragma unroll
        for (int k = 0; k < IT; k++) { const i64 v = tile + (i64)k * 32 + lane; pk[k] = x[k] ? A.out_vtx[v] : 0; }
        u32 my_nch = 0;
#pragma unroll
        for (int k = 0; k < IT; k++) {
            q[k] = 0; cls[k] = 0;
            const u32 d = vtx_deg(pk[k]);
            if (x[k] && x[k] >= thr_fwd(rmax_fx, d)) {
                const i64 v = tile + (i64)k * 32 + lane;
                const u64 a = mulhi64(x[k], A.alpha_q64);
                const u64 y = x[k] - a;
                u64 left;
                if (d == 0) { left = 0; tot.dang += y; }
                else { q[k] = y / (u64)d; left = y - q[k] * (u64)d; }
                A.r[v] = left;
                A.reserve[v] += a;
                tot.resv += a; tot.edges += d; tot.pushed += 1;
                cls[k] = d == 0 ? 0u : (d <= PUSH_SMALL_MAX ? 1u : 2u);
                if (cls[k] == 2) my_nch += (d + PUSH_CHUNK - 1) / PUSH_CHUNK;
            }
        }
        u32 ms[IT], small_total = 0;
#pragma unroll
        for (int k = 0; k < IT; k++) { ms[k] = __ballot_sync(0xffffffffu, cls[k] == 1); small_total += __popc(ms[k]); }
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
#pragma unroll
            for (int k = 0; k < IT; k++) {
                if (cls[k] == 1) { u32 slot = off + __popc(ms[k] & lanemask_lt()); A.s_pack[slot] = pk[k]; A.s_q[slot] = q[k]; }
                off += __popc(ms[k]);
                if (cls[k] == 2) {
                    u32 nch = (vtx_deg(pk[k]) + PUSH_CHUNK - 1) / PUSH_CHUNK;
                    A.l_pack[cb] = pk[k]; A.l_q[cb] = q[k];
                    for (u32 c = 0; c < nch; c++) A.chunks[cb + c] = make_uint2(cb, c);
                    cb += nch;
                }
            }
        }
    }
    phase_a_publish(ctl, par, tot, sm_cnt);
}

// --------------------------------------------------------------------------------- phase B
// One relaxed edge.  LIST: returning atomic + crossing test; otherwise a fire-and-forget reduction.
// Hub accumulator in shared memory: value = hi * 2^32 + lo.  Shared memory has no native 64-bit add (atomicAdd on u64
// compiles to an LDS + ATOMS.CAS retry loop, 14 % of all stall samples of the kernel in round 1's profile); two native
// 32-bit atomics do it exactly: the returning add on the low word detects its own wrap-around, which is carried into
// the high word.
__device__ __forceinline__ void hot_add(u32* sm_hot, u32 slot, u64 q) {
    const u32 lo = (u32)q;
    const u32 old = atomicAdd(&sm_hot[slot], lo);
    const u32 hi = (u32)(q >> 32) + ((old + lo) < old ? 1u : 0u);
    if (hi) atomicAdd(&sm_hot[HOT_SLOTS + slot], hi);
}

template <bool LIST>
__device__ __forceinline__ bool relax_edge(const PushArgs& A, u64 rmax_fx, u32& u, u64 q, u32* sm_hot, u32& deg) {
    // `u` comes from push_idx: hub targets are encoded as HOT_BIT | slot.
    if (LIST) {
        if (u & HOT_BIT) u = __ldg(&A.hot_id[u & HOT_MASK]);
        u64 old = atomicAdd((ull*)&A.r[u], (ull)q);
        deg = __ldg(&A.out_deg[u]);
        u64 th = thr_fwd(rmax_fx, deg);
        return old < th && old + q >= th;
    } else {
        // dense frontiers: adds to hub vertices are accumulated per CTA in shared memory and flushed once
        // per phase -- same-address RED.64 serialises at ~0.7 ns in L2, and a hub receives up to
        // max-in-degree adds per iteration (scripts/microbench: 1/256 of the adds on one address costs 30 %)
        if (u & HOT_BIT) hot_add(sm_hot, u & HOT_MASK, q);
        else red_add_u64(&A.r[u], q);
        return false;
    }
}

template <bool LIST>
__device__ void push_phase_b(const PushArgs& A, int par, u64 rmax_fx, i64 gtid, i64 gsize, u64* sm_excl /*[warps][32]*/,
                             u64* sm_beg, u64* sm_q, u32* sm_hot /*[2][HOT_SLOTS], all zero on entry and exit; unused when LIST*/) {
    PushCtl* ctl = A.ctl;
    const u64 pol_stream = l2_policy_evict_first();  // the edge array is streamed once per iteration: do not let it evict r[] / degrees
    const u64 staged = *(volatile u64*)&ctl->stage[par];
    const i64 ns = (i64)(u32)staged;
    const u32 nch = (u32)(staged >> 32);
    const i64 gwarp = gtid >> 5, nwarps = gsize >> 5;
    const u32 lane = lane_id();
    const int wib = threadIdx.x >> 5;
    u32* next = A.front[par ^ 1];
    u32* next_cnt = &ctl->cnt[par ^ 1];
    u64* w_excl = sm_excl + wib * 32;
    u64* w_beg = sm_beg + wib * 32;
    u64* w_q = sm_q + wib * 32;

    // (1) short rows: up to 32 vertices per warp, lanes sweep the concatenated edge list.  When the
    // frontier is small the batch shrinks so that every warp of the grid gets some rows (a warp works
    // through its batch serially, one dependent atomic round trip per 32 edges).
    const i64 bw = min((i64)32, max((i64)1, ((i64)ns + nwarps - 1) / nwarps));
    for (i64 wbase = gwarp * bw; wbase < ns; wbase += nwarps * bw) {
        i64 item = wbase + lane;
        u64 pk = 0, q = 0;
        if (lane < bw && item < ns) { pk = A.s_pack[item]; q = A.s_q[item]; }
        u32 d = vtx_deg(pk);
        u32 incl = d;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        u32 excl = incl - d;
        u32 total = __shfl_sync(0xffffffffu, incl, 31);
        w_excl[lane] = excl;
        w_beg[lane] = vtx_begin(pk) - excl;  // edge address of concatenated position e is w_beg[owner] + e
        w_q[lane] = q;
        __syncwarp();
        for (u32 eb = 0; eb < total; eb += 128) {  // four independent edges per lane in flight
            u32 u[4], dg[4];
            u64 qq[4];
            bool cr[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                u32 e = eb + 32 * k + lane;
                u[k] = 0xffffffffu;
                qq[k] = 0;
                if (e < total) {
                    int o = 0;
#pragma unroll
                    for (int s = 16; s > 0; s >>= 1) { int p = o + s; if ((u32)w_excl[p] <= e) o = p; }
                    u[k] = ld_hint_u32(&A.push_idx[w_beg[o] + e], pol_stream);
                    qq[k] = w_q[o];
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) { dg[k] = 0; cr[k] = u[k] != 0xffffffffu ? relax_edge<LIST>(A, rmax_fx, u[k], qq[k], sm_hot, dg[k]) : false; }
            if (LIST) warp_append4_u32(next, next_cnt, (ull*)&ctl->nedge[par ^ 1], cr, u, dg, (u32*)nullptr, 0u);
        }
        __syncwarp();
    }

    // (2) long rows: one warp per PUSH_CHUNK-edge chunk; SCAN mode keeps the whole chunk's ids in flight (PUSH_CHUNK / 32 = 8
    // loads per lane, then 8 reductions), LIST mode four (the crossing test needs the ids, degrees and flags live)
    for (i64 c = gwarp; c < nch; c += nwarps) {
        uint2 desc = A.chunks[c];
        u64 pk = A.l_pack[desc.x];
        u64 q = A.l_q[desc.x];
        u32 d = vtx_deg(pk);
        u32 off = desc.y * PUSH_CHUNK;
        u32 len = min((u32)PUSH_CHUNK, d - off);
        const u32* row = A.push_idx + vtx_begin(pk) + off;
        if (!LIST) {
            u32 u[PUSH_CHUNK / 32];
#pragma unroll
            for (int k = 0; k < PUSH_CHUNK / 32; k++) { u32 e = 32 * k + lane; u[k] = e < len ? ld_hint_u32(&row[e], pol_stream) : 0xffffffffu; }
#pragma unroll
            for (int k = 0; k < PUSH_CHUNK / 32; k++) {
                if (u[k] != 0xffffffffu) { if (u[k] & HOT_BIT) hot_add(sm_hot, u[k] & HOT_MASK, q); else red_add_u64(&A.r[u[k]], q); }
            }
            continue;
        }
        for (u32 eb = 0; eb < len; eb += 128) {
            u32 u[4], dg[4];
            bool cr[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { u32 e = eb + 32 * k + lane; u[k] = e < len ? ld_hint_u32(&row[e], pol_stream) : 0xffffffffu; }
#pragma unroll
            for (int k = 0; k < 4; k++) { dg[k] = 0; cr[k] = u[k] != 0xffffffffu ? relax_edge<LIST>(A, rmax_fx, u[k], q, sm_hot, dg[k]) : false; }
            if (LIST) warp_append4_u32(next, next_cnt, (ull*)&ctl->nedge[par ^ 1], cr, u, dg, (u32*)nullptr, 0u);
        }
    }

    if (!LIST && A.n_hot) {  // flush the CTA's hub accumulators (uniform: every thread of the CTA gets here)
        __syncthreads();
        for (u32 sl = threadIdx.x; sl < A.n_hot; sl += blockDim.x) {
            const u64 x = ((u64)sm_hot[HOT_SLOTS + sl] << 32) | (u64)sm_hot[sl];
            if (x) { red_add_u64(&A.r[__ldg(&A.hot_id[sl])], x); sm_hot[sl] = 0; sm_hot[HOT_SLOTS + sl] = 0; }
        }
        __syncthreads();
    }
    // (3) dangling mass -> source (J/Forward_Push.java:101-113) and the iteration's bookkeeping
    if (gtid == 0) {
        u64 dsum = *(volatile u64*)&ctl->dang[par];
        if (dsum) {
            u64 old = atomicAdd((ull*)&A.r[A.src], (ull)dsum);
            u64 th = thr_fwd(rmax_fx, A.out_deg[A.src]);
            if (LIST && old < th && old + dsum >= th) { u32 idx = atomicAdd(next_cnt, 1u); next[idx] = (u32)A.src; atomicAdd((ull*)&ctl->nedge[par ^ 1], (ull)A.out_deg[A.src]); }
        }
        ctl->iters += 1;
        ctl->pushes += *(volatile u32*)&ctl->fcnt[par];
        ctl->edges += *(volatile u64*)&ctl->fedges[par];
    }
}

// --------------------------------------------------------------------------------- block-level helpers (walk preparation)
__device__ __forceinline__ u64 block_sum_u64(u64 v, u64* sm /*>= 32 entries*/) {
    v = warp_sum_u64(v);
    __syncthreads();
    if (lane_id() == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    u64 t = 0;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : 0;
        t = warp_sum_u64(t);
        if (threadIdx.x == 0) sm[0] = t;
    }
    __syncthreads();
    t = sm[0];
    __syncthreads();
    return t;
}

// exclusive scan of (a, b) pairs over the block; returns this thread's exclusive prefix and the block totals
__device__ __forceinline__ void block_excl_scan2(u64 a, u64 b, u64* sm /*>= 64 entries*/, u64& ea, u64& eb, u64& ta, u64& tb) {
    const u32 lane = lane_id(), w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    u64 ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u64 xa = __shfl_up_sync(0xffffffffu, ia, o), xb = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += xa; ib += xb; }
    }
    __syncthreads();
    if (lane == 31) { sm[w] = ia; sm[32 + w] = ib; }
    __syncthreads();
    u64 wa = 0, wb = 0;
    ta = 0; tb = 0;
    for (u32 k = 0; k < nw; k++) {
        u64 xa = sm[k], xb = sm[32 + k];
        if (k < w) { wa += xa; wb += xb; }
        ta += xa; tb += xb;
    }
    ea = wa + ia - a;
    eb = wb + ib - b;
    __syncthreads();
}

// --------------------------------------------------------------------------------- single-CTA iterations
// The head and the tail of every round are frontiers of a few (hundred) vertices, and a grid-wide iteration costs
// ~22 us however little it does (two grid barriers and a dozen dependent L2 round trips on the counters).  While the
// frontier holds at most PUSH_MINI_MAX vertices, CTA 0 runs the same synchronous iterations alone: one frontier
// vertex per thread, frontier / staging / counters in shared memory, __syncthreads instead of grid barriers
// (3-8 us per iteration).  The other CTAs wait at the next grid barrier.  Same schedule, same bits.  Entry needs the
// frontier's edge count too (ctl->nedge, accumulated by the LIST appends): 30 vertices can own 50 K edges.
// Ordering inside the CTA rests on __syncthreads alone (store -> barrier -> atomic on one address from one SM).  Do NOT add
// __threadfence() here: with a fence anywhere in the kernel ptxas stops turning result-less atomics into REDG (it emits
// ATOMG with an RZ destination instead) and the dense phase B drops from ~200 to ~140 G edges/s.
__device__ __noinline__ void push_mini(const PushArgs& A, int par, u64 rmax_fx, u64* m_excl, u64* m_beg, u64* m_q, u64* scratch /*64*/,
                          u32* mf /*PUSH_MINI_MAX*/, u32* m_ctr /*2*/) {
    PushCtl* ctl = A.ctl;
    const u32 tid = threadIdx.x;
    u32 mc = __ldcg(&ctl->cnt[par]);
    if (tid < mc) mf[tid] = __ldcg(&A.front[par][tid]);
    u32 ne = 0;
    u64* acc = scratch + 48;  // [iters, pushes, edges, reserve] of this run, kept by thread 0 (the scans use scratch[0..47])
    if (tid == 0) { acc[0] = 0; acc[1] = 0; acc[2] = ctl->edges; acc[3] = 0; }
    __syncthreads();
    for (;;) {
        u64 t0 = 0, t1 = 0;
        if (tid == 0) t0 = globaltimer_ns();
        // phase A: one frontier vertex per thread
        u64 a = 0, dangy = 0, pk = 0, q = 0;
        u32 d = 0;
        if (tid < mc) {
            const u32 v = mf[tid];
            const u64 x = __ldcg(&A.r[v]);  // r[] is updated by L2 atomics between iterations: never through L1
            pk = A.out_vtx[v];
            d = vtx_deg(pk);
            a = mulhi64(x, A.alpha_q64);
            const u64 y = x - a;
            u64 left;
            if (d == 0) { left = 0; dangy = y; }
            else { q = y / (u64)d; left = y - q * (u64)d; }
            A.r[v] = left;
            A.reserve[v] += a;
        }
        u64 e_d, e_a, t_d, t_a;
        block_excl_scan2((u64)d, a, scratch, e_d, e_a, t_d, t_a);
        const u64 dang_total = block_sum_u64(dangy, scratch);
        m_excl[tid] = tid < mc ? e_d : ~0ull;
        m_beg[tid] = vtx_begin(pk) - e_d;  // edge address of concatenated position e is m_beg[owner] + e
        m_q[tid] = q;
        if (tid == 0) { m_ctr[0] = 0; m_ctr[1] = 0; t1 = globaltimer_ns(); }
        __syncthreads();
        // phase B: threads sweep the concatenated edge list, four independent edges per thread in flight
        u32* next = A.front[par ^ 1];
        for (u64 eb = 0; eb < t_d; eb += 4 * PUSH_THREADS) {
            u32 u[4], dg[4];
            u64 qq[4];
            bool cr[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const u64 e = eb + (u64)k * PUSH_THREADS + tid;
                u[k] = 0xffffffffu;
                qq[k] = 0;
                if (e < t_d) {
                    int o = 0;
#pragma unroll
                    for (int s = PUSH_MINI_POW2 / 2; s > 0; s >>= 1) { int p = o + s; if (p < PUSH_MINI_MAX && m_excl[p] <= e) o = p; }
                    u[k] = __ldg(&A.push_idx[m_beg[o] + e]);
                    qq[k] = m_q[o];
                }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) { dg[k] = 0; cr[k] = u[k] != 0xffffffffu ? relax_edge<true>(A, rmax_fx, u[k], qq[k], nullptr, dg[k]) : false; }
            warp_append4_u32(next, &m_ctr[0], &m_ctr[1], cr, u, dg, mf, (u32)PUSH_MINI_MAX);
        }
        __syncthreads();
        if (tid == 0) {
            if (dang_total) {  // dangling mass -> source (J/Forward_Push.java:101-113)
                u64 old = atomicAdd((ull*)&A.r[A.src], (ull)dang_total);
                u64 th = thr_fwd(rmax_fx, A.out_deg[A.src]);
                if (old < th && old + dang_total >= th) {
                    u32 idx = m_ctr[0]++;
                    next[idx] = (u32)A.src;
                    if (idx < PUSH_MINI_MAX) mf[idx] = (u32)A.src;
                    m_ctr[1] += min(A.out_deg[A.src], 1u << 20);
                }
            }
            acc[0] += 1; acc[1] += mc; acc[2] += t_d; acc[3] += t_a;
            u32 k = ctl->dbg_count;
            if (k < DBG_ITERS) {
                u64* dd = A.dbg + (size_t)k * DBG_FIELDS;
                dd[0] = t0; dd[1] = t1; dd[2] = globaltimer_ns(); dd[3] = mc; dd[4] = acc[2]; dd[5] = 4;
                ctl->dbg_count = k + 1;
            }
        }
        __syncthreads();
        par ^= 1;
        mc = m_ctr[0];
        ne = m_ctr[1];
        __syncthreads();
        if (mc == 0 || mc > PUSH_MINI_MAX || ne > PUSH_MINI_EDGES) break;
    }
    if (tid == 0) {  // hand the state back to the grid-wide loop: list in front[par], its size in cnt[par], all else clean
        ctl->cnt[par] = mc; ctl->cnt[par ^ 1] = 0;
        ctl->stage[0] = ctl->stage[1] = 0; ctl->dang[0] = ctl->dang[1] = 0;
        ctl->fcnt[0] = ctl->fcnt[1] = 0; ctl->fedges[0] = ctl->fedges[1] = 0; ctl->nedge[par] = ne; ctl->nedge[par ^ 1] = 0;
        ctl->iters += acc[0]; ctl->pushes += acc[1]; ctl->edges = acc[2]; ctl->reserve_total += acc[3];
        ctl->mini_par = (u32)par;
    }
}

// --------------------------------------------------------------------------------- the small regime
// Frontiers between the single-CTA limit and (midi_v vertices, midi_e edges) are too big for one CTA and far too small
// for the whole grid: a grid-wide iteration costs ~30 us there, nearly all of it barrier and dependent-round-trip latency.  The
// first K CTAs (one per SM) iterate them with a barrier among themselves (sub_sync, ~1.5 us) and hand frontiers that
// shrink further to CTA 0's push_mini without leaving the regime; the other CTAs wait at the next grid barrier.  Same
// phase functions, same schedule, same bits as the grid-wide iteration.
__device__ void push_small(const PushArgs& A, int par, u64 rmax_fx, u32 K, u32& bar_target, u64* sm_excl, u64* sm_beg, u64* sm_q, u32* sm_hot,
                           u64* sm_scan, u32* sm_mf, u32* sm_mctr, u32* sm_cnt) {
    PushCtl* ctl = A.ctl;
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)K * blockDim.x;
    for (;;) {
        const u32 c = *(volatile u32*)&ctl->cnt[par];
        const u64 ne = *(volatile u64*)&ctl->nedge[par];
        if (c == 0) break;
        if (A.mini && c <= PUSH_MINI_MAX && ne <= PUSH_MINI_EDGES) {
            if (blockIdx.x == 0) push_mini(A, par, rmax_fx, sm_excl, sm_beg, sm_q, sm_scan, sm_mf, sm_mctr);
            if (K > 1) sub_sync(&ctl->sub_bar, K, bar_target); else __syncthreads();
            par = (int)*(volatile u32*)&ctl->mini_par;
            continue;
        }
        if (K <= 1 || c > A.midi_v || ne > A.midi_e) break;
        u64 t0 = 0, t1 = 0;
        if (gtid == 0) t0 = globaltimer_ns();
        push_phase_a_list(A, par, rmax_fx, gtid, gsize, sm_cnt);
        sub_sync(&ctl->sub_bar, K, bar_target);
        if (gtid == 0) t1 = globaltimer_ns();
        push_phase_b<true>(A, par, rmax_fx, gtid, gsize, sm_excl, sm_beg, sm_q, sm_hot);
        sub_sync(&ctl->sub_bar, K, bar_target);
        if (gtid == 0) {
            u32 k = ctl->dbg_count;
            if (k < DBG_ITERS) {
                u64* d = A.dbg + (size_t)k * DBG_FIELDS;
                d[0] = t0; d[1] = t1; d[2] = globaltimer_ns(); d[3] = c; d[4] = ctl->edges; d[5] = 8;
                ctl->dbg_count = k + 1;
            }
        }
        par ^= 1;
    }
    if (gtid == 0) ctl->mini_par = (u32)par;
}

// --------------------------------------------------------------------------------- walk preparation (inside the cooperative kernel)
// FORA whole graph, J/Fora_Whole_Graph.java:116-131: every vertex with residue x moves mulhi(x,alpha)
// into its reserve, y = x - that; N = (long)(omega * r_sum); omega_v = ceil(y / r_sum * N); each of its
// walks carries y / omega_v (remainder on walk 0).  Dangling starts return themselves without a draw
// (J/Monte_Carlo.java:106-108), so their y goes straight to reserve[v].
// FORA top-k, J/Fora_Topk.java:148-160: no alpha split, N = (long)(omega * r_sum * (1-alpha)),
// omega_v = ceil(x * N), estimate = copy of the push reserve plus walk contributions; residues stay.
// Output: compacted walk sources in ascending vertex order with the exclusive prefix of walk counts.
template <bool TOPK>
__device__ void walk_prep(const PushArgs& A, cg::grid_group& grid, u64* sm) {
    PushCtl* ctl = A.ctl;
    const i64 n = A.n;
    const i64 per = (n + gridDim.x - 1) / gridDim.x;
    const i64 v_lo = min(n, (i64)blockIdx.x * per), v_hi = min(n, v_lo + per);
    u64* target = TOPK ? A.est : A.reserve;

    double rsum_d;
    i64 N;
    if (!TOPK) {
        u64 sy = 0;
        for (i64 v = v_lo + threadIdx.x; v < v_hi; v += blockDim.x) {
            u64 x = A.r[v];
            if (x) sy += x - mulhi64(x, A.alpha_q64);
        }
        sy = block_sum_u64(sy, sm);
        if (threadIdx.x == 0 && sy) atomicAdd((ull*)&ctl->rsum_fx, (ull)sy);
        grid.sync();
        u64 rsum_fx = *(volatile u64*)&ctl->rsum_fx;
        rsum_d = __dmul_rn((double)rsum_fx, FX_INV);
        N = (i64)__dmul_rn(A.omega, rsum_d);
    } else {
        u64 rsum_fx = FX_ONE - *(volatile u64*)&ctl->reserve_total;
        rsum_d = __dmul_rn((double)rsum_fx, FX_INV);
        double rsum_rw = __dmul_rn(rsum_d, 1.0 - A.alpha);
        N = (i64)__dmul_rn(A.omega, rsum_rw);
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->rsum_fx = rsum_fx;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->N = N;

    // pass 2: per-CTA totals of (sources, walks)
    u64 c_src = 0, c_w = 0;
    for (i64 v = v_lo + threadIdx.x; v < v_hi; v += blockDim.x) {
        u64 x = A.r[v];
        if (x && N > 0 && A.out_deg[v] != 0) {
            u64 y = TOPK ? x : x - mulhi64(x, A.alpha_q64);
            double yd = __dmul_rn((double)y, FX_INV);
            double t = TOPK ? __dmul_rn(yd, (double)N) : __dmul_rn(__ddiv_rn(yd, rsum_d), (double)N);
            i64 w = (i64)ceil(t);
            if (w > 0) { c_src += 1; c_w += (u64)w; }
        }
    }
    c_src = block_sum_u64(c_src, sm);
    c_w = block_sum_u64(c_w, sm);
    if (threadIdx.x == 0) { A.blk_a[blockIdx.x] = c_src; A.blk_b[blockIdx.x] = c_w; }
    grid.sync();

    // pass 3: CTA offsets, then ordered compaction of the chunk
    u64 o_src = 0, o_w = 0;
    for (u32 b = threadIdx.x; b < blockIdx.x; b += blockDim.x) { o_src += A.blk_a[b]; o_w += A.blk_b[b]; }
    o_src = block_sum_u64(o_src, sm);
    o_w = block_sum_u64(o_w, sm);
    for (i64 tile = v_lo; tile < v_hi; tile += blockDim.x) {
        i64 v = tile + threadIdx.x;
        u64 has = 0, w = 0, y = 0;
        if (v < v_hi) {
            u64 x = A.r[v];
            u64 base = (TOPK && !A.prep_adds_push_part) ? 0ull : A.reserve[v];  // top-k shards > 0 hold walk contributions only
            if (x) {
                u64 a = TOPK ? 0 : mulhi64(x, A.alpha_q64);
                y = x - a;
                if (!TOPK) { if (A.prep_adds_push_part) base += a; A.r[v] = 0; }
                if (N > 0) {
                    if (A.out_deg[v] == 0) { if (A.prep_adds_push_part) base += y; }
                    else {
                        double yd = __dmul_rn((double)y, FX_INV);
                        double t = TOPK ? __dmul_rn(yd, (double)N) : __dmul_rn(__ddiv_rn(yd, rsum_d), (double)N);
                        i64 wi = (i64)ceil(t);
                        if (wi > 0) { has = 1; w = (u64)wi; }
                    }
                }
            }
            if (TOPK || x) target[v] = base;
        }
        u64 e_src, e_w, t_src, t_w;
        block_excl_scan2(has, w, sm, e_src, e_w, t_src, t_w);
        if (has) {
            u64 slot = o_src + e_src;
            u64 inc = y / w;
            A.ws_v[slot] = (u32)v;
            A.ws_begin[slot] = o_w + e_w;
            A.ws_inc[slot] = inc;
            A.ws_rem[slot] = y - inc * w;
        }
        o_src += t_src; o_w += t_w;
    }
    if (blockIdx.x == gridDim.x - 1) {  // totals + sentinels for the walk kernel's 33-entry windows
        for (u32 k = threadIdx.x; k < 40; k += blockDim.x) A.ws_begin[o_src + k] = o_w;
        if (threadIdx.x == 0) { ctl->total_walks = o_w; ctl->n_sources = (u32)o_src; }
    }
}

#include "pull.cuh"

// --------------------------------------------------------------------------------- the persistent query kernel
__global__ void __launch_bounds__(PUSH_THREADS, PUSH_CTAS_PER_SM) k_push_query(const __grid_constant__ PushArgs A) {
    cg::grid_group grid = cg::this_grid();
    __shared__ u64 sm_excl[(PUSH_THREADS / 32) * 32];
    __shared__ u64 sm_beg[(PUSH_THREADS / 32) * 32];
    __shared__ u64 sm_q[(PUSH_THREADS / 32) * 32];
    __shared__ __align__(8) u32 sm_cnt[16];
    __shared__ u64 sm_scan[64];
    __shared__ u32 sm_mf[PUSH_MINI_MAX];
    __shared__ u32 sm_mctr[2];
    extern __shared__ __align__(16) u32 sm_hot[];   // [2][HOT_SLOTS]: low and high words of the per-CTA hub accumulators
    PushCtl* ctl = A.ctl;
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 gsize = (i64)gridDim.x * blockDim.x;
    for (u32 sl = threadIdx.x; sl < 2 * HOT_SLOTS; sl += blockDim.x) sm_hot[sl] = 0;
    u32 bar_target = 0;  // sub-grid barrier: arrivals expected at this thread's next sub_sync
    const u32 K = (A.midi_ctas > 1 && (u32)A.midi_ctas <= gridDim.x) ? (u32)A.midi_ctas : 1u;
    __syncthreads();

    if (A.init) {
        for (i64 v = gtid; v < A.n; v += gsize) { A.r[v] = 0; A.reserve[v] = 0; }
        if (gtid == 0) {
            ctl->cnt[0] = ctl->cnt[1] = 0; ctl->stage[0] = ctl->stage[1] = 0; ctl->dang[0] = ctl->dang[1] = 0; ctl->fcnt[0] = ctl->fcnt[1] = 0; ctl->fedges[0] = ctl->fedges[1] = 0; ctl->nedge[0] = ctl->nedge[1] = 0;
            ctl->reserve_total = 0; ctl->pushes = 0; ctl->edges = 0; ctl->iters = 0; ctl->rounds = 0;
            ctl->rmax = A.rmax0; ctl->rmax_fx = rmax_to_fx_dev(A.rmax0, A.alpha_q64);
            ctl->steps = 0; ctl->walks_done = 0; ctl->err = 0; ctl->src_dangling = 0; ctl->cont = 0; ctl->dbg_count = 0;
        }
        grid.sync();
    }
    if (gtid == 0) {
        ctl->rsum_fx = 0; ctl->N = 0; ctl->total_walks = 0; ctl->n_sources = 0; ctl->sub_bar = 0;
        if (!A.init) {  // top-k round: host supplied the round's rmax; counters of both parities must be clean
            ctl->cnt[0] = ctl->cnt[1] = 0; ctl->stage[0] = ctl->stage[1] = 0; ctl->dang[0] = ctl->dang[1] = 0; ctl->fcnt[0] = ctl->fcnt[1] = 0; ctl->fedges[0] = ctl->fedges[1] = 0; ctl->nedge[0] = ctl->nedge[1] = 0;
            ctl->rmax = A.rmax0; ctl->rmax_fx = rmax_to_fx_dev(A.rmax0, A.alpha_q64);
        }
    }
    const bool dangling_src = A.out_deg[A.src] == 0;  // J/Forward_Push.java:72-76, J/Fora_Topk.java:128-132
    if (dangling_src) {
        if (gtid == 0) {
            A.reserve[A.src] = A.prep_adds_push_part ? FX_ONE : 0;
            ctl->reserve_total = FX_ONE;
            ctl->src_dangling = 1;
            ctl->rmax_fx = 0;
        }
        if (A.mode == PMODE_TOPK) {
            grid.sync();
            for (i64 v = gtid; v < A.n; v += gsize) A.est[v] = A.reserve[v];
        }
        return;
    }
    if (A.init && gtid == 0) {
        A.r[A.src] = FX_ONE;
        // the first frontier is known: {src} if it is above threshold -- no dense sweep needed to find it
        if (FX_ONE >= thr_fwd(ctl->rmax_fx, A.out_deg[A.src])) { A.front[0][0] = (u32)A.src; ctl->cnt[0] = 1; ctl->nedge[0] = A.out_deg[A.src]; }
    }
    grid.sync();

    int par = 0, qb = 0;
    bool first_round = true;
    u64* sm_hq = reinterpret_cast<u64*>(sm_hot);
    for (;;) {  // rounds (FORA: rmax halves between rounds)
        const u64 rmax_fx = *(volatile u64*)&ctl->rmax_fx;
        bool have_list = first_round && A.init;  // later rounds start with a dense sweep: rmax just changed
        bool pull_mode = false;                  // true: the frontier of iteration `par` exists as dense amounts in q[qb] (pull.cuh)
        first_round = false;
        for (;;) {
            u64 t0 = 0, t1 = 0;
            if (gtid == 0) t0 = globaltimer_ns();
            if (pull_mode) {
                const u32 pf = *(volatile u32*)&ctl->fcnt[par];
                const u64 pe = *(volatile u64*)&ctl->fedges[par];
                const u64 pd = *(volatile u64*)&ctl->dang[par];
                if (pf == 0) break;
                if ((i64)pe >= A.pull_exit) {   // iteration `par` in pull form, fused with phase A of the next one
                    pull_main(A, par, rmax_fx, A.q[qb], A.q[qb ^ 1], sm_hq, sm_cnt, pd);
                    grid.sync();
                    if (gtid == 0) t1 = globaltimer_ns();
                    pull_long(A, par, rmax_fx, A.q[qb ^ 1], sm_cnt, pd, pf, pe);
                    grid.sync();
                    if (gtid == 0) {
                        u32 k = ctl->dbg_count;
                        if (k < DBG_ITERS) {
                            u64* d = A.dbg + (size_t)k * DBG_FIELDS;
                            d[0] = t0; d[1] = t1; d[2] = globaltimer_ns(); d[3] = pf; d[4] = ctl->edges; d[5] = 16;
                            ctl->dbg_count = k + 1;
                        }
                    }
                    par ^= 1; qb ^= 1;
                    continue;
                }
                pull_unstage(A, par, A.q[qb]);   // frontier too small for a pull pass: back to the push direction
                pull_mode = false;
            } else if (have_list) {
                const u32 c = *(volatile u32*)&ctl->cnt[par];
                if (c == 0) break;
                const u64 ne = *(volatile u64*)&ctl->nedge[par];
                const bool mini_ok = A.mini && c <= PUSH_MINI_MAX && ne <= PUSH_MINI_EDGES;
                if (mini_ok || (K > 1 && c <= A.midi_v && ne <= A.midi_e)) {
                    if (blockIdx.x < K) push_small(A, par, rmax_fx, K, bar_target, sm_excl, sm_beg, sm_q, sm_hot, sm_scan, sm_mf, sm_mctr, sm_cnt);
                    grid.sync();
                    par = (int)*(volatile u32*)&ctl->mini_par;
                    continue;
                }
                push_phase_a_list(A, par, rmax_fx, gtid, gsize, sm_cnt);
            } else {
#ifdef SWEEP_PLAIN
                push_sweep_plain(A, par, rmax_fx, gtid, gsize, sm_cnt);
#else
                push_sweep_warp(A, par, rmax_fx, gtid, gsize, sm_cnt);
#endif
            }
            grid.sync();
            const u32 fcnt = *(volatile u32*)&ctl->fcnt[par];
            const u64 fedges = *(volatile u64*)&ctl->fedges[par];
            if (fcnt == 0) break;
            const bool next_scan = (i64)fedges >= A.scan_threshold;
            if (gtid == 0) t1 = globaltimer_ns();
            if (next_scan) push_phase_b<false>(A, par, rmax_fx, gtid, gsize, sm_excl, sm_beg, sm_q, sm_hot);
            else push_phase_b<true>(A, par, rmax_fx, gtid, gsize, sm_excl, sm_beg, sm_q, sm_hot);
            grid.sync();
            if (gtid == 0) {
                u32 k = ctl->dbg_count;
                if (k < DBG_ITERS) {
                    u64* d = A.dbg + (size_t)k * DBG_FIELDS;
                    d[0] = t0; d[1] = t1; d[2] = globaltimer_ns(); d[3] = fcnt; d[4] = ctl->edges;
                    d[5] = (have_list ? 0 : 1) | (next_scan ? 2 : 0);
                    ctl->dbg_count = k + 1;
                }
            }
            have_list = !next_scan;
            par ^= 1;
            if (next_scan && A.pull_enter >= 0 && (i64)fedges >= A.pull_enter) {
                // a fat iteration is usually followed by another: capture the next frontier as dense amounts
                // (instead of staging lists) and continue in the pull direction
                u64 s0 = 0;
                if (gtid == 0) s0 = globaltimer_ns();
                pull_sweep_q(A, par, rmax_fx, A.q[qb], sm_cnt);
                grid.sync();
                pull_mode = true;
                if (gtid == 0) {
                    u32 k = ctl->dbg_count;
                    if (k < DBG_ITERS) {
                        u64* d = A.dbg + (size_t)k * DBG_FIELDS;
                        const u64 now = globaltimer_ns();
                        d[0] = s0; d[1] = now; d[2] = now; d[3] = 0; d[4] = ctl->edges; d[5] = 32;
                        ctl->dbg_count = k + 1;
                    }
                }
            }
        }
        if (gtid == 0) {
            // deterministic stand-in for the wall-clock balancer of J/Fora_Whole_Graph.java:93-103
            u64 rsum_fx = FX_ONE - ctl->reserve_total;
            double rsum_w = __dmul_rn(__dmul_rn((double)rsum_fx, FX_INV), 1.0 - A.alpha);
            double cost = (double)(ctl->pushes + ctl->edges);
            double rmax = __dmul_rn(ctl->rmax, 0.5);
            u64 rounds = ctl->rounds + 1;
            ctl->rounds = rounds;
            bool cont = A.mode == PMODE_FORA && cost < __dmul_rn(__dmul_rn(A.walk_cost, rsum_w), A.omega) &&
                        rounds < (u64)A.max_rounds;
            ctl->rmax = rmax;
            if (cont) ctl->rmax_fx = rmax_to_fx_dev(rmax, A.alpha_q64);
            ctl->cont = cont ? 1u : 0u;
        }
        grid.sync();
        if (*(volatile u32*)&ctl->cont == 0) break;
        // parity is kept: the iteration that ended the round left this parity's counters at zero
    }

    if (A.mode == PMODE_PUSH) return;
    if (A.zero_reserve_after_push) {
        for (i64 v = gtid; v < A.n; v += gsize) A.reserve[v] = 0;
        grid.sync();
    }
    if (A.mode == PMODE_FORA) walk_prep<false>(A, grid, sm_excl);
    else walk_prep<true>(A, grid, sm_excl);
}

// --------------------------------------------------------------------------------- stepped wrappers (trace API)
__global__ void __launch_bounds__(PUSH_THREADS) k_step_init(PushArgs A) {
    PushCtl* ctl = A.ctl;
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 v = gtid; v < A.n; v += gsize) { A.r[v] = 0; A.reserve[v] = 0; }
    if (gtid == 0) {
        ctl->cnt[0] = ctl->cnt[1] = 0; ctl->stage[0] = ctl->stage[1] = 0; ctl->dang[0] = ctl->dang[1] = 0; ctl->fcnt[0] = ctl->fcnt[1] = 0; ctl->fedges[0] = ctl->fedges[1] = 0; ctl->nedge[0] = ctl->nedge[1] = 0;
        ctl->reserve_total = 0; ctl->pushes = 0; ctl->edges = 0; ctl->iters = 0; ctl->rounds = 0;
        ctl->rmax = A.rmax0; ctl->rmax_fx = rmax_to_fx_dev(A.rmax0, A.alpha_q64);
        ctl->steps = 0; ctl->walks_done = 0; ctl->err = 0; ctl->src_dangling = 0; ctl->cont = 0;
        ctl->rsum_fx = 0; ctl->N = 0; ctl->total_walks = 0; ctl->n_sources = 0;
    }
}
__global__ void __launch_bounds__(PUSH_THREADS) k_step_seed(PushArgs A) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (A.out_deg[A.src] == 0) { A.reserve[A.src] = FX_ONE; A.ctl->reserve_total = FX_ONE; A.ctl->src_dangling = 1; }
        else A.r[A.src] = FX_ONE;
    }
}
__global__ void __launch_bounds__(PUSH_THREADS) k_step_scan(PushArgs A, int par) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    push_scan(A, par, A.ctl->rmax_fx, gtid, gsize);
}
__global__ void __launch_bounds__(PUSH_THREADS) k_step_a(PushArgs A, int par) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    __shared__ __align__(8) u32 sm_cnt[16];
    push_phase_a_list(A, par, A.ctl->rmax_fx, gtid, gsize, sm_cnt);
}
__global__ void __launch_bounds__(PUSH_THREADS) k_step_b(PushArgs A, int par) {
    __shared__ u64 sm_excl[(PUSH_THREADS / 32) * 32];
    __shared__ u64 sm_beg[(PUSH_THREADS / 32) * 32];
    __shared__ u64 sm_q[(PUSH_THREADS / 32) * 32];
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    push_phase_b<true>(A, par, A.ctl->rmax_fx, gtid, gsize, sm_excl, sm_beg, sm_q, nullptr);
}
