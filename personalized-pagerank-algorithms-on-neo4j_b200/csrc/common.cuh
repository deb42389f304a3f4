// common.cuh -- shared device helpers and host-side engine structs of libpprb200.so (sm_100a only).
//
// Arithmetic contract (DESIGN.md "Fixed point"): residues and reserves are uint64 holding
// value * 2^62; all accumulation is integer so atomics commute and every schedule gives the same
// bits.  Reference formulas are cited per kernel (J/ = src/main/java/joezie/fora_neo4j/ of the
// reference repository).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pprb200.h"

typedef uint64_t u64;
typedef int64_t i64;
typedef uint32_t u32;
typedef int32_t i32;
typedef unsigned long long ull;

#define FX_ONE (1ull << 62)
#define FX_INV (1.0 / 4611686018427387904.0)  // 2^-62
#define FX_SCALE 4611686018427387904.0

// Packed vertex record: (first edge offset << 24) | degree.  One 8-byte load gives both, so a walk
// step touches one sector of the vertex table and one of the edge array.
#define VTX_DEG_BITS 24
#define VTX_DEG_MASK ((1ull << VTX_DEG_BITS) - 1)
#define VTX_MAX_DEG ((1ll << VTX_DEG_BITS) - 1)
#define VTX_MAX_EDGES (1ll << 40)

// Forward-push work classes
#ifndef PUSH_SMALL_MAX
#define PUSH_SMALL_MAX 64   // degree <= this: warp-cooperative gather over 32 vertices
#endif
#define PUSH_MINI_MAX PUSH_THREADS  // frontiers up to this size are iterated by one CTA without grid barriers
#define PUSH_MINI_POW2 (PUSH_THREADS <= 256 ? 256 : PUSH_THREADS <= 512 ? 512 : 1024)   // power of two >= PUSH_MINI_MAX
#define PUSH_MINI_EDGES 4096        // ... owning at most this many edges
#define PUSH_CHUNK 256      // larger rows are cut into chunks of this many edges, one warp per chunk

#define DBG_ITERS 4096
#define DBG_FIELDS 6   // t_start, t_after_A, t_after_B (globaltimer ns), frontier size, cumulative edges, mode
#ifndef HOT_SLOTS
#define HOT_SLOTS 4096      // hub vertices whose incoming adds are staged in shared memory (32 KB per CTA); a power of two
#endif
#define HOT_BIT 0x80000000u
#define HOT_MASK ((unsigned)HOT_SLOTS - 1u)
#ifndef PUSH_THREADS
#define PUSH_THREADS 512
#endif
// "small regime": frontiers of at most midi_v vertices owning at most midi_e edges are iterated by the first
// midi_ctas CTAs with a barrier among themselves; the rest of the grid waits at the next grid barrier
#define PUSH_MIDI_CTAS 32
#define PUSH_MIDI_V 4096
#define PUSH_MIDI_E (1 << 16)
#ifndef PUSH_CTAS_PER_SM
// One 512-thread CTA per SM: the push kernel then gets 126 registers per thread and does not spill (with two CTAs per
// SM, 64 registers, it spilled 380 bytes and the prefetching dense sweep lost 25 %: 3.62 -> 3.22 ms per query at R-MAT
// scale 22, 1.09 -> 1.01 ms at scale 20; profiles/r02_push_variants.txt).  16 warps per SM are enough for the dense
// phase B (same 4.0 ms either way), and the barrier-bound small iterations get cheaper with half the CTAs.
#define PUSH_CTAS_PER_SM 1
#endif
#define WALK_THREADS 256
#ifndef WALK_PER_LANE
#define WALK_PER_LANE 2      // walks in flight per lane
#endif
#ifndef WALK_CTAS_PER_SM
#define WALK_CTAS_PER_SM 4
#endif
#ifndef WALK_REFILL_MIN
#define WALK_REFILL_MIN 12
#endif

#define CUDA_TRY(expr)                                                       \
    do {                                                                     \
        cudaError_t e__ = (expr);                                            \
        if (e__ != cudaSuccess) return ctx_fail_cuda(ctx, e__, #expr, __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ u64 mulhi64(u64 a, u64 b) { return __umul64hi(a, b); }
__device__ __forceinline__ u64 vtx_begin(u64 p) { return p >> VTX_DEG_BITS; }
__device__ __forceinline__ u32 vtx_deg(u64 p) { return (u32)(p & VTX_DEG_MASK); }

// thr(u) of the forward push: J/Forward_Push.java:132 tests r/d >= rmax; d == 0 gives +inf >= rmax for
// any positive residue, i.e. threshold "one unit".
__device__ __forceinline__ u64 thr_fwd(u64 rmax_fx, u32 d) {
    if (d == 0) return 1ull;
    u64 hi = __umul64hi(rmax_fx, (u64)d);
    return hi ? ~0ull : rmax_fx * (u64)d;
}

struct Philox4 { u32 x0, x1, x2, x3; };
__device__ __forceinline__ Philox4 philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, u32 k0, u32 k1) {
    const u32 M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        u32 hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        u32 hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        u32 n0 = hi1 ^ c1 ^ k0;
        u32 n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x0 = c0; o.x1 = c1; o.x2 = c2; o.x3 = c3;
    return o;
}

// Loads with L2 eviction-priority hints: streamed edge arrays are marked evict-first so that they do not push the
// randomly accessed per-vertex tables (vertex records, residues, degrees) out of the 126 MB L2.
__device__ __forceinline__ u64 l2_policy_evict_first() { u64 p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 l2_policy_evict_last() { u64 p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 l2_policy_keep_fraction_25() { u64 p; asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.25;" : "=l"(p)); return p; }
__device__ __forceinline__ u64 l2_policy_keep_fraction_15() { u64 p; asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, 0.15;" : "=l"(p)); return p; }
__device__ __forceinline__ u32 ld_hint_u32(const void* p, u64 pol) {
    u32 v; asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ u64 ld_hint_u64(const void* p, u64 pol) {
    u64 v; asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol)); return v;
}

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
#endif

// ------------------------------------------------------------------------------------------ device-resident control block
// One per workspace; written by the kernels, read back by the host once per call.
struct PushCtl {
    // frontier bookkeeping (index = iteration parity)
    u32 cnt[2];        // frontier sizes
    u64 stage[2];      // forward push: staged work of the iteration, (chunk descriptors << 32) | short rows
    u32 ns[2];         // reverse push: staged short rows
    u32 nl[2];         // reverse push: staged long rows
    u32 nchunks[2];    // reverse push: chunk descriptors
    u64 dang[2];       // dangling mass headed for the source
    u32 fcnt[2];       // vertices pushed by the iteration (both frontier modes)
    u64 fedges[2];     // edges relaxed by the iteration
    u64 nedge[2];      // out-degree sum of the frontier list being built (LIST appends; degrees clamped to 2^20)
    // work counters
    u64 reserve_total; // sum of everything moved into reserve by pushes
    u64 pushes, edges, iters, rounds;
    u64 rmax_fx;
    double rmax;
    // walk preparation
    u64 rsum_fx;       // sum of (1-alpha) parts handed to walks
    i64 N;             // (long)(omega * r_sum)
    u64 total_walks;
    u32 n_sources;
    u32 cont;          // balance loop: 1 = run another rmax-halving round
    u64 steps;         // walk moves executed (this shard)
    u64 walks_done;    // walks executed (this shard)
    // radix select
    u32 sel_count;     // survivors written
    u32 sel_pos;       // positives seen
    u64 sel_kth;
    u32 err;           // device-side error flags
    u32 src_dangling;  // 1 = the query source has no out-edges (result is {src: 1})
    u32 err_a;         // error raised during a phase A (kept apart from err so each flag is only read behind a barrier)
    u32 mini_par;      // forward push: frontier parity after a run of single-CTA iterations
    u32 dbg_count;     // timeline entries written
    u32 sub_bar;       // forward push: arrival counter of the sub-grid barrier (small regime)
    u64 pairs;         // all-pair backward search: (v, t) pairs written by the batch
    u64 pairs_needed;  // ... and the number that qualified (> capacity means the batch must be rerun)
};

// ------------------------------------------------------------------------------------------ host structs
struct ppr_ctx {
    int device;
    int sm_count;
    size_t total_mem;
    std::string name;
    std::string err;
    cudaStream_t stream;   // compute
    cudaStream_t stream2;  // second compute stream of the batch pipeline
    cudaStream_t copy;     // result download (batch pipeline)
    cudaEvent_t ev[8];
    double walk_cost;
    int max_rounds;
    i64 scan_threshold;    // < 0: default (n)
    i32 batch_streams;     // ppr_fora_batch: 1 = queries back to back, 2 = two queries in flight
    i32 bwd_batch;         // all-pair backward search: targets per batch (0 = automatic)
    i32 hot_min_indeg;     // a vertex needs at least this in-degree to get a hub slot (graphs created afterwards)
    i32 midi_ctas;         // forward push, small regime: CTAs taking part (<= 1: off), vertex and edge limits
    i64 midi_v, midi_e;
    i32 walk_hint;         // walk kernel variant: -1 automatic (beyond the L2: eviction hints + 32-bit row pointers), 0 plain, 1 hints, 2 hints + 32-bit row pointers
    double pull_enter, pull_exit;   // forward push: pull-direction passes from / until this share of all edges in the frontier (enter < 0: off)
    // measurement helpers (pprb200_debug.h): per context, created on first use, destroyed by ppr_shutdown
    cudaEvent_t timer_ev[4]; bool timer_init;
    void* flush_buf;
    // top-k over caller vectors (ppr_topk): grow-only device scratch
    void* topk_buf; size_t topk_bytes;
    // multi-GPU (ppr_comm_init): NCCL communicator of this rank, bound at run time (comm.hpp)
    void* comm; i32 rank, world;
};

struct DevGraph {
    i64 n, m;
    u64* out_vtx;   // [n] packed (begin<<24 | out-degree)
    i32* col_idx;   // [m]
    u32* out_deg;   // [n]
    u32* row32;     // [n+1] 32-bit CSR row pointers (NULL when m >= 2^32): the walk kernel's vertex table beyond the L2
    u64* in_vtx;    // [n] packed over the CSC
    i32* row_idx;   // [m]
    u64* invdeg;    // [n] (2^64-1)/out_deg, 0 for dangling
    u32* push_idx;  // [m] col_idx with hub targets encoded as HOT_BIT | slot
    u32* hot_id;    // [n_hot] slot -> vertex
    u32 n_hot;
    // pull-direction iterations of the forward push (pull.cuh)
    u32* pull_idx;      // [m] row_idx with the highest out-degree sources encoded as HOT_BIT | slot
    u32* pull_hot_id;   // [n_pull_hot] slot -> vertex
    u32 n_pull_hot;
    u32* long_v; u32 n_long;             // vertices with more than PULL_LONG in-edges
    uint2* long_chunk; u32 n_long_chunk; // (vertex, chunk number)
    u32* pull_task; u32 n_pull_task;     // n_pull_task + 1 boundaries, in 32-vertex groups
};

struct Workspace {
    i64 n;
    u64* r;         // residue
    u64* reserve;   // reserve / result
    u64* est;       // FORA top-k estimate (push reserve + walk contributions)
    u64* q[2];      // pull-direction iterations: dense per-edge amounts of the current / next frontier
    u32* front[2];  // frontier lists
    u64* s_pack; u64* s_q;   // staged small items (packed vertex record, per-edge amount)
    u64* l_pack; u64* l_q;   // staged large items
    uint2* chunks;  i64 chunk_cap;
    // walk sources
    u32* ws_v; u64* ws_begin; u64* ws_inc; u64* ws_rem;
    u64* blk_a; u64* blk_b; u64* blk_c;  // per-CTA partials of the in-kernel scans
    double* out_f64;   // conversion buffer for downloads
    u32* hist;         // radix-select histograms
    i32* sel_ids; u64* sel_vals; i64 sel_cap;
    u64* dbg;          // per-iteration timeline of the last push kernel: DBG_FIELDS u64 per iteration
    PushCtl* ctl;      // device
    PushCtl* ctl_host; // pinned mirror
    int coop_blocks;   // grid size of the cooperative push / reverse-push kernels
    int sel_blocks;    // grid size of the cooperative radix select (its own occupancy: 256-thread CTAs, up to four per SM)
};

struct ppr_graph {
    ppr_ctx* ctx;
    DevGraph d;
    Workspace ws[2];    // two workspaces: batch pipeline overlaps download of query i with query i+1
    int last_ws;        // workspace holding the last result
    int last_kind;      // 0 none, 1 reserve vector, 2 est vector
    std::vector<i64> h_rowptr_dummy;
    // stepped push (trace) state
    bool trace_active; int trace_parity; i32 trace_src; u64 trace_alpha_q64;
};

struct ppr_pairs {   // inverted index of Base_Whole_Graph: row v = (target, pi(v, target)) pairs, targets ascending (or the top-k rows)
    i64 n;
    i64 count;
    std::vector<i64> src_ptr;
    // device-resident result (the normal case): ppr_pairs_export copies straight into the caller's arrays
    int device; i32* d_tgt; u64* d_val;
    // host-resident result (top-k rows): on_host
    bool on_host;
    std::vector<i32> tgt;
    std::vector<u64> val;
};

int32_t ctx_fail(ppr_ctx* ctx, int32_t code, const std::string& msg);
int32_t ctx_fail_cuda(ppr_ctx* ctx, cudaError_t e, const char* expr, const char* file, int line);

// host-side parameter helpers (restated from J/Algo_Conf.java and the algorithm classes; see api.cu)
static inline u64 h_alpha_q64(double alpha) { return (u64)__builtin_ldexp(alpha, 64); }
static inline u64 h_stop_threshold(double alpha) { return (u64)__builtin_ceil(__builtin_ldexp(alpha, 53)); }
static inline u64 h_walk_key(u64 seed, u64 tag) { return seed + tag * 0x9E3779B97F4A7C15ull; }
