// pprb200.cu -- host side of libpprb200.so: contexts, graph export, kernel launches, the C ABI.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo (see Makefile).  There is no CPU path:
// every entry point that computes launches the kernels in push.cuh / walk.cuh / select.cuh /
// backward.cuh on the context's device and reports PPR_E_CUDA when that is impossible.
//
// Parameter derivations restate J/Algo_Conf.java:29-81, J/Fora_Whole_Graph.java:86-87,
// J/Fora_Topk.java:109-125 and J/Monte_Carlo.java:145 (J/ = src/main/java/joezie/fora_neo4j/).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "backward.cuh"
#include "bwdcta.cuh"
#include "comm.hpp"
#include "metrics.cuh"
#include "common.cuh"
#include "../../include/pprb200_debug.h"
#include "push.cuh"
#include "select.cuh"
#include "walk.cuh"

#define PUSH_DYN_SMEM ((size_t)(HOT_SLOTS > PULL_SLOTS ? HOT_SLOTS : PULL_SLOTS) * 8)   // dynamic shared memory of k_push_query: the per-CTA hub accumulators (push) / hub amounts (pull)
static thread_local std::string g_init_err;

int32_t ctx_fail(ppr_ctx* ctx, int32_t code, const std::string& msg) {
    if (ctx) ctx->err = msg; else g_init_err = msg;
    return code;
}
int32_t ctx_fail_cuda(ppr_ctx* ctx, cudaError_t e, const char* expr, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, expr);
    cudaGetLastError();
    return ctx_fail(ctx, e == cudaErrorMemoryAllocation ? PPR_E_NOMEM : PPR_E_CUDA, buf);
}

// ============================================================================================= graph build kernels
// all-pair backward search: reciprocal out-degree of the SOURCE of every in-edge, laid out like row_idx (built on first use)
__global__ void k_build_edge_inv(const i32* __restrict__ row_idx, const u64* __restrict__ invdeg, i64 m, u64* __restrict__ edge_inv) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 e = gtid; e < m; e += gsize) edge_inv[e] = invdeg[row_idx[e]];
}
__global__ void k_build_row32(const i64* __restrict__ ptr, u32* __restrict__ row32, i64 n) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 v = gtid; v <= n; v += gsize) row32[v] = (u32)ptr[v];
}
__global__ void k_build_vtx(const i64* __restrict__ ptr, u64* __restrict__ vtx, u32* __restrict__ deg, u64* __restrict__ invdeg,
                            i64 n) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 v = gtid; v < n; v += gsize) {
        i64 b = ptr[v], e = ptr[v + 1];
        u64 d = (u64)(e - b);
        vtx[v] = ((u64)b << VTX_DEG_BITS) | d;
        if (deg) deg[v] = (u32)d;
        if (invdeg) invdeg[v] = d ? ~0ull / d : 0ull;
    }
}

// CSC build on the device (the reference gets `incoming[][]` from HeavyGraph, J/PPR.java:136-152): in-degree histogram,
// exclusive scan, then a STABLE sort of the edges by target -- the CSR is ordered by source, so sources ascend inside
// every column, the canonical order of graphio/oracle.  The sort and the scan are CUB calls (one-time export, not a query
// kernel); everything else is below.
__global__ void k_indeg(const i32* __restrict__ col_idx, i64 m, i64 n, ull* __restrict__ indeg, u32* __restrict__ err) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 e = gtid; e < m; e += gsize) {
        i32 c = col_idx[e];
        if (c < 0 || (i64)c >= n) *err = 1u;
        else atomicAdd(&indeg[c], 1ull);
    }
}
__global__ void k_check_indeg(const ull* __restrict__ indeg, i64 n, u32* __restrict__ err) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 v = gtid; v < n; v += gsize) if (indeg[v] > (ull)VTX_MAX_DEG) *err = 2u;
}
// source vertex of every edge: one warp per row, coalesced stores
__global__ void k_edge_src(const u64* __restrict__ out_vtx, i64 n, i32* __restrict__ src) {
    const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((i64)gridDim.x * blockDim.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    for (i64 v = gw; v < n; v += nw) {
        const u64 pk = out_vtx[v];
        const u64 b = vtx_begin(pk), d = vtx_deg(pk);
        for (u64 e = lane; e < d; e += 32) src[b + e] = (i32)v;
    }
}
__global__ void k_set_slots(const u32* __restrict__ hot_id, u32 n_hot, i32* __restrict__ slot_of) {
    u32 s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_hot) slot_of[hot_id[s]] = (i32)s;
}
// push_idx = col_idx with hub targets replaced by HOT_BIT | slot (read by the push kernels only)
__global__ void k_encode_push_idx(const i32* __restrict__ col_idx, const i32* __restrict__ slot_of, i64 m, u32* __restrict__ push_idx) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 e = gtid; e < m; e += gsize) {
        i32 c = col_idx[e];
        i32 sl = slot_of[c];
        push_idx[e] = sl >= 0 ? (HOT_BIT | (u32)sl) : (u32)c;
    }
}

// ============================================================================================= context
extern "C" const char* ppr_version(void) { return "pprb200 0.1 (sm_100a)"; }

extern "C" int32_t ppr_init(int32_t device, ppr_ctx** out) {
    if (!out) return ctx_fail(nullptr, PPR_E_INVALID, "ppr_init: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return ctx_fail(nullptr, PPR_E_CUDA, std::string("ppr_init: no CUDA device (") + cudaGetErrorString(e) +
                                                  "); this library has no CPU fallback");
    }
    if (device < 0 || device >= count) return ctx_fail(nullptr, PPR_E_INVALID, "ppr_init: device ordinal out of range");
    ppr_ctx* ctx = new (std::nothrow) ppr_ctx();
    if (!ctx) return ctx_fail(nullptr, PPR_E_NOMEM, "ppr_init: host allocation failed");
    ctx->device = device;
    ctx->walk_cost = -1.0;  // auto: 6 when the graph lives in L2, 24 otherwise (measured on B200)
    ctx->max_rounds = 64;
    ctx->scan_threshold = -1;
    ctx->hot_min_indeg = 64;
    ctx->bwd_batch = 0;
    ctx->batch_streams = 0;
    ctx->walk_hint = -1;
    { const char* e = getenv("PPR_WALK_VARIANT"); if (e) ctx->walk_hint = atoi(e); }
    // pull-direction passes are OFF by default: inside the 16-warp-per-SM persistent kernel a pass is a chain of dependent
    // gathers and costs ~950 us at R-MAT scale 22 (profiles/r02_pull_direction.txt), more than the fattest push iteration
    ctx->pull_enter = -1.0; ctx->pull_exit = 0.0;
    { const char* e; if ((e = getenv("PPR_PULL_ENTER"))) ctx->pull_enter = atof(e); if ((e = getenv("PPR_PULL_EXIT"))) ctx->pull_exit = atof(e); }
    ctx->timer_init = false; ctx->flush_buf = nullptr; ctx->topk_buf = nullptr; ctx->topk_bytes = 0;
    ctx->comm = nullptr; ctx->rank = 0; ctx->world = 1;
    {   // small-regime knobs of the forward push (debug overrides: PPR_MIDI_CTAS / PPR_MIDI_V / PPR_MIDI_E)
        const char* e;
        ctx->midi_ctas = (e = getenv("PPR_MIDI_CTAS")) ? atoi(e) : PUSH_MIDI_CTAS;
        ctx->midi_v = (e = getenv("PPR_MIDI_V")) ? atoll(e) : PUSH_MIDI_V;
        ctx->midi_e = (e = getenv("PPR_MIDI_E")) ? atoll(e) : PUSH_MIDI_E;
    }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        delete ctx;
        return ctx_fail_cuda(nullptr, e, "cudaSetDevice/cudaGetDeviceProperties", __FILE__, __LINE__);
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->total_mem = prop.totalGlobalMem;
    ctx->name = prop.name;
    if (!prop.cooperativeLaunch) { delete ctx; return ctx_fail(nullptr, PPR_E_CUDA, "device lacks cooperative launch"); }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return ctx_fail_cuda(nullptr, e, "cudaStreamCreate", __FILE__, __LINE__);
    }
    for (int i = 0; i < 8; i++) cudaEventCreate(&ctx->ev[i]);
    *out = ctx;
    return PPR_OK;
}

extern "C" int32_t ppr_shutdown(ppr_ctx* ctx) {
    if (!ctx) return PPR_E_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy);
    for (int i = 0; i < 8; i++) cudaEventDestroy(ctx->ev[i]);
    if (ctx->timer_init) for (int i = 0; i < 4; i++) cudaEventDestroy(ctx->timer_ev[i]);
    cudaFree(ctx->flush_buf); cudaFree(ctx->topk_buf);
    if (ctx->comm) { NcclApi* api = nccl_api(nullptr); if (api) api->CommDestroy((ncclComm_t)ctx->comm); }
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->stream2);
    cudaStreamDestroy(ctx->copy);
    delete ctx;
    return PPR_OK;
}

extern "C" const char* ppr_last_error(ppr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_err.c_str(); }

extern "C" int32_t ppr_device_info(ppr_ctx* ctx, int32_t* sm_count, int64_t* total_mem, char* name, int32_t name_cap) {
    if (!ctx) return PPR_E_INVALID;
    if (sm_count) *sm_count = ctx->sm_count;
    if (total_mem) *total_mem = (int64_t)ctx->total_mem;
    if (name && name_cap > 0) { strncpy(name, ctx->name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    return PPR_OK;
}

extern "C" int32_t ppr_set_balance(ppr_ctx* ctx, double walk_cost, int32_t max_rounds) {
    if (!ctx) return PPR_E_INVALID;
    if (max_rounds < 1) return ctx_fail(ctx, PPR_E_INVALID, "ppr_set_balance: max_rounds >= 1");
    ctx->walk_cost = walk_cost;
    ctx->max_rounds = max_rounds;
    return PPR_OK;
}

extern "C" int32_t ppr_set_tuning(ppr_ctx* ctx, int64_t scan_threshold_edges, int32_t hot_min_indeg) {
    if (!ctx) return PPR_E_INVALID;
    ctx->scan_threshold = scan_threshold_edges;
    if (hot_min_indeg >= 0) ctx->hot_min_indeg = hot_min_indeg;
    return PPR_OK;
}

extern "C" int32_t ppr_set_batch_streams(ppr_ctx* ctx, int32_t streams) {
    if (!ctx || streams < 0 || streams > 2) return PPR_E_INVALID;
    ctx->batch_streams = streams;
    return PPR_OK;
}

extern "C" int32_t ppr_set_bwd_batch(ppr_ctx* ctx, int32_t targets_per_batch) {
    if (!ctx || targets_per_batch < 0) return PPR_E_INVALID;
    ctx->bwd_batch = targets_per_batch;
    return PPR_OK;
}

extern "C" int32_t ppr_timer_record(ppr_ctx* ctx, int32_t slot) {
    if (!ctx || slot < 0 || slot > 3) return PPR_E_INVALID;
    cudaSetDevice(ctx->device);
    if (!ctx->timer_init) { for (int i = 0; i < 4; i++) CUDA_TRY(cudaEventCreate(&ctx->timer_ev[i])); ctx->timer_init = true; }
    // the compute stream waits for everything queued on the download stream and the second compute stream
    CUDA_TRY(cudaEventRecord(ctx->ev[6], ctx->copy));
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev[6], 0));
    CUDA_TRY(cudaEventRecord(ctx->ev[7], ctx->stream2));
    CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev[7], 0));
    CUDA_TRY(cudaEventRecord(ctx->timer_ev[slot], ctx->stream));
    return PPR_OK;
}
extern "C" int32_t ppr_timer_elapsed_ms(ppr_ctx* ctx, int32_t a, int32_t b, float* ms_out) {
    if (!ctx || a < 0 || a > 3 || b < 0 || b > 3 || !ms_out || !ctx->timer_init) return PPR_E_INVALID;
    cudaSetDevice(ctx->device);
    CUDA_TRY(cudaEventSynchronize(ctx->timer_ev[b]));
    CUDA_TRY(cudaEventElapsedTime(ms_out, ctx->timer_ev[a], ctx->timer_ev[b]));
    return PPR_OK;
}
extern "C" int32_t ppr_flush_l2(ppr_ctx* ctx) {
    if (!ctx) return PPR_E_INVALID;
    cudaSetDevice(ctx->device);
    const size_t bytes = (size_t)256 << 20;
    if (!ctx->flush_buf) CUDA_TRY(cudaMalloc(&ctx->flush_buf, bytes));
    CUDA_TRY(cudaMemsetAsync(ctx->flush_buf, 0x5a, bytes, ctx->stream));
    return PPR_OK;
}
extern "C" int32_t ppr_set_pull(ppr_ctx* ctx, double enter_share, double exit_share) {
    if (!ctx) return PPR_E_INVALID;
    ctx->pull_enter = enter_share; ctx->pull_exit = exit_share;
    return PPR_OK;
}

extern "C" int32_t ppr_set_walk_variant(ppr_ctx* ctx, int32_t hint) {
    if (!ctx || hint < -1 || hint > 5) return PPR_E_INVALID;
    ctx->walk_hint = hint;
    return PPR_OK;
}
extern "C" int32_t ppr_device_sync(ppr_ctx* ctx) {
    if (!ctx) return PPR_E_INVALID;
    cudaSetDevice(ctx->device);
    CUDA_TRY(cudaDeviceSynchronize());
    return PPR_OK;
}

extern "C" int32_t ppr_host_alloc(ppr_ctx* ctx, int64_t bytes, void** out) {
    if (!ctx || !out || bytes < 0) return PPR_E_INVALID;
    cudaSetDevice(ctx->device);
    CUDA_TRY(cudaHostAlloc(out, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault));
    return PPR_OK;
}
extern "C" int32_t ppr_host_free(ppr_ctx* ctx, void* p) {
    if (!ctx) return PPR_E_INVALID;
    if (p) CUDA_TRY(cudaFreeHost(p));
    return PPR_OK;
}

// ============================================================================================= workspace
template <typename T>
static cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)); }

static void ws_free(Workspace& w) {
    cudaFree(w.r); cudaFree(w.reserve); cudaFree(w.est); cudaFree(w.front[0]); cudaFree(w.front[1]); cudaFree(w.q[0]); cudaFree(w.q[1]);
    cudaFree(w.s_pack); cudaFree(w.s_q); cudaFree(w.l_pack); cudaFree(w.l_q); cudaFree(w.chunks);
    cudaFree(w.ws_v); cudaFree(w.ws_begin); cudaFree(w.ws_inc); cudaFree(w.ws_rem);
    cudaFree(w.blk_a); cudaFree(w.blk_b); cudaFree(w.blk_c); cudaFree(w.out_f64); cudaFree(w.hist);
    cudaFree(w.sel_ids); cudaFree(w.sel_vals); cudaFree(w.ctl); cudaFree(w.dbg);
    if (w.ctl_host) cudaFreeHost(w.ctl_host);
    memset(&w, 0, sizeof w);
}

static int32_t ws_alloc(ppr_ctx* ctx, Workspace& w, i64 n, i64 chunk_cap) {
    memset(&w, 0, sizeof w);
    w.n = n;
    w.chunk_cap = chunk_cap;
    size_t N = (size_t)n;
    CUDA_TRY(dalloc(&w.r, N)); CUDA_TRY(dalloc(&w.reserve, N)); CUDA_TRY(dalloc(&w.est, N));
    CUDA_TRY(dalloc(&w.front[0], N)); CUDA_TRY(dalloc(&w.front[1], N));
    CUDA_TRY(dalloc(&w.q[0], N)); CUDA_TRY(dalloc(&w.q[1], N));
    CUDA_TRY(dalloc(&w.s_pack, N)); CUDA_TRY(dalloc(&w.s_q, N)); CUDA_TRY(dalloc(&w.l_pack, std::max(N, (size_t)chunk_cap))); CUDA_TRY(dalloc(&w.l_q, std::max(N, (size_t)chunk_cap)));
    CUDA_TRY(dalloc(&w.chunks, (size_t)chunk_cap));
    CUDA_TRY(dalloc(&w.ws_v, N)); CUDA_TRY(dalloc(&w.ws_begin, N + 64)); CUDA_TRY(dalloc(&w.ws_inc, N)); CUDA_TRY(dalloc(&w.ws_rem, N));
    int occ = 0;
    if (PUSH_DYN_SMEM) CUDA_TRY(cudaFuncSetAttribute(k_push_query, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PUSH_DYN_SMEM));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_push_query, PUSH_THREADS, PUSH_DYN_SMEM));
    int occ2 = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_topk_select, SEL_THREADS, 0));
    int occ3 = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ3, k_bwd_batch, PUSH_THREADS, 0));
    if (occ2 < 1) return ctx_fail(ctx, PPR_E_CUDA, "cooperative kernels do not fit on an SM");
    w.sel_blocks = ctx->sm_count * std::min(occ2, 4);
    occ = std::min(occ, occ3);
    if (occ < 1) return ctx_fail(ctx, PPR_E_CUDA, "cooperative kernels do not fit on an SM");
    w.coop_blocks = ctx->sm_count * std::min(occ, 4);
    CUDA_TRY(dalloc(&w.blk_a, (size_t)w.coop_blocks)); CUDA_TRY(dalloc(&w.blk_b, (size_t)w.coop_blocks));
    CUDA_TRY(dalloc(&w.blk_c, (size_t)w.coop_blocks * (PUSH_THREADS / 32)));   // also the per-warp counts of the reverse push
    CUDA_TRY(dalloc(&w.out_f64, N));
    CUDA_TRY(dalloc(&w.hist, (size_t)9 * 256));
    w.sel_cap = std::min<i64>(n, (i64)1 << 22);
    CUDA_TRY(dalloc(&w.sel_ids, (size_t)w.sel_cap)); CUDA_TRY(dalloc(&w.sel_vals, (size_t)w.sel_cap));
    CUDA_TRY(dalloc(&w.dbg, (size_t)DBG_ITERS * DBG_FIELDS));
    CUDA_TRY(dalloc(&w.ctl, 1));
    CUDA_TRY(cudaMemset(w.ctl, 0, sizeof(PushCtl)));
    CUDA_TRY(cudaHostAlloc((void**)&w.ctl_host, sizeof(PushCtl), cudaHostAllocDefault));
    memset(w.ctl_host, 0, sizeof(PushCtl));
    return PPR_OK;
}

// ============================================================================================= backward workspace
struct BwdWorkspace {
    i32 T; i64 n;
    u64* r; u64* reserve; u64* front[2]; i64 front_cap;
    u64* s_pack; u64* s_y; u32* s_t; u64* l_pack; u64* l_y; u32* l_t;
    uint2* chunks; i64 chunk_cap;
    u32* out_v; u32* out_t; u64* out_val; i64 out_cap;
};
static void bws_free(BwdWorkspace& b) {
    cudaFree(b.r); cudaFree(b.reserve); cudaFree(b.front[0]); cudaFree(b.front[1]);
    cudaFree(b.s_pack); cudaFree(b.s_y); cudaFree(b.s_t); cudaFree(b.l_pack); cudaFree(b.l_y); cudaFree(b.l_t);
    cudaFree(b.chunks); cudaFree(b.out_v); cudaFree(b.out_t); cudaFree(b.out_val);
    memset(&b, 0, sizeof b);
}
static int32_t bws_alloc(ppr_ctx* ctx, BwdWorkspace& b, i32 T, i64 n, i64 chunks_per_target) {
    memset(&b, 0, sizeof b);
    b.T = T; b.n = n;
    size_t tn = (size_t)T * (size_t)n;
    b.front_cap = (i64)std::min<size_t>(tn, (size_t)1 << 24);
    // one iteration of one target can stage ceil(indeg / PUSH_CHUNK) descriptors for every long row of the graph
    // (chunks_per_target, from the in-degree histogram at export): dense graphs need more than T * n of them
    b.chunk_cap = (i64)std::max<size_t>(std::min<size_t>((size_t)T * (size_t)std::max<i64>(chunks_per_target, 1), (size_t)1 << 26), (size_t)std::max<i64>(chunks_per_target, 1));
    b.chunk_cap = std::max<i64>(b.chunk_cap, (i64)std::min<size_t>(tn, (size_t)1 << 24));
    b.out_cap = (i64)std::min<size_t>(tn, (size_t)1 << 25);
    CUDA_TRY(dalloc(&b.r, tn)); CUDA_TRY(dalloc(&b.reserve, tn));
    CUDA_TRY(dalloc(&b.front[0], (size_t)b.front_cap)); CUDA_TRY(dalloc(&b.front[1], (size_t)b.front_cap));
    CUDA_TRY(dalloc(&b.s_pack, (size_t)b.front_cap)); CUDA_TRY(dalloc(&b.s_y, (size_t)b.front_cap)); CUDA_TRY(dalloc(&b.s_t, (size_t)b.front_cap));
    CUDA_TRY(dalloc(&b.l_pack, (size_t)b.chunk_cap)); CUDA_TRY(dalloc(&b.l_y, (size_t)b.chunk_cap)); CUDA_TRY(dalloc(&b.l_t, (size_t)b.chunk_cap));   // indexed by a long row's first chunk slot
    CUDA_TRY(dalloc(&b.chunks, (size_t)b.chunk_cap));
    CUDA_TRY(dalloc(&b.out_v, (size_t)b.out_cap)); CUDA_TRY(dalloc(&b.out_t, (size_t)b.out_cap)); CUDA_TRY(dalloc(&b.out_val, (size_t)b.out_cap));
    return PPR_OK;
}

// workspace of the CTA-per-target reverse push (bwdcta.cuh): per-slot rows + the growing pair buffers
struct BwdCtaWorkspace {
    int slots, bd; i64 n;
    u64* r; u64* reserve; u32* front; u32* pushed; u32* touched; u64* st_pk; u64* st_y;
    u32* ticket; ull* counters;
    u64* key[2]; u64* val[2]; i64 cap;     // pairs (double-buffered for the radix sort)
    u32* hist; size_t hist_len;
};
static void bcw_free(BwdCtaWorkspace& b) {
    cudaFree(b.r); cudaFree(b.reserve); cudaFree(b.front); cudaFree(b.pushed); cudaFree(b.touched); cudaFree(b.st_pk); cudaFree(b.st_y);
    cudaFree(b.ticket); cudaFree(b.counters); cudaFree(b.key[0]); cudaFree(b.key[1]); cudaFree(b.val[0]); cudaFree(b.val[1]); cudaFree(b.hist);
    memset(&b, 0, sizeof b);
}
static int32_t bcw_alloc(ppr_ctx* ctx, BwdCtaWorkspace& b, int slots, int bd, i64 n) {
    memset(&b, 0, sizeof b);
    b.slots = slots; b.bd = bd; b.n = n;
    const size_t sn = (size_t)slots * (size_t)n;
    CUDA_TRY(dalloc(&b.r, sn)); CUDA_TRY(dalloc(&b.reserve, sn)); CUDA_TRY(dalloc(&b.front, 2 * sn));
    CUDA_TRY(dalloc(&b.pushed, sn)); CUDA_TRY(dalloc(&b.touched, sn)); CUDA_TRY(dalloc(&b.st_pk, sn)); CUDA_TRY(dalloc(&b.st_y, sn));
    CUDA_TRY(dalloc(&b.ticket, 1)); CUDA_TRY(dalloc(&b.counters, 16));
    CUDA_TRY(cudaMemsetAsync(b.r, 0, sn * 8, ctx->stream));          // rows are zero outside a search; every search cleans up after itself
    CUDA_TRY(cudaMemsetAsync(b.reserve, 0, sn * 8, ctx->stream));
    return PPR_OK;
}
static int32_t bcw_reserve_pairs(ppr_ctx* ctx, BwdCtaWorkspace& b, i64 need, i64 keep) {   // grow the pair buffers, keeping the first `keep` pairs of buffer 0
    if (need <= b.cap) return PPR_OK;
    const i64 cap = std::max<i64>(need, b.cap + b.cap / 2);
    u64 *k0 = nullptr, *v0 = nullptr, *k1 = nullptr, *v1 = nullptr;
    cudaFree(b.key[1]); cudaFree(b.val[1]); b.key[1] = nullptr; b.val[1] = nullptr;
    CUDA_TRY(dalloc(&k0, (size_t)cap)); CUDA_TRY(dalloc(&v0, (size_t)cap));
    if (keep > 0) {
        CUDA_TRY(cudaMemcpyAsync(k0, b.key[0], (size_t)keep * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(v0, b.val[0], (size_t)keep * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    cudaFree(b.key[0]); cudaFree(b.val[0]);
    b.key[0] = k0; b.val[0] = v0;
    CUDA_TRY(dalloc(&k1, (size_t)cap)); CUDA_TRY(dalloc(&v1, (size_t)cap));
    b.key[1] = k1; b.val[1] = v1; b.cap = cap;
    return PPR_OK;
}

// ============================================================================================= graph
struct GraphExtra {  // host-side bookkeeping kept beside ppr_graph
    bool ws1_ready;
    BwdWorkspace bws; bool bws_ready;
    BwdCtaWorkspace bcw; bool bcw_ready;
    int walk_blocks;
    double last_divisor;
    // ppr_fora_batch: events and the pinned control-block snapshots are created once and reused by every call
    std::vector<cudaEvent_t> batch_ev;      // 4 timing events per query (only when the caller asks for stats)
    cudaEvent_t done_ev[2], copied_ev[2];   // per workspace: result ready / result downloaded (no timing)
    // ring of export buffers: a query's doubles leave its workspace at once, so the next queries compute while up to BATCH_RING
    // downloads are queued (sources without out-edges finish in ~0.1 ms but still return a dense n-vector)
    double* ring[8]; cudaEvent_t ring_done[8], ring_copied[8]; bool ring_ready;
    bool batch_ev_ready;
    PushCtl* snaps; size_t snaps_cap;       // pinned
    i64 bwd_chunks_per_target;              // chunk descriptors one reverse-push target can need in one iteration
    // on-device error metrics (metrics.cuh): [0] ground truth, [1] estimate, as doubles; the last power-iteration result
    double* metric_vec[2]; bool metric_set[2];
    double* pm_result; bool last_is_pm;
    i32* sort_ids; u64* sort_vals; i64 sort_cap;   // ordered top-k survivors (k_rank_sort)
    std::vector<u32> h_indeg;                      // in-degrees (all-pair backward search: heaviest targets first)
    i32* bwd_order; i64 bwd_order_cap;             // device copy of a call's target order
    u64* edge_inv;                                 // [m] per in-edge reciprocal degree (lazily built by the all-pair search)
};
static std::vector<std::pair<ppr_graph*, GraphExtra*>> g_extras;
static std::mutex g_extras_mu;   // graphs of different contexts may be created / queried / destroyed from different threads
static GraphExtra* extra_of(ppr_graph* g) {
    std::lock_guard<std::mutex> lk(g_extras_mu);
    for (auto& p : g_extras) if (p.first == g) return p.second;
    return nullptr;
}

extern "C" int32_t ppr_graph_destroy(ppr_graph* g) {
    if (!g) return PPR_E_INVALID;
    ppr_ctx* ctx = g->ctx;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copy);
    GraphExtra* x = extra_of(g);
    ws_free(g->ws[0]);
    if (x && x->ws1_ready) ws_free(g->ws[1]);
    if (x && x->bws_ready) bws_free(x->bws);
    if (x && x->bcw_ready) bcw_free(x->bcw);
    if (x) {
        for (auto& e : x->batch_ev) cudaEventDestroy(e);
        if (x->batch_ev_ready) for (int i = 0; i < 2; i++) { cudaEventDestroy(x->done_ev[i]); cudaEventDestroy(x->copied_ev[i]); }
        if (x->snaps) cudaFreeHost(x->snaps);
        if (x->ring_ready) for (int i = 0; i < 8; i++) { cudaFree(x->ring[i]); cudaEventDestroy(x->ring_done[i]); cudaEventDestroy(x->ring_copied[i]); }
        cudaFree(x->metric_vec[0]); cudaFree(x->metric_vec[1]); cudaFree(x->pm_result); cudaFree(x->sort_ids); cudaFree(x->sort_vals); cudaFree(x->bwd_order); cudaFree(x->edge_inv);
    }
    cudaFree(g->d.out_vtx); cudaFree(g->d.col_idx); cudaFree(g->d.out_deg); cudaFree(g->d.row32);
    cudaFree(g->d.in_vtx); cudaFree(g->d.row_idx); cudaFree(g->d.invdeg); cudaFree(g->d.push_idx); cudaFree(g->d.hot_id);
    cudaFree(g->d.pull_idx); cudaFree(g->d.pull_hot_id); cudaFree(g->d.long_v); cudaFree(g->d.long_chunk); cudaFree(g->d.pull_task);
    {
        std::lock_guard<std::mutex> lk(g_extras_mu);
        for (size_t i = 0; i < g_extras.size(); i++) if (g_extras[i].first == g) { delete g_extras[i].second; g_extras.erase(g_extras.begin() + i); break; }
    }
    delete g;
    return PPR_OK;
}

extern "C" int32_t ppr_graph_create(ppr_ctx* ctx, int64_t n, int64_t m, const int64_t* row_ptr, const int32_t* col_idx,
                                    ppr_graph** out) {
    if (!ctx) return PPR_E_INVALID;
    if (!out || !row_ptr || (m > 0 && !col_idx)) return ctx_fail(ctx, PPR_E_INVALID, "ppr_graph_create: NULL argument");
    *out = nullptr;
    if (n < 1 || n > 0x7fffffffll) return ctx_fail(ctx, PPR_E_RANGE, "ppr_graph_create: need 1 <= n < 2^31");
    if (m < 0 || m >= VTX_MAX_EDGES) return ctx_fail(ctx, PPR_E_RANGE, "ppr_graph_create: need 0 <= m < 2^40");
    if (row_ptr[0] != 0 || row_ptr[n] != m) return ctx_fail(ctx, PPR_E_INVALID, "ppr_graph_create: row_ptr[0] != 0 or row_ptr[n] != m");
    i64 chunk_cap = 1;
    for (i64 v = 0; v < n; v++) {
        i64 d = row_ptr[v + 1] - row_ptr[v];
        if (d < 0) return ctx_fail(ctx, PPR_E_INVALID, "ppr_graph_create: row_ptr not monotone");
        if (d > VTX_MAX_DEG) return ctx_fail(ctx, PPR_E_RANGE, "ppr_graph_create: out-degree >= 2^24");
        if (d > PUSH_SMALL_MAX) chunk_cap += (d + PUSH_CHUNK - 1) / PUSH_CHUNK;
    }
    cudaSetDevice(ctx->device);
    ppr_graph* g = new (std::nothrow) ppr_graph();
    GraphExtra* x = new (std::nothrow) GraphExtra();
    if (!g || !x) { delete g; delete x; return ctx_fail(ctx, PPR_E_NOMEM, "ppr_graph_create: host allocation failed"); }
    memset(&g->d, 0, sizeof g->d); memset(g->ws, 0, sizeof g->ws);
    g->ctx = ctx; g->last_ws = 0; g->last_kind = 0; g->trace_active = false;
    x->ws1_ready = false; x->bws_ready = false; x->walk_blocks = 0; x->last_divisor = 0.0; memset(&x->bws, 0, sizeof x->bws);
    x->bcw_ready = false; memset(&x->bcw, 0, sizeof x->bcw);
    x->metric_vec[0] = x->metric_vec[1] = nullptr; x->metric_set[0] = x->metric_set[1] = false; x->pm_result = nullptr; x->last_is_pm = false; x->sort_ids = nullptr; x->sort_vals = nullptr; x->sort_cap = 0; x->bwd_order = nullptr; x->bwd_order_cap = 0; x->edge_inv = nullptr;
    x->batch_ev_ready = false; x->snaps = nullptr; x->snaps_cap = 0; x->bwd_chunks_per_target = 1;
    x->ring_ready = false; for (int i = 0; i < 8; i++) x->ring[i] = nullptr;
    { std::lock_guard<std::mutex> lk(g_extras_mu); g_extras.push_back(std::make_pair(g, x)); }
    g->d.n = n; g->d.m = m;
    i64* d_ptr = nullptr;
    int32_t rc = PPR_OK;
    auto fail = [&](int32_t code) { cudaFree(d_ptr); ppr_graph_destroy(g); return code; };
#define G_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return fail(ctx_fail_cuda(ctx, e__, #expr, __FILE__, __LINE__)); } while (0)
    G_TRY(dalloc(&g->d.out_vtx, (size_t)n)); G_TRY(dalloc(&g->d.out_deg, (size_t)n)); G_TRY(dalloc(&g->d.invdeg, (size_t)n));
    G_TRY(dalloc(&g->d.in_vtx, (size_t)n)); G_TRY(dalloc(&g->d.col_idx, (size_t)m)); G_TRY(dalloc(&g->d.row_idx, (size_t)m));
    G_TRY(dalloc(&d_ptr, (size_t)n + 1));
    G_TRY(dalloc(&g->d.push_idx, (size_t)m)); G_TRY(dalloc(&g->d.hot_id, (size_t)HOT_SLOTS));
    int blocks = ctx->sm_count * 8;
    G_TRY(cudaMemcpyAsync(d_ptr, row_ptr, ((size_t)n + 1) * sizeof(i64), cudaMemcpyHostToDevice, ctx->stream));
    k_build_vtx<<<blocks, 256, 0, ctx->stream>>>(d_ptr, g->d.out_vtx, g->d.out_deg, g->d.invdeg, n);
    G_TRY(cudaGetLastError());
    if (m < ((i64)1 << 32)) {
        G_TRY(dalloc(&g->d.row32, (size_t)n + 1));
        k_build_row32<<<blocks, 256, 0, ctx->stream>>>(d_ptr, g->d.row32, n);
        G_TRY(cudaGetLastError());
    }
    if (m > 0) G_TRY(cudaMemcpyAsync(g->d.col_idx, col_idx, (size_t)m * sizeof(i32), cudaMemcpyHostToDevice, ctx->stream));

    // CSC on the device: histogram -> scan -> stable sort by target (temporaries freed at the end of this block)
    ull* d_indeg = nullptr; u32* d_err = nullptr; i32* d_src = nullptr; u32* d_keys = nullptr; void* d_tmp = nullptr; i32* d_slot = nullptr;
    auto free_tmp = [&]() { cudaFree(d_indeg); cudaFree(d_err); cudaFree(d_src); cudaFree(d_keys); cudaFree(d_tmp); cudaFree(d_slot);
                            d_indeg = nullptr; d_err = nullptr; d_src = nullptr; d_keys = nullptr; d_tmp = nullptr; d_slot = nullptr; };
#define B_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { free_tmp(); return fail(ctx_fail_cuda(ctx, e__, #expr, __FILE__, __LINE__)); } } while (0)
    B_TRY(dalloc(&d_indeg, (size_t)n + 1)); B_TRY(dalloc(&d_err, 1));
    B_TRY(cudaMemsetAsync(d_indeg, 0, ((size_t)n + 1) * sizeof(ull), ctx->stream));
    B_TRY(cudaMemsetAsync(d_err, 0, sizeof(u32), ctx->stream));
    if (m > 0) { k_indeg<<<blocks, 256, 0, ctx->stream>>>(g->d.col_idx, m, n, d_indeg, d_err); B_TRY(cudaGetLastError()); }
    k_check_indeg<<<blocks, 256, 0, ctx->stream>>>(d_indeg, n, d_err);
    B_TRY(cudaGetLastError());
    u32 h_err = 0;
    std::vector<ull> h_indeg;
    try { h_indeg.resize((size_t)n); } catch (...) { free_tmp(); return fail(ctx_fail(ctx, PPR_E_NOMEM, "ppr_graph_create: host allocation failed")); }
    B_TRY(cudaMemcpyAsync(&h_err, d_err, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    B_TRY(cudaMemcpyAsync(h_indeg.data(), d_indeg, (size_t)n * sizeof(ull), cudaMemcpyDeviceToHost, ctx->stream));
    B_TRY(cudaStreamSynchronize(ctx->stream));
    if (h_err == 1u) { free_tmp(); return fail(ctx_fail(ctx, PPR_E_INVALID, "ppr_graph_create: col_idx out of range")); }
    if (h_err == 2u) { free_tmp(); return fail(ctx_fail(ctx, PPR_E_RANGE, "ppr_graph_create: in-degree >= 2^24")); }
    {   // column pointers: exclusive scan of the n + 1 counts (the last one is 0, so entry n receives m)
        size_t tb = 0;
        B_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb, d_indeg, reinterpret_cast<ull*>(d_ptr), n + 1, ctx->stream));
        B_TRY(cudaMalloc(&d_tmp, std::max<size_t>(tb, 16)));
        B_TRY(cub::DeviceScan::ExclusiveSum(d_tmp, tb, d_indeg, reinterpret_cast<ull*>(d_ptr), n + 1, ctx->stream));
        k_build_vtx<<<blocks, 256, 0, ctx->stream>>>(d_ptr, g->d.in_vtx, nullptr, nullptr, n);
        B_TRY(cudaGetLastError());
        B_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_tmp); d_tmp = nullptr;
    }
    if (m > 0) {   // rows of the CSC: sources of the edges, stably sorted by target
        B_TRY(dalloc(&d_src, (size_t)m)); B_TRY(dalloc(&d_keys, (size_t)m));
        k_edge_src<<<blocks, 256, 0, ctx->stream>>>(g->d.out_vtx, n, d_src);
        B_TRY(cudaGetLastError());
        int bits = 1;
        while (bits < 31 && ((i64)1 << bits) < n) bits++;
        size_t tb = 0;
        B_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, reinterpret_cast<const u32*>(g->d.col_idx), d_keys, d_src, g->d.row_idx, (i64)m, 0, bits, ctx->stream));
        B_TRY(cudaMalloc(&d_tmp, std::max<size_t>(tb, 16)));
        B_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tb, reinterpret_cast<const u32*>(g->d.col_idx), d_keys, d_src, g->d.row_idx, (i64)m, 0, bits, ctx->stream));
        B_TRY(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_tmp); d_tmp = nullptr; cudaFree(d_src); d_src = nullptr; cudaFree(d_keys); d_keys = nullptr;
    }
    // hub set: the (up to) HOT_SLOTS vertices of highest in-degree (ties: lower id first), at least hot_min_indeg in-edges
    // each -- an O(n) selection on the host from the downloaded histogram; the O(m) encoding runs on the device
    std::vector<u32> hot_id;
    try {
        std::vector<std::pair<i64, i32>> cand;
        const i64 min_in = std::max<i64>(1, ctx->hot_min_indeg);
        for (i64 v = 0; v < n; v++) { i64 d = (i64)h_indeg[(size_t)v]; if (d >= min_in) cand.push_back(std::make_pair(-d, (i32)v)); }
        i64 bch = 1;
        for (i64 v = 0; v < n; v++) { i64 d = (i64)h_indeg[(size_t)v]; if (d > PUSH_SMALL_MAX) bch += (d + PUSH_CHUNK - 1) / PUSH_CHUNK; }
        x->bwd_chunks_per_target = bch;
        x->h_indeg.resize((size_t)n);
        for (i64 v = 0; v < n; v++) x->h_indeg[(size_t)v] = (u32)std::min<ull>(h_indeg[(size_t)v], 0xffffffffull);
        size_t keep = std::min<size_t>(cand.size(), HOT_SLOTS);
        std::partial_sort(cand.begin(), cand.begin() + keep, cand.end());
        for (size_t k = 0; k < keep; k++) hot_id.push_back((u32)cand[k].second);
    } catch (...) { free_tmp(); return fail(ctx_fail(ctx, PPR_E_NOMEM, "ppr_graph_create: host allocation failed")); }
    g->d.n_hot = (u32)hot_id.size();
    if (!hot_id.empty()) B_TRY(cudaMemcpyAsync(g->d.hot_id, hot_id.data(), hot_id.size() * sizeof(u32), cudaMemcpyHostToDevice, ctx->stream));
    if (m > 0) {
        B_TRY(dalloc(&d_slot, (size_t)n));
        B_TRY(cudaMemsetAsync(d_slot, 0xff, (size_t)n * sizeof(i32), ctx->stream));   // -1 everywhere
        if (!hot_id.empty()) { k_set_slots<<<(unsigned)((hot_id.size() + 255) / 256), 256, 0, ctx->stream>>>(g->d.hot_id, g->d.n_hot, d_slot); B_TRY(cudaGetLastError()); }
        k_encode_push_idx<<<blocks, 256, 0, ctx->stream>>>(g->d.col_idx, d_slot, m, g->d.push_idx);
        B_TRY(cudaGetLastError());
    }
    // pull-direction iterations (pull.cuh): hub SOURCES (highest out-degree) encoded in a copy of the CSC, the list of
    // long rows cut into chunks, and a split of the 32-vertex groups into tasks with equal numbers of in-edges
    {
        std::vector<u32> phot, longv, task;
        std::vector<uint2> lchunk;
        try {
            std::vector<std::pair<i64, i32>> cand;
            for (i64 v = 0; v < n; v++) { i64 d = row_ptr[v + 1] - row_ptr[v]; if (d >= 64) cand.push_back(std::make_pair(-d, (i32)v)); }
            size_t keep = std::min<size_t>(cand.size(), PULL_SLOTS);
            std::partial_sort(cand.begin(), cand.begin() + keep, cand.end());
            for (size_t k = 0; k < keep; k++) phot.push_back((u32)cand[k].second);
            const i64 groups = (n + 31) / 32;
            std::vector<i64> gw((size_t)groups, 0);
            i64 total_w = 0;
            for (i64 v = 0; v < n; v++) {
                const i64 d = (i64)h_indeg[(size_t)v];
                if (d > PULL_LONG) {
                    longv.push_back((u32)v);
                    for (i64 c = 0; c * PULL_LCHUNK < d; c++) lchunk.push_back(make_uint2((u32)v, (u32)c));
                } else gw[(size_t)(v >> 5)] += d;
            }
            for (i64 gidx = 0; gidx < groups; gidx++) { gw[(size_t)gidx] += 96; total_w += gw[(size_t)gidx]; }   // + the fixed cost of a group
            const i64 nt = std::max<i64>(1, std::min<i64>(groups, 65536));
            task.assign((size_t)nt + 1, (u32)groups);
            task[0] = 0;
            i64 acc = 0, t = 1;
            for (i64 gidx = 0; gidx < groups && t < nt; gidx++) {
                acc += gw[(size_t)gidx];
                while (t < nt && acc * nt >= total_w * t) { task[(size_t)t] = (u32)(gidx + 1); t++; }
            }
        } catch (...) { free_tmp(); return fail(ctx_fail(ctx, PPR_E_NOMEM, "ppr_graph_create: host allocation failed")); }
        g->d.n_pull_hot = (u32)phot.size(); g->d.n_long = (u32)longv.size(); g->d.n_long_chunk = (u32)lchunk.size(); g->d.n_pull_task = (u32)task.size() - 1;
        B_TRY(dalloc(&g->d.pull_hot_id, phot.size())); B_TRY(dalloc(&g->d.long_v, longv.size())); B_TRY(dalloc(&g->d.long_chunk, lchunk.size()));
        B_TRY(dalloc(&g->d.pull_task, task.size())); B_TRY(dalloc(&g->d.pull_idx, (size_t)m));
        if (!phot.empty()) B_TRY(cudaMemcpyAsync(g->d.pull_hot_id, phot.data(), phot.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (!longv.empty()) B_TRY(cudaMemcpyAsync(g->d.long_v, longv.data(), longv.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (!lchunk.empty()) B_TRY(cudaMemcpyAsync(g->d.long_chunk, lchunk.data(), lchunk.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream));
        B_TRY(cudaMemcpyAsync(g->d.pull_task, task.data(), task.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        if (m > 0) {
            B_TRY(cudaMemsetAsync(d_slot, 0xff, (size_t)n * sizeof(i32), ctx->stream));
            if (!phot.empty()) { k_set_slots<<<(unsigned)((phot.size() + 255) / 256), 256, 0, ctx->stream>>>(g->d.pull_hot_id, g->d.n_pull_hot, d_slot); B_TRY(cudaGetLastError()); }
            k_encode_push_idx<<<blocks, 256, 0, ctx->stream>>>(g->d.row_idx, d_slot, m, g->d.pull_idx);
            B_TRY(cudaGetLastError());
        }
        B_TRY(cudaStreamSynchronize(ctx->stream));   // the host vectors go out of scope
    }
    B_TRY(cudaStreamSynchronize(ctx->stream));
    free_tmp();
#undef B_TRY
    cudaFree(d_ptr); d_ptr = nullptr;
#undef G_TRY
    rc = ws_alloc(ctx, g->ws[0], n, chunk_cap);
    if (rc != PPR_OK) { ppr_graph_destroy(g); return rc; }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_walk<2, WALK_PER_LANE>, WALK_THREADS, 0) != cudaSuccess || occ < 1) occ = 1;
    x->walk_blocks = ctx->sm_count * occ;
    *out = g;
    return PPR_OK;
}

extern "C" int64_t ppr_graph_node_count(const ppr_graph* g) { return g ? g->d.n : -1; }
extern "C" int64_t ppr_graph_rel_count(const ppr_graph* g) { return g ? g->d.m : -1; }

__global__ void k_unpack_ptr(const u64* __restrict__ vtx, i64* __restrict__ ptr, i64 n, i64 m) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 v = gtid; v < n; v += gsize) ptr[v] = (i64)vtx_begin(vtx[v]);
    if (gtid == 0) ptr[n] = m;
}
extern "C" int32_t ppr_graph_export_csc(ppr_graph* g, int64_t* col_ptr, int32_t* row_idx) {
    if (!g || !col_ptr || !row_idx) return PPR_E_INVALID;
    ppr_ctx* ctx = g->ctx;
    cudaSetDevice(ctx->device);
    i64* d_ptr = nullptr;
    CUDA_TRY(dalloc(&d_ptr, (size_t)g->d.n + 1));
    k_unpack_ptr<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(g->d.in_vtx, d_ptr, g->d.n, g->d.m);
    cudaError_t e = cudaMemcpyAsync(col_ptr, d_ptr, ((size_t)g->d.n + 1) * sizeof(i64), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && g->d.m) e = cudaMemcpyAsync(row_idx, g->d.row_idx, (size_t)g->d.m * sizeof(i32), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_ptr);
    CUDA_TRY(e);
    return PPR_OK;
}

// ============================================================================================= launch helpers
static void fill_push_args(PushArgs& A, ppr_graph* g, Workspace& w) {
    memset(&A, 0, sizeof A);
    A.n = g->d.n; A.out_vtx = g->d.out_vtx; A.col_idx = g->d.col_idx; A.out_deg = g->d.out_deg;
    A.push_idx = g->d.push_idx; A.hot_id = g->d.hot_id; A.n_hot = g->d.n_hot;
    A.r = w.r; A.reserve = w.reserve; A.est = w.est; A.front[0] = w.front[0]; A.front[1] = w.front[1];
    A.s_pack = w.s_pack; A.s_q = w.s_q; A.l_pack = w.l_pack; A.l_q = w.l_q; A.chunks = w.chunks; A.mini = getenv("PPR_NO_MINI") ? 0 : 1;
    A.ws_v = w.ws_v; A.ws_begin = w.ws_begin; A.ws_inc = w.ws_inc; A.ws_rem = w.ws_rem;
    A.blk_a = w.blk_a; A.blk_b = w.blk_b; A.ctl = w.ctl; A.dbg = w.dbg;
    A.prep_adds_push_part = 1;
    A.scan_threshold = g->ctx->scan_threshold < 0 ? std::max<i64>(1, g->d.n / 4) : g->ctx->scan_threshold;
    A.midi_ctas = g->ctx->midi_ctas; A.midi_v = (u32)g->ctx->midi_v; A.midi_e = (u64)g->ctx->midi_e;
    A.in_vtx = g->d.in_vtx; A.pull_idx = g->d.pull_idx; A.pull_hot_id = g->d.pull_hot_id; A.n_pull_hot = g->d.n_pull_hot;
    A.long_v = g->d.long_v; A.n_long = g->d.n_long; A.long_chunk = g->d.long_chunk; A.n_long_chunk = g->d.n_long_chunk;
    A.pull_task = g->d.pull_task; A.n_pull_task = g->d.n_pull_task; A.q[0] = w.q[0]; A.q[1] = w.q[1];
    // direction switch: a pull pass costs about as much as pushing a quarter of all edges
    const double pe = g->ctx->pull_enter, px = std::min(g->ctx->pull_exit, g->ctx->pull_enter);
    A.pull_enter = (pe < 0 || g->d.m == 0) ? -1 : (i64)(pe * (double)g->d.m);
    A.pull_exit = (pe < 0 || g->d.m == 0) ? 0 : (i64)(px * (double)g->d.m);
}

static cudaError_t launch_push(ppr_ctx* ctx, Workspace& w, PushArgs& A, cudaStream_t st = nullptr, int blocks = 0) {
    void* args[] = {&A};
    return cudaLaunchCooperativeKernel((const void*)k_push_query, dim3(blocks ? blocks : w.coop_blocks), dim3(PUSH_THREADS), args, PUSH_DYN_SMEM, st ? st : ctx->stream);
}

static cudaError_t launch_walk(ppr_ctx* ctx, ppr_graph* g, Workspace& w, u64* target, u64 seed, u64 tag, double alpha, i32 src,
                               int restart_at_source, int no_zero_hop, u64 shard, u64 nshards, cudaStream_t st = nullptr) {
    WalkArgs W;
    memset(&W, 0, sizeof W);
    W.out_vtx = g->d.out_vtx; W.col_idx = g->d.col_idx; W.ws_v = w.ws_v; W.ws_begin = w.ws_begin; W.ws_inc = w.ws_inc;
    W.ws_rem = w.ws_rem; W.target = target; W.ctl = w.ctl; W.shard = shard; W.nshards = nshards;
    u64 key = h_walk_key(seed, tag);
    W.k0 = (u32)key; W.k1 = (u32)(key >> 32);
    W.tstop = h_stop_threshold(alpha);
    W.src = src; W.restart_at_source = restart_at_source; W.no_zero_hop = no_zero_hop;
    W.row32 = g->d.row32;
    // graph beyond the L2: eviction hints, and the half-size 32-bit row pointers as the vertex table when they exist
    int variant = ctx->walk_hint >= 0 ? ctx->walk_hint : (4.0 * (double)g->d.m + 8.0 * (double)g->d.n > 96.0 * 1024 * 1024 ? 2 : 0);
    if (variant >= 2 && !g->d.row32) variant = 1;
    if (variant == 3) k_walk<3, WALK_PER_LANE><<<extra_of(g)->walk_blocks, WALK_THREADS, 0, st ? st : ctx->stream>>>(W);
    else if (variant == 4) k_walk<4, WALK_PER_LANE><<<extra_of(g)->walk_blocks, WALK_THREADS, 0, st ? st : ctx->stream>>>(W);
    else if (variant == 5) k_walk<5, WALK_PER_LANE><<<extra_of(g)->walk_blocks, WALK_THREADS, 0, st ? st : ctx->stream>>>(W);
    else if (variant == 2) k_walk<2, WALK_PER_LANE><<<extra_of(g)->walk_blocks, WALK_THREADS, 0, st ? st : ctx->stream>>>(W);
    else if (variant == 1) k_walk<1, WALK_PER_LANE><<<extra_of(g)->walk_blocks, WALK_THREADS, 0, st ? st : ctx->stream>>>(W);
    else k_walk<0, WALK_PER_LANE><<<extra_of(g)->walk_blocks, WALK_THREADS, 0, st ? st : ctx->stream>>>(W);
    return cudaGetLastError();
}

static void stats_from_ctl(ppr_stats* st, const PushCtl* c, double omega) {
    if (!st) return;
    memset(st, 0, sizeof *st);
    st->pushes = (i64)c->pushes; st->edges = (i64)c->edges; st->iters = (i64)c->iters; st->rounds = (i64)c->rounds;
    st->walks = (i64)c->walks_done; st->walks_total = (i64)c->total_walks; st->steps = (i64)c->steps;
    st->walk_sources = (i64)c->n_sources; st->rmax_fx = c->rmax_fx; st->rsum_fx = c->rsum_fx; st->omega = omega;
}

static int32_t check_query(ppr_graph* g, i32 v, double alpha, const char* who) {
    if (!g) return PPR_E_INVALID;
    if (v < 0 || v >= g->d.n) return ctx_fail(g->ctx, PPR_E_INVALID, std::string(who) + ": vertex id out of range");
    if (!(alpha > 0.0 && alpha < 1.0)) return ctx_fail(g->ctx, PPR_E_INVALID, std::string(who) + ": alpha must be in (0, 1)");
    return PPR_OK;
}

static int32_t export_vec(ppr_graph* g, Workspace& w, const u64* src, double divisor, double* out, cudaStream_t s) {
    ppr_ctx* ctx = g->ctx;
    k_export_f64<<<ctx->sm_count * 8, 256, 0, s>>>(src, w.out_f64, g->d.n, divisor);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, w.out_f64, (size_t)g->d.n * sizeof(double), cudaMemcpyDeviceToHost, s));
    return PPR_OK;
}

static int32_t read_ctl(ppr_graph* g, Workspace& w) {
    ppr_ctx* ctx = g->ctx;
    CUDA_TRY(cudaMemcpyAsync(w.ctl_host, w.ctl, sizeof(PushCtl), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PPR_OK;
}

// Push/walk balance constant.  A walk costs ~6 relaxed edges while col_idx and the vertex table fit in the
// 126 MB L2 (scale 20: 0.10 ns per walk vs 0.013 ns per edge) and ~15-25 once walks miss to HBM (scale 22:
// 0.14 ns vs 0.0095 ns); measured sweep in profiles/.  ppr_set_balance(walk_cost >= 0) overrides.
static double effective_walk_cost(const ppr_graph* g) {
    if (g->ctx->walk_cost >= 0) return g->ctx->walk_cost;
    double footprint = 4.0 * (double)g->d.m + 8.0 * (double)g->d.n;
    return footprint > 96.0 * 1024 * 1024 ? 24.0 : 6.0;
}

struct ForaParams { double delta, pfail, L, rmax, omega; };
// J/Algo_Conf.java:45-53 and J/Fora_Whole_Graph.java:86-87
static ForaParams fora_params(i64 n, i64 m, double alpha, double eps, const ppr_conf* conf) {
    ForaParams p;
    p.delta = 1.0 / (double)n; p.pfail = 1.0 / (double)n;
    if (conf && conf->delta > 0) p.delta = conf->delta;
    if (conf && conf->pfail > 0) p.pfail = conf->pfail;
    p.L = std::log(2.0 / p.pfail);
    p.rmax = eps * std::sqrt(p.delta / 3.0 / (double)m / p.L) / (1.0 - alpha);
    p.omega = (eps + 2.0) * p.L / eps / eps / p.delta;
    return p;
}

// ============================================================================================= forward push
extern "C" int32_t ppr_forward_push(ppr_graph* g, int32_t src, double alpha, double rmax, double* reserve_out,
                                    double* residue_out, double* rsum_out, ppr_stats* stats) {
    int32_t rc = check_query(g, src, alpha, "ppr_forward_push");
    if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    if (!(rmax > 0)) return ctx_fail(ctx, PPR_E_INVALID, "ppr_forward_push: rmax must be > 0");
    cudaSetDevice(ctx->device);
    Workspace& w = g->ws[0];
    g->trace_active = false;
    PushArgs A;
    fill_push_args(A, g, w);
    A.src = src; A.alpha_q64 = h_alpha_q64(alpha); A.alpha = alpha; A.rmax0 = rmax; A.mode = PMODE_PUSH; A.init = 1;
    A.max_rounds = 1;
    CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    CUDA_TRY(launch_push(ctx, w, A));
    CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
    int launches = 1;
    if (reserve_out) { rc = export_vec(g, w, w.reserve, 0.0, reserve_out, ctx->stream); if (rc) return rc; launches++; }
    if (residue_out) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // out_f64 is reused
        rc = export_vec(g, w, w.r, 0.0, residue_out, ctx->stream); if (rc) return rc; launches++;
    }
    rc = read_ctl(g, w); if (rc) return rc;
    g->last_ws = 0; g->last_kind = 1; extra_of(g)->last_divisor = 0.0; extra_of(g)->last_is_pm = false;
    if (rsum_out) *rsum_out = (double)(FX_ONE - w.ctl_host->reserve_total) * FX_INV;
    if (stats) {
        stats_from_ctl(stats, w.ctl_host, 0.0);
        cudaEventElapsedTime(&stats->ms_push, ctx->ev[0], ctx->ev[1]);
        stats->ms_total = stats->ms_push; stats->launches = launches;
        stats->rsum_fx = FX_ONE - w.ctl_host->reserve_total;
    }
    return PPR_OK;
}

// ============================================================================================= FORA
static int32_t fora_launch(ppr_graph* g, Workspace& w, i32 src, double alpha, double eps, u64 seed, int flags,
                           const ppr_conf* conf, i32 shard, i32 nshards, cudaEvent_t e0, cudaEvent_t e1, cudaEvent_t e2, double* omega_out,
                           cudaStream_t st = nullptr, int blocks = 0) {
    ppr_ctx* ctx = g->ctx;
    ForaParams p = fora_params(g->d.n, g->d.m, alpha, eps, conf);
    PushArgs A;
    fill_push_args(A, g, w);
    A.src = src; A.alpha_q64 = h_alpha_q64(alpha); A.alpha = alpha; A.omega = p.omega; A.walk_cost = effective_walk_cost(g);
    if (flags & PPR_BALANCE_PER_SHARD) A.walk_cost = A.walk_cost / (double)nshards;   // walks are nshards times cheaper than on one GPU
    A.rmax0 = p.rmax; A.max_rounds = ctx->max_rounds; A.mode = PMODE_FORA; A.init = 1;
    A.zero_reserve_after_push = shard != 0; A.prep_adds_push_part = shard == 0;
    if (omega_out) *omega_out = p.omega;
    if (!st) st = ctx->stream;
    if (e0) CUDA_TRY(cudaEventRecord(e0, st));
    CUDA_TRY(launch_push(ctx, w, A, st, blocks));
    if (e1) CUDA_TRY(cudaEventRecord(e1, st));
    CUDA_TRY(launch_walk(ctx, g, w, w.reserve, seed, 0, alpha, src, (flags & PPR_RESTART_AT_SOURCE) ? 1 : 0, 1, (u64)shard, (u64)nshards, st));
    if (e2) CUDA_TRY(cudaEventRecord(e2, st));
    return PPR_OK;
}

static int32_t check_fora(ppr_graph* g, i32 src, double alpha, double eps, i32 shard, i32 nshards, const char* who) {
    int32_t rc = check_query(g, src, alpha, who);
    if (rc) return rc;
    if (!(eps > 0)) return ctx_fail(g->ctx, PPR_E_INVALID, std::string(who) + ": eps must be > 0");
    if (nshards < 1 || shard < 0 || shard >= nshards) return ctx_fail(g->ctx, PPR_E_INVALID, std::string(who) + ": bad shard");
    if (g->d.m < 1) return ctx_fail(g->ctx, PPR_E_INVALID, std::string(who) + ": graph has no edges");
    return PPR_OK;
}

extern "C" int32_t ppr_fora_shard(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed, int32_t flags,
                                  const ppr_conf* conf, int32_t shard, int32_t nshards, ppr_stats* stats) {
    int32_t rc = check_fora(g, src, alpha, eps, shard, nshards, "ppr_fora");
    if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    cudaSetDevice(ctx->device);
    Workspace& w = g->ws[0];
    g->trace_active = false;
    double omega = 0;
    rc = fora_launch(g, w, src, alpha, eps, seed, flags, conf, shard, nshards, ctx->ev[0], ctx->ev[1], ctx->ev[2], &omega);
    if (rc) return rc;
    rc = read_ctl(g, w); if (rc) return rc;
    g->last_ws = 0; g->last_kind = 1; extra_of(g)->last_divisor = 0.0; extra_of(g)->last_is_pm = false;
    if (stats) {
        stats_from_ctl(stats, w.ctl_host, omega);
        cudaEventElapsedTime(&stats->ms_push, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&stats->ms_walk, ctx->ev[1], ctx->ev[2]);
        cudaEventElapsedTime(&stats->ms_total, ctx->ev[0], ctx->ev[2]);
        stats->launches = 2;
    }
    return PPR_OK;
}

extern "C" int32_t ppr_fora(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed, int32_t flags,
                            const ppr_conf* conf, double* ppr_out, ppr_stats* stats) {
    if (g && !ppr_out) return ctx_fail(g->ctx, PPR_E_INVALID, "ppr_fora: ppr_out is NULL");
    int32_t rc = ppr_fora_shard(g, src, alpha, eps, seed, flags, conf, 0, 1, stats);
    if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    rc = export_vec(g, g->ws[0], g->ws[0].reserve, 0.0, ppr_out, ctx->stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (stats) stats->launches += 1;
    return PPR_OK;
}

extern "C" int32_t ppr_fora_batch(ppr_graph* g, const int32_t* srcs, int32_t nq, double alpha, double eps, uint64_t seed,
                                  int32_t flags, const ppr_conf* conf, double* ppr_out, ppr_stats* stats_out) {
    if (!g) return PPR_E_INVALID;
    ppr_ctx* ctx = g->ctx;
    if (!srcs || nq < 0) return ctx_fail(ctx, PPR_E_INVALID, "ppr_fora_batch: NULL argument");
    const bool resident = ppr_out == nullptr;  // results stay on the device (only the last query of each workspace survives)
    for (int q = 0; q < nq; q++) { int32_t rc = check_fora(g, srcs[q], alpha, eps, 0, 1, "ppr_fora_batch"); if (rc) return rc; }
    cudaSetDevice(ctx->device);
    GraphExtra* x = extra_of(g);
    if (!x->ws1_ready) {
        int32_t rc = ws_alloc(ctx, g->ws[1], g->d.n, g->ws[0].chunk_cap);
        if (rc) { ws_free(g->ws[1]); return rc; }
        x->ws1_ready = true;
    }
    g->trace_active = false;
    const i64 n = g->d.n;
    // events and snapshot memory live with the graph: nothing is created per call once they have grown to the batch size
    if (!x->batch_ev_ready) {
        for (int i = 0; i < 2; i++) {
            CUDA_TRY(cudaEventCreateWithFlags(&x->done_ev[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&x->copied_ev[i], cudaEventDisableTiming));
        }
        x->batch_ev_ready = true;
    }
    constexpr int BATCH_RING = 6;
    if (!resident && !x->ring_ready) {
        for (int i = 0; i < 8; i++) { x->ring[i] = nullptr; }
        for (int i = 0; i < BATCH_RING; i++) {
            CUDA_TRY(dalloc(&x->ring[i], (size_t)n));
            CUDA_TRY(cudaEventCreateWithFlags(&x->ring_done[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&x->ring_copied[i], cudaEventDisableTiming));
        }
        for (int i = BATCH_RING; i < 8; i++) { CUDA_TRY(cudaEventCreateWithFlags(&x->ring_done[i], cudaEventDisableTiming)); CUDA_TRY(cudaEventCreateWithFlags(&x->ring_copied[i], cudaEventDisableTiming)); }
        x->ring_ready = true;
    }
    if (stats_out) {
        try { x->batch_ev.reserve((size_t)nq * 3); } catch (...) { return ctx_fail(ctx, PPR_E_NOMEM, "ppr_fora_batch: host allocation failed"); }
        while (x->batch_ev.size() < (size_t)nq * 3) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); x->batch_ev.push_back(e); }
        if (x->snaps_cap < (size_t)nq) {
            if (x->snaps) cudaFreeHost(x->snaps);
            x->snaps = nullptr; x->snaps_cap = 0;
            CUDA_TRY(cudaHostAlloc((void**)&x->snaps, (size_t)nq * sizeof(PushCtl), cudaHostAllocDefault));
            x->snaps_cap = (size_t)nq;
        }
    }
    std::vector<double> omegas;
    try { omegas.assign((size_t)std::max(nq, 1), 0.0); } catch (...) { return ctx_fail(ctx, PPR_E_NOMEM, "ppr_fora_batch: host allocation failed"); }
    cudaError_t ce = cudaSuccess;
    int32_t rc = PPR_OK;
#define Q_TRY(expr) do { if (ce == cudaSuccess) { ce = (expr); if (ce != cudaSuccess) rc = ctx_fail_cuda(ctx, ce, #expr, __FILE__, __LINE__); } } while (0)
    // automatic: two queries in flight only while both working sets fit in the L2 together (measured with each query on half of
    // the SMs: +17.5 % at R-MAT scale 20 (358 -> 421 queries/s), -1 % at scale 22)
    const bool small_graph = 4.0 * (double)g->d.m + 8.0 * (double)g->d.n <= 96.0 * 1024 * 1024;
    // (two cooperative kernels are only resident together when an SM holds two of their CTAs)
    const bool two_per_sm = g->ws[0].coop_blocks >= 2 * ctx->sm_count;
    const bool dual = nq > 1 && (ctx->batch_streams == 2 || (ctx->batch_streams == 0 && small_graph));
    // with one cooperative CTA per SM two queries share the GPU as two halves of the SMs
    const int dual_blocks = two_per_sm ? ctx->sm_count : std::max(1, ctx->sm_count / 2);
    bool ring_used[8] = {false, false, false, false, false, false, false, false};
    for (int q = 0; q < nq && rc == PPR_OK; q++) {
        Workspace& w = g->ws[q & 1];
        const int slot = q % BATCH_RING;
        cudaEvent_t* e = stats_out ? &x->batch_ev[(size_t)q * 3] : nullptr;
        // Two compute streams: query q+1 (other workspace, other stream) runs concurrently with query q, so the
        // nearly idle tail iterations of one push overlap the dense phases / walks of the other.  Each cooperative
        // kernel then takes one CTA per SM so that both are resident together.
        cudaStream_t st = dual ? (q & 1 ? ctx->stream2 : ctx->stream) : ctx->stream;
        rc = fora_launch(g, w, srcs[q], alpha, eps, seed + (u64)q, flags, conf, 0, 1, e ? e[0] : nullptr, e ? e[1] : nullptr, e ? e[2] : nullptr,
                         &omegas[q], st, dual ? dual_blocks : 0);
        if (rc) break;
        if (!resident) {
            // the doubles leave the workspace for a ring buffer at once; the download of query q (copy stream) overlaps the
            // kernels of the next BATCH_RING - 1 queries
            if (ring_used[slot]) Q_TRY(cudaStreamWaitEvent(st, x->ring_copied[slot], 0));   // ring[slot] has been downloaded
            k_export_f64<<<ctx->sm_count * 8, 256, 0, st>>>(w.reserve, x->ring[slot], n, 0.0);
            Q_TRY(cudaGetLastError());
            Q_TRY(cudaEventRecord(x->ring_done[slot], st));
            Q_TRY(cudaStreamWaitEvent(ctx->copy, x->ring_done[slot], 0));
            Q_TRY(cudaMemcpyAsync(ppr_out + (size_t)q * n, x->ring[slot], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy));
            Q_TRY(cudaEventRecord(x->ring_copied[slot], ctx->copy));
            ring_used[slot] = true;
        }
        if (stats_out) Q_TRY(cudaMemcpyAsync(&x->snaps[q], w.ctl, sizeof(PushCtl), cudaMemcpyDeviceToHost, st));
    }
#undef Q_TRY
    cudaError_t s1 = cudaStreamSynchronize(ctx->stream), s2 = cudaStreamSynchronize(ctx->copy);
    { cudaError_t s3 = cudaStreamSynchronize(ctx->stream2); if (s1 == cudaSuccess) s1 = s3; }
    if (rc == PPR_OK && (s1 != cudaSuccess || s2 != cudaSuccess)) rc = ctx_fail_cuda(ctx, s1 != cudaSuccess ? s1 : s2, "batch sync", __FILE__, __LINE__);
    if (rc == PPR_OK) {
        if (stats_out) for (int q = 0; q < nq; q++) {
            ppr_stats* st = &stats_out[q];
            stats_from_ctl(st, &x->snaps[q], omegas[q]);
            cudaEvent_t* e = &x->batch_ev[(size_t)q * 3];
            cudaEventElapsedTime(&st->ms_push, e[0], e[1]);
            cudaEventElapsedTime(&st->ms_walk, e[1], e[2]);
            cudaEventElapsedTime(&st->ms_total, e[0], e[2]);
            st->launches = resident ? 2 : 3;
        }
        if (nq > 0) { g->last_ws = (nq - 1) & 1; g->last_kind = 1; x->last_divisor = 0.0; x->last_is_pm = false; }
    }
    return rc;
}

// ============================================================================================= Monte-Carlo
extern "C" int32_t ppr_monte_carlo_shard(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed,
                                         const ppr_conf* conf, int32_t shard, int32_t nshards, ppr_stats* stats) {
    int32_t rc = check_query(g, src, alpha, "ppr_monte_carlo");
    if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    if (!(eps > 0)) return ctx_fail(ctx, PPR_E_INVALID, "ppr_monte_carlo: eps must be > 0");
    if (nshards < 1 || shard < 0 || shard >= nshards) return ctx_fail(ctx, PPR_E_INVALID, "ppr_monte_carlo: bad shard");
    cudaSetDevice(ctx->device);
    Workspace& w = g->ws[0];
    g->trace_active = false;
    // J/Algo_Conf.java:29-35, J/Monte_Carlo.java:145,149
    double delta = 1.0 / (double)g->d.n, pfail = 1.0 / (double)g->d.n;
    if (conf && conf->delta > 0) delta = conf->delta;
    if (conf && conf->pfail > 0) pfail = conf->pfail;
    double omega = 3 * std::log(2 / pfail) / eps / eps / delta;
    if (!(omega < 1e18)) return ctx_fail(ctx, PPR_E_RANGE, "ppr_monte_carlo: omega too large");
    u64 W = omega >= 1.0 ? (u64)std::floor(omega) : 0;
    CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    k_mc_setup<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(w.reserve, g->d.n, w.ws_v, w.ws_begin, w.ws_inc, w.ws_rem, w.ctl,
                                                          g->d.out_deg, src, W, (u64)shard, (u64)nshards);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
    CUDA_TRY(launch_walk(ctx, g, w, w.reserve, seed, 0, alpha, src, 1, 0, (u64)shard, (u64)nshards));
    k_mc_dangling<<<1, 32, 0, ctx->stream>>>(w.reserve, w.ctl, src, W, (u64)shard, (u64)nshards);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ctx->ev[2], ctx->stream));
    rc = read_ctl(g, w); if (rc) return rc;
    g->last_ws = 0; g->last_kind = 1; extra_of(g)->last_divisor = omega; extra_of(g)->last_is_pm = false;
    if (stats) {
        stats_from_ctl(stats, w.ctl_host, omega);
        stats->walks_total = (i64)W;
        cudaEventElapsedTime(&stats->ms_walk, ctx->ev[1], ctx->ev[2]);
        cudaEventElapsedTime(&stats->ms_total, ctx->ev[0], ctx->ev[2]);
        stats->launches = 3;
    }
    return PPR_OK;
}

extern "C" int32_t ppr_monte_carlo(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed, const ppr_conf* conf,
                                   double* ppr_out, ppr_stats* stats) {
    if (g && !ppr_out) return ctx_fail(g->ctx, PPR_E_INVALID, "ppr_monte_carlo: ppr_out is NULL");
    int32_t rc = ppr_monte_carlo_shard(g, src, alpha, eps, seed, conf, 0, 1, stats);
    if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    rc = export_vec(g, g->ws[0], g->ws[0].reserve, extra_of(g)->last_divisor, ppr_out, ctx->stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (stats) stats->launches += 1;
    return PPR_OK;
}

// ============================================================================================= result access
extern "C" int32_t ppr_result_devptr(ppr_graph* g, uint64_t* devptr_out, int64_t* n_out) {
    if (!g || !devptr_out) return PPR_E_INVALID;
    if (g->last_kind == 0) return ctx_fail(g->ctx, PPR_E_STATE, "ppr_result_devptr: no query has run");
    Workspace& w = g->ws[g->last_ws];
    *devptr_out = (uint64_t)(uintptr_t)(g->last_kind == 2 ? w.est : w.reserve);
    if (n_out) *n_out = g->d.n;
    return PPR_OK;
}

extern "C" int32_t ppr_result_export(ppr_graph* g, double divisor, double* out) {
    if (!g || !out) return PPR_E_INVALID;
    ppr_ctx* ctx = g->ctx;
    if (g->last_kind == 0) return ctx_fail(ctx, PPR_E_STATE, "ppr_result_export: no query has run");
    cudaSetDevice(ctx->device);
    Workspace& w = g->ws[g->last_ws];
    int32_t rc = export_vec(g, w, g->last_kind == 2 ? w.est : w.reserve, divisor, out, ctx->stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PPR_OK;
}

extern "C" int32_t ppr_last_raw(ppr_graph* g, uint64_t* reserve_fx, uint64_t* residue_fx) {
    if (!g) return PPR_E_INVALID;
    ppr_ctx* ctx = g->ctx;
    if (g->last_kind == 0) return ctx_fail(ctx, PPR_E_STATE, "ppr_last_raw: no query has run");
    cudaSetDevice(ctx->device);
    Workspace& w = g->ws[g->last_ws];
    size_t bytes = (size_t)g->d.n * sizeof(u64);
    if (reserve_fx) CUDA_TRY(cudaMemcpyAsync(reserve_fx, g->last_kind == 2 ? w.est : w.reserve, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (residue_fx) CUDA_TRY(cudaMemcpyAsync(residue_fx, w.r, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PPR_OK;
}

extern "C" int32_t ppr_debug_timeline(ppr_graph* g, uint64_t* out, int64_t cap_entries, int64_t* count_out) {
    if (!g || !count_out) return PPR_E_INVALID;
    ppr_ctx* ctx = g->ctx;
    cudaSetDevice(ctx->device);
    Workspace& w = g->ws[g->last_ws];
    int32_t rc = read_ctl(g, w); if (rc) return rc;
    i64 cnt = std::min<i64>(w.ctl_host->dbg_count, DBG_ITERS);
    *count_out = cnt;
    if (out && cnt) {
        i64 c = std::min<i64>(cnt, cap_entries);
        CUDA_TRY(cudaMemcpy(out, w.dbg, (size_t)c * DBG_FIELDS * sizeof(u64), cudaMemcpyDeviceToHost));
    }
    return PPR_OK;
}

// ============================================================================================= top-k
static int32_t run_select(ppr_graph* g, Workspace& w, const u64* vals, i64 n, i64 k) {
    ppr_ctx* ctx = g->ctx;
    SelArgs S;
    S.vals = vals; S.n = n; S.k = k; S.hist = w.hist; S.sel_ids = w.sel_ids; S.sel_vals = w.sel_vals; S.sel_cap = w.sel_cap; S.ctl = w.ctl;
    void* args[] = {&S};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_topk_select, dim3(w.sel_blocks), dim3(SEL_THREADS), args, 0, ctx->stream));
    return PPR_OK;
}

// survivors -> (value desc, id asc), the order of getTopKNodeIds (J/Fora_Topk.java:82-99; ties are unordered there)
static int32_t fetch_survivors(ppr_graph* g, Workspace& w, std::vector<i32>& ids, std::vector<u64>& vals) {
    // the <= k + ties survivors are put in order (value descending, id ascending) on the device (k_rank_sort) and downloaded once
    ppr_ctx* ctx = g->ctx;
    i64 cnt = w.ctl_host->sel_count;
    if (cnt > w.sel_cap) return ctx_fail(ctx, PPR_E_CAPACITY, "top-k: too many tied survivors for the selection buffer");
    ids.resize((size_t)cnt); vals.resize((size_t)cnt);
    if (cnt) {
        GraphExtra* x = extra_of(g);
        if (x->sort_cap < cnt) {
            cudaFree(x->sort_ids); cudaFree(x->sort_vals); x->sort_ids = nullptr; x->sort_vals = nullptr; x->sort_cap = 0; x->bwd_order = nullptr; x->bwd_order_cap = 0; x->edge_inv = nullptr;
            const i64 cap = std::max<i64>(cnt, 4096);
            CUDA_TRY(dalloc(&x->sort_ids, (size_t)cap)); CUDA_TRY(dalloc(&x->sort_vals, (size_t)cap));
            x->sort_cap = cap;
        }
        k_rank_sort<<<(unsigned)std::min<i64>((cnt + 127) / 128, 4096), 128, 0, ctx->stream>>>(w.sel_ids, w.sel_vals, cnt, x->sort_ids, x->sort_vals);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(ids.data(), x->sort_ids, (size_t)cnt * sizeof(i32), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(vals.data(), x->sort_vals, (size_t)cnt * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return PPR_OK;
}

static int32_t fora_topk_impl(ppr_graph* g, int32_t src, double alpha, double eps, int32_t k, uint64_t seed, int32_t flags,
                             const ppr_conf* conf, int32_t shard, int32_t nshards, ppr_reduce_fn reduce, void* user,
                             int32_t* ids_out, double* vals_out, int32_t cap, int32_t* count_out, ppr_stats* stats) {
    int32_t rc = check_fora(g, src, alpha, eps, shard, nshards, "ppr_fora_topk");
    if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    if (k < 1 || !count_out || cap < 0 || (cap > 0 && (!ids_out || !vals_out))) return ctx_fail(ctx, PPR_E_INVALID, "ppr_fora_topk: bad k / buffers");
    if (nshards > 1 && !reduce) return ctx_fail(ctx, PPR_E_INVALID, "ppr_fora_topk_shard: nshards > 1 needs a reduce hook");
    cudaSetDevice(ctx->device);
    Workspace& w = g->ws[0];
    g->trace_active = false;
    const i64 n = g->d.n, m = g->d.m;
    // J/Algo_Conf.java:71-81 (note the integer division n / k) and J/Fora_Topk.java:109-114
    double min_delta = 1.0 / (double)n;
    double delta_local = 1.0 / (double)k;
    double pfail = 1.0 / (double)n / (double)n / std::log((double)(n / (i64)k));
    if (conf && conf->delta > 0) delta_local = conf->delta;
    if (conf && conf->pfail > 0) pfail = conf->pfail;
    if (conf && conf->min_delta > 0) min_delta = conf->min_delta;
    double epsilon = eps * 0.5;
    double L = std::log(2.0 / pfail);
    float ms_push = 0, ms_walk = 0, ms_sel = 0;
    int launches = 0;
    int round = 0;
    double omega_local = 0;
    bool dangling = false;
    CUDA_TRY(cudaEventRecord(ctx->ev[4], ctx->stream));
    while (delta_local >= min_delta) {
        double rmax_local = epsilon * std::sqrt(delta_local / 3.0 / (double)m / L);
        omega_local = (epsilon + 2.0) * L / epsilon / epsilon / delta_local;
        rmax_local *= std::sqrt((double)m * rmax_local) * 3.0;  // J/Fora_Topk.java:133
        PushArgs A;
        fill_push_args(A, g, w);
        A.src = src; A.alpha_q64 = h_alpha_q64(alpha); A.alpha = alpha; A.omega = omega_local; A.rmax0 = rmax_local;
        A.max_rounds = 1; A.mode = PMODE_TOPK; A.init = round == 0;
        A.prep_adds_push_part = shard == 0;   // only shard 0's estimate starts from the push reserve
        CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
        CUDA_TRY(launch_push(ctx, w, A));
        CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
        CUDA_TRY(launch_walk(ctx, g, w, w.est, seed, (u64)round, alpha, src, (flags & PPR_RESTART_AT_SOURCE) ? 1 : 0, 0, (u64)shard, (u64)nshards));
        CUDA_TRY(cudaEventRecord(ctx->ev[2], ctx->stream));
        if (nshards > 1) {  // sum the round's estimates over the shards (caller's collective) before selecting
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (reduce(user, (uint64_t)(uintptr_t)w.est, n) != 0) return ctx_fail(ctx, PPR_E_CALLBACK, "ppr_fora_topk_shard: the reduce hook failed");
            CUDA_TRY(cudaEventRecord(ctx->ev[2], ctx->stream));   // the collective is not counted as select time
        }
        rc = run_select(g, w, w.est, n, k); if (rc) return rc;
        CUDA_TRY(cudaEventRecord(ctx->ev[3], ctx->stream));
        rc = read_ctl(g, w); if (rc) return rc;
        launches += 3;
        float t;
        cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]); ms_push += t;
        cudaEventElapsedTime(&t, ctx->ev[1], ctx->ev[2]); ms_walk += t;
        cudaEventElapsedTime(&t, ctx->ev[2], ctx->ev[3]); ms_sel += t;
        if (w.ctl_host->src_dangling) { dangling = true; break; }  // J/Fora_Topk.java:128-132
        round++;
        double kth_d = (double)w.ctl_host->sel_kth * FX_INV;
        if (kth_d >= (1 + epsilon) * delta_local || delta_local <= min_delta) break;
        delta_local = std::max(min_delta, delta_local / 4.0);
    }
    CUDA_TRY(cudaEventRecord(ctx->ev[5], ctx->stream));
    g->last_ws = 0; g->last_kind = 2; extra_of(g)->last_divisor = 0.0; extra_of(g)->last_is_pm = false;
    std::vector<i32> ids; std::vector<u64> vals;
    rc = fetch_survivors(g, w, ids, vals); if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (stats) {
        stats_from_ctl(stats, w.ctl_host, omega_local);
        stats->rounds = dangling ? 0 : round;
        stats->ms_push = ms_push; stats->ms_walk = ms_walk; stats->ms_select = ms_sel;
        cudaEventElapsedTime(&stats->ms_total, ctx->ev[4], ctx->ev[5]);
        stats->launches = launches;
    }
    *count_out = (int32_t)ids.size();
    if ((i64)ids.size() > cap) return ctx_fail(ctx, PPR_E_CAPACITY, "ppr_fora_topk: output buffers too small");
    for (size_t i = 0; i < ids.size(); i++) { ids_out[i] = ids[i]; vals_out[i] = (double)vals[i] * FX_INV; }
    return PPR_OK;
}

extern "C" int32_t ppr_fora_topk(ppr_graph* g, int32_t src, double alpha, double eps, int32_t k, uint64_t seed, int32_t flags,
                                 const ppr_conf* conf, int32_t* ids_out, double* vals_out, int32_t cap, int32_t* count_out, ppr_stats* stats) {
    return fora_topk_impl(g, src, alpha, eps, k, seed, flags, conf, 0, 1, nullptr, nullptr, ids_out, vals_out, cap, count_out, stats);
}

extern "C" int32_t ppr_fora_topk_shard(ppr_graph* g, int32_t src, double alpha, double eps, int32_t k, uint64_t seed, int32_t flags,
                                       const ppr_conf* conf, int32_t shard, int32_t nshards, ppr_reduce_fn reduce, void* user,
                                       int32_t* ids_out, double* vals_out, int32_t cap, int32_t* count_out, ppr_stats* stats) {
    return fora_topk_impl(g, src, alpha, eps, k, seed, flags, conf, shard, nshards, reduce, user, ids_out, vals_out, cap, count_out, stats);
}

extern "C" int32_t ppr_topk(ppr_ctx* ctx, const double* vec, int64_t n, int32_t k, int32_t* ids_out, double* vals_out,
                            int32_t cap, int32_t* count_out, double* kth_out) {
    if (!ctx) return PPR_E_INVALID;
    if (!vec || n < 1 || !count_out || cap < 0 || (cap > 0 && (!ids_out || !vals_out))) return ctx_fail(ctx, PPR_E_INVALID, "ppr_topk: bad arguments");
    cudaSetDevice(ctx->device);
    // grow-only device scratch kept by the context (getTopKNodeIds is called once per query by Forward_Push / Monte_Carlo /
    // Power_Method): vector, bit image, survivors and ordered survivors, histograms, control block
    const size_t N = (size_t)n, al = 256;
    auto up = [&](size_t x) { return (x + al - 1) / al * al; };
    const size_t o_vec = 0, o_bits = o_vec + up(N * 8), o_sv = o_bits + up(N * 8), o_ov = o_sv + up(N * 8), o_si = o_ov + up(N * 8),
                 o_oi = o_si + up(N * 4), o_hist = o_oi + up(N * 4), o_ctl = o_hist + up(9 * 256 * 4), total = o_ctl + up(sizeof(PushCtl));
    if (ctx->topk_bytes < total) {
        cudaFree(ctx->topk_buf); ctx->topk_buf = nullptr; ctx->topk_bytes = 0;
        CUDA_TRY(cudaMalloc(&ctx->topk_buf, total));
        ctx->topk_bytes = total;
    }
    char* base = static_cast<char*>(ctx->topk_buf);
    double* d_vec = reinterpret_cast<double*>(base + o_vec); u64* d_bits = reinterpret_cast<u64*>(base + o_bits);
    u64* sel_vals = reinterpret_cast<u64*>(base + o_sv); u64* ord_vals = reinterpret_cast<u64*>(base + o_ov);
    i32* sel_ids = reinterpret_cast<i32*>(base + o_si); i32* ord_ids = reinterpret_cast<i32*>(base + o_oi);
    u32* hist = reinterpret_cast<u32*>(base + o_hist); PushCtl* ctl = reinterpret_cast<PushCtl*>(base + o_ctl);
    PushCtl hctl;
    CUDA_TRY(cudaMemcpyAsync(d_vec, vec, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_import_f64<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_vec, d_bits, n);
    CUDA_TRY(cudaGetLastError());
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_topk_select, SEL_THREADS, 0));
    SelArgs S;
    S.vals = d_bits; S.n = n; S.k = k; S.hist = hist; S.sel_ids = sel_ids; S.sel_vals = sel_vals; S.sel_cap = n; S.ctl = ctl;
    void* args[] = {&S};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_topk_select, dim3(ctx->sm_count * std::max(1, std::min(occ, 4))), dim3(SEL_THREADS), args, 0, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(&hctl, ctl, sizeof hctl, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const i64 cnt = hctl.sel_count;
    *count_out = (int32_t)cnt;
    if (kth_out) { double kd; u64 kb = hctl.sel_kth; memcpy(&kd, &kb, 8); *kth_out = kb ? kd : 0.0; }
    if (cnt > cap) return ctx_fail(ctx, PPR_E_CAPACITY, "ppr_topk: output buffers too small");
    if (cnt) {   // value descending, id ascending, ordered on the device; the bit images are the doubles themselves
        k_rank_sort<<<(unsigned)std::min<i64>((cnt + 127) / 128, 4096), 128, 0, ctx->stream>>>(sel_ids, sel_vals, cnt, ord_ids, ord_vals);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(ids_out, ord_ids, (size_t)cnt * sizeof(i32), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaMemcpyAsync(vals_out, ord_vals, (size_t)cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return PPR_OK;
}

// ============================================================================================= multi-GPU (NCCL behind the ABI)
#define NCCL_TRY(api, expr) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) return ctx_fail(ctx, PPR_E_COMM, std::string("NCCL: ") + (api)->GetErrorString(r__) + " in " #expr); } while (0)

extern "C" int32_t ppr_comm_unique_id(void* id_out) {
    if (!id_out) return PPR_E_INVALID;
    std::string err;
    NcclApi* api = nccl_api(&err);
    if (!api) return ctx_fail(nullptr, PPR_E_COMM, err);
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) return ctx_fail(nullptr, PPR_E_COMM, std::string("ncclGetUniqueId: ") + api->GetErrorString(r));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id_out, &id, sizeof id);
    return PPR_OK;
}

extern "C" int32_t ppr_comm_init(ppr_ctx* ctx, const void* id, int32_t rank, int32_t world) {
    if (!ctx) return PPR_E_INVALID;
    if (!id || world < 1 || rank < 0 || rank >= world) return ctx_fail(ctx, PPR_E_INVALID, "ppr_comm_init: bad arguments");
    if (ctx->comm) return ctx_fail(ctx, PPR_E_STATE, "ppr_comm_init: the context already has a communicator");
    std::string err;
    NcclApi* api = nccl_api(&err);
    if (!api) return ctx_fail(ctx, PPR_E_COMM, err);
    cudaSetDevice(ctx->device);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    ncclComm_t comm = nullptr;
    NCCL_TRY(api, api->CommInitRank(&comm, world, uid, rank));
    ctx->comm = comm; ctx->rank = rank; ctx->world = world;
    return PPR_OK;
}

extern "C" int32_t ppr_comm_destroy(ppr_ctx* ctx) {
    if (!ctx) return PPR_E_INVALID;
    if (ctx->comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        NcclApi* api = nccl_api(nullptr);
        if (api) api->CommDestroy((ncclComm_t)ctx->comm);
        ctx->comm = nullptr; ctx->rank = 0; ctx->world = 1;
    }
    return PPR_OK;
}

extern "C" int32_t ppr_comm_info(ppr_ctx* ctx, int32_t* rank, int32_t* world, int32_t* nccl_version) {
    if (!ctx) return PPR_E_INVALID;
    if (rank) *rank = ctx->rank;
    if (world) *world = ctx->world;
    if (nccl_version) { *nccl_version = 0; NcclApi* api = nccl_api(nullptr); if (api) { int v = 0; api->GetVersion(&v); *nccl_version = v; } }
    return PPR_OK;
}

// in-place integer sum of n raw uint64 values over the ranks, on the library's stream (no-op for a single rank)
static int32_t comm_allreduce_u64(ppr_ctx* ctx, u64* buf, i64 n) {
    if (ctx->world <= 1) return PPR_OK;
    NcclApi* api = nccl_api(nullptr);
    if (!api || !ctx->comm) return ctx_fail(ctx, PPR_E_COMM, "sharded call without a communicator (ppr_comm_init)");
    NCCL_TRY(api, api->AllReduce(buf, buf, (size_t)n, ncclUint64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return PPR_OK;
}

extern "C" int32_t ppr_comm_max_f64(ppr_ctx* ctx, double* inout, int32_t count) {
    if (!ctx || !inout || count < 0 || count > 64) return PPR_E_INVALID;
    cudaSetDevice(ctx->device);
    CUDA_TRY(cudaDeviceSynchronize());
    if (ctx->world <= 1) return PPR_OK;
    NcclApi* api = nccl_api(nullptr);
    if (!api || !ctx->comm) return ctx_fail(ctx, PPR_E_COMM, "ppr_comm_max_f64 without a communicator (ppr_comm_init)");
    double* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, 64 * sizeof(double)));
    cudaError_t e = cudaMemcpyAsync(d, inout, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t r = e == cudaSuccess ? api->AllReduce(d, d, (size_t)count, ncclDouble, ncclMax, (ncclComm_t)ctx->comm, ctx->stream) : ncclSuccess;
    if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpyAsync(inout, d, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    if (r != ncclSuccess) return ctx_fail(ctx, PPR_E_COMM, std::string("NCCL: ") + api->GetErrorString(r));
    CUDA_TRY(e);
    return PPR_OK;
}

static int32_t require_comm(ppr_graph* g, const char* who) {
    if (!g) return PPR_E_INVALID;
    if (g->ctx->world > 1 && !g->ctx->comm) return ctx_fail(g->ctx, PPR_E_COMM, std::string(who) + ": no communicator");
    return PPR_OK;
}

extern "C" int32_t ppr_fora_sharded(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed, int32_t flags,
                                    const ppr_conf* conf, double* ppr_out, ppr_stats* stats) {
    int32_t rc = require_comm(g, "ppr_fora_sharded"); if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    rc = ppr_fora_shard(g, src, alpha, eps, seed, flags, conf, ctx->rank, ctx->world, stats); if (rc) return rc;
    Workspace& w = g->ws[0];
    CUDA_TRY(cudaEventRecord(ctx->ev[3], ctx->stream));
    rc = comm_allreduce_u64(ctx, w.reserve, g->d.n); if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (ppr_out) { rc = export_vec(g, w, w.reserve, 0.0, ppr_out, ctx->stream); if (rc) return rc; }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (stats) { cudaEventElapsedTime(&stats->ms_comm, ctx->ev[3], ctx->ev[4]); stats->ms_total += stats->ms_comm; stats->launches += ppr_out ? 2 : 1; }
    return PPR_OK;
}

extern "C" int32_t ppr_monte_carlo_sharded(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed, const ppr_conf* conf,
                                           double* ppr_out, ppr_stats* stats) {
    int32_t rc = require_comm(g, "ppr_monte_carlo_sharded"); if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    rc = ppr_monte_carlo_shard(g, src, alpha, eps, seed, conf, ctx->rank, ctx->world, stats); if (rc) return rc;
    Workspace& w = g->ws[0];
    CUDA_TRY(cudaEventRecord(ctx->ev[3], ctx->stream));
    rc = comm_allreduce_u64(ctx, w.reserve, g->d.n); if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (ppr_out) { rc = export_vec(g, w, w.reserve, extra_of(g)->last_divisor, ppr_out, ctx->stream); if (rc) return rc; }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (stats) { cudaEventElapsedTime(&stats->ms_comm, ctx->ev[3], ctx->ev[4]); stats->ms_total += stats->ms_comm; stats->launches += ppr_out ? 2 : 1; }
    return PPR_OK;
}

static int32_t nccl_reduce_hook(void* user, uint64_t dev_ptr, int64_t n) {
    ppr_ctx* ctx = (ppr_ctx*)user;
    if (comm_allreduce_u64(ctx, (u64*)(uintptr_t)dev_ptr, n) != PPR_OK) return 1;
    return cudaStreamSynchronize(ctx->stream) == cudaSuccess ? 0 : 1;
}

extern "C" int32_t ppr_fora_topk_sharded(ppr_graph* g, int32_t src, double alpha, double eps, int32_t k, uint64_t seed, int32_t flags,
                                         const ppr_conf* conf, int32_t* ids_out, double* vals_out, int32_t cap, int32_t* count_out, ppr_stats* stats) {
    int32_t rc = require_comm(g, "ppr_fora_topk_sharded"); if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    return fora_topk_impl(g, src, alpha, eps, k, seed, flags, conf, ctx->rank, ctx->world, ctx->world > 1 ? nccl_reduce_hook : nullptr, ctx,
                          ids_out, vals_out, cap, count_out, stats);
}

// ============================================================================================= power iteration
extern "C" int32_t ppr_power_iteration(ppr_graph* g, int32_t src, double alpha, int32_t iters, double* out) {
    int32_t rc = check_query(g, src, alpha, "ppr_power_iteration");
    if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    if (!out || iters < 0) return ctx_fail(ctx, PPR_E_INVALID, "ppr_power_iteration: bad arguments");
    cudaSetDevice(ctx->device);
    const i64 n = g->d.n;
    const int nblk = ctx->sm_count * 4;
    double *r = nullptr, *r2 = nullptr, *q = nullptr, *res = nullptr, *blk = nullptr;
    auto done = [&]() { cudaFree(r); cudaFree(r2); cudaFree(q); cudaFree(res); cudaFree(blk); };
    cudaError_t e;
#define P_TRY(expr) do { e = (expr); if (e != cudaSuccess) { done(); return ctx_fail_cuda(ctx, e, #expr, __FILE__, __LINE__); } } while (0)
    P_TRY(dalloc(&r, (size_t)n)); P_TRY(dalloc(&r2, (size_t)n)); P_TRY(dalloc(&q, (size_t)n)); P_TRY(dalloc(&res, (size_t)n)); P_TRY(dalloc(&blk, (size_t)nblk));
    P_TRY(cudaMemsetAsync(r, 0, (size_t)n * 8, ctx->stream)); P_TRY(cudaMemsetAsync(res, 0, (size_t)n * 8, ctx->stream));
    double one = 1.0;
    P_TRY(cudaMemcpyAsync(r + src, &one, 8, cudaMemcpyHostToDevice, ctx->stream));
    for (int it = 0; it < iters; it++) {
        k_pm_scale<<<nblk, 256, 0, ctx->stream>>>(r, q, res, g->d.out_deg, n, alpha, blk);
        k_pm_pull<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(q, r2, g->d.in_vtx, g->d.row_idx, n, src, blk, nblk);
        std::swap(r, r2);
    }
    P_TRY(cudaGetLastError());
    P_TRY(cudaMemcpyAsync(out, res, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    {   // the result also stays on the device for ppr_metrics_keep (on-device error metrics against this ground truth)
        GraphExtra* x = extra_of(g);
        if (!x->pm_result) P_TRY(dalloc(&x->pm_result, (size_t)n));
        P_TRY(cudaMemcpyAsync(x->pm_result, res, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        x->last_is_pm = true;
    }
    P_TRY(cudaStreamSynchronize(ctx->stream));
#undef P_TRY
    done();
    return PPR_OK;
}

// ============================================================================================= error metrics on the device
// Gen_Util.computeError (J/Gen_Util.java:259-326).  ppr_metrics_keep snapshots the result of the last query (as doubles, on
// the device) as ground truth (slot 0: normally the power iteration) or estimate (slot 1); ppr_metrics_compute evaluates
// MAX_ERR, PRECISION and NDCG of slot 1 against slot 0 without moving either vector to the host.
extern "C" int32_t ppr_metrics_keep(ppr_graph* g, int32_t slot) {
    if (!g) return PPR_E_INVALID;
    ppr_ctx* ctx = g->ctx;
    if (slot < 0 || slot > 1) return ctx_fail(ctx, PPR_E_INVALID, "ppr_metrics_keep: slot must be 0 (ground truth) or 1 (estimate)");
    GraphExtra* x = extra_of(g);
    cudaSetDevice(ctx->device);
    const i64 n = g->d.n;
    if (!x->metric_vec[slot]) CUDA_TRY(dalloc(&x->metric_vec[slot], (size_t)n));
    if (x->last_is_pm) {
        CUDA_TRY(cudaMemcpyAsync(x->metric_vec[slot], x->pm_result, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        if (g->last_kind == 0) return ctx_fail(ctx, PPR_E_STATE, "ppr_metrics_keep: no query has run");
        Workspace& w = g->ws[g->last_ws];
        k_export_f64<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(g->last_kind == 2 ? w.est : w.reserve, x->metric_vec[slot], n, x->last_divisor);
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    x->metric_set[slot] = true;
    return PPR_OK;
}

// tie-inclusive top-k of a double vector resident on the device, ordered (value descending, id ascending), left on the device
static int32_t device_topk_f64(ppr_graph* g, Workspace& w, const double* vec, i64 k, i32* ids_out, u64* vals_out, i64 cap, i64* count) {
    ppr_ctx* ctx = g->ctx;
    const i64 n = g->d.n;
    k_import_f64<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(vec, w.q[0], n);   // order-preserving bit image; q[0] is free outside a push
    CUDA_TRY(cudaGetLastError());
    int32_t rc = run_select(g, w, w.q[0], n, k); if (rc) return rc;
    rc = read_ctl(g, w); if (rc) return rc;
    const i64 c = w.ctl_host->sel_count;
    if (c > w.sel_cap || c > cap) return ctx_fail(ctx, PPR_E_CAPACITY, "ppr_metrics_compute: too many tied top-k entries");
    if (c) { k_rank_sort<<<(unsigned)std::min<i64>((c + 127) / 128, 4096), 128, 0, ctx->stream>>>(w.sel_ids, w.sel_vals, c, ids_out, vals_out); CUDA_TRY(cudaGetLastError()); }
    *count = c;
    return PPR_OK;
}

extern "C" int32_t ppr_metrics_compute(ppr_graph* g, int32_t k, double* max_err, double* precision, double* ndcg) {
    if (!g) return PPR_E_INVALID;
    ppr_ctx* ctx = g->ctx;
    GraphExtra* x = extra_of(g);
    if (!x->metric_set[0] || !x->metric_set[1]) return ctx_fail(ctx, PPR_E_STATE, "ppr_metrics_compute: keep a ground truth (slot 0) and an estimate (slot 1) first");
    if (k < 0) return ctx_fail(ctx, PPR_E_INVALID, "ppr_metrics_compute: k < 0");
    cudaSetDevice(ctx->device);
    const i64 n = g->d.n;
    Workspace& w = g->ws[0];
    ull* d_bits = nullptr; double* d_out = nullptr; double* d_gain = nullptr; i32* ids[2] = {nullptr, nullptr}; u64* vals[2] = {nullptr, nullptr};
    const i64 cap = w.sel_cap;
    auto done = [&]() { cudaFree(d_bits); cudaFree(d_out); cudaFree(d_gain); cudaFree(ids[0]); cudaFree(ids[1]); cudaFree(vals[0]); cudaFree(vals[1]); };
#define M_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { done(); return ctx_fail_cuda(ctx, e__, #expr, __FILE__, __LINE__); } } while (0)
    M_TRY(dalloc(&d_bits, 1)); M_TRY(dalloc(&d_out, 2));
    M_TRY(cudaMemsetAsync(d_bits, 0, 8, ctx->stream)); M_TRY(cudaMemsetAsync(d_out, 0, 16, ctx->stream));
    k_max_err<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(x->metric_vec[0], x->metric_vec[1], n, d_bits);
    M_TRY(cudaGetLastError());
    i64 cnt[2] = {0, 0};
    if (k > 0 && (precision || ndcg)) {
        for (int s = 0; s < 2; s++) {
            M_TRY(dalloc(&ids[s], (size_t)cap)); M_TRY(dalloc(&vals[s], (size_t)cap));
            int32_t rc = device_topk_f64(g, w, x->metric_vec[s], k, ids[s], vals[s], cap, &cnt[s]);
            if (rc) { done(); return rc; }
        }
        M_TRY(dalloc(&d_gain, (size_t)std::max<i64>(cnt[1], 1)));
        k_topk_metrics<<<1, 256, 0, ctx->stream>>>(ids[0], vals[0], cnt[0], ids[1], cnt[1], d_gain, d_out);
        M_TRY(cudaGetLastError());
    }
    ull h_bits = 0; double h_out[2] = {0, 0};
    M_TRY(cudaMemcpyAsync(&h_bits, d_bits, 8, cudaMemcpyDeviceToHost, ctx->stream));
    M_TRY(cudaMemcpyAsync(h_out, d_out, 16, cudaMemcpyDeviceToHost, ctx->stream));
    M_TRY(cudaStreamSynchronize(ctx->stream));
#undef M_TRY
    done();
    if (max_err) memcpy(max_err, &h_bits, 8);
    if (precision) *precision = h_out[0];
    if (ndcg) *ndcg = h_out[1];
    return PPR_OK;
}

// host-side stage timer for tuning (PPR_DEBUG_TIMING=1 prints to stderr)
#include <chrono>
struct StageTimer {
    bool on; std::chrono::steady_clock::time_point t;
    StageTimer() : on(getenv("PPR_DEBUG_TIMING") != nullptr), t(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[pprb200] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

// ============================================================================================= backward search
static int32_t bwd_run_batch(ppr_graph* g, BwdWorkspace& b, i32 t0, i32 T, double alpha, double threshold, PushCtl* hctl) {
    ppr_ctx* ctx = g->ctx;
    Workspace& w = g->ws[0];
    const i64 n = g->d.n;
    BwdArgs A;
    memset(&A, 0, sizeof A);
    A.n = n; A.in_vtx = g->d.in_vtx; A.row_idx = g->d.row_idx; A.invdeg = g->d.invdeg;
    A.r = b.r; A.reserve = b.reserve; A.front[0] = b.front[0]; A.front[1] = b.front[1]; A.front_cap = b.front_cap;
    A.s_pack = b.s_pack; A.s_y = b.s_y; A.s_t = b.s_t; A.l_pack = b.l_pack; A.l_y = b.l_y; A.l_t = b.l_t;
    A.chunks = b.chunks; A.chunk_cap = b.chunk_cap; A.wcnt = reinterpret_cast<u32*>(w.blk_c); A.dbg = w.dbg; A.ctl = w.ctl; A.t0 = t0; A.T = T;
    A.alpha_q64 = h_alpha_q64(alpha);
    // J/Backward_Search.java:89 (r > rmax) and J/Base_Whole_Graph.java:83 (pi >= threshold), in fixed point
    double fl = threshold > 0 ? std::floor(threshold * FX_SCALE) : 0.0;
    double cl = threshold > 0 ? std::ceil(threshold * FX_SCALE) : 0.0;
    A.rmax_fx = fl >= 18446744073709551615.0 ? ~0ull : (u64)fl;
    A.rmax_fx = std::max<u64>(A.rmax_fx, ~0ull / A.alpha_q64 + 2);   // floor: a push must move >= 1 unit into the reserve
    A.thr_fx = cl >= 18446744073709551615.0 ? ~0ull : (u64)cl;
    A.out_v = b.out_v; A.out_t = b.out_t; A.out_val = b.out_val; A.out_cap = b.out_cap;
    CUDA_TRY(cudaMemsetAsync(b.r, 0, (size_t)T * n * 8, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(b.reserve, 0, (size_t)T * n * 8, ctx->stream));
    void* args[] = {&A};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_bwd_batch, dim3(w.coop_blocks), dim3(PUSH_THREADS), args, 0, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(w.ctl_host, w.ctl, sizeof(PushCtl), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *hctl = *w.ctl_host;
    return PPR_OK;
}

static int32_t bwd_ensure(ppr_graph* g, i32 T) {
    GraphExtra* x = extra_of(g);
    if (x->bws_ready && x->bws.T >= T) return PPR_OK;
    if (x->bws_ready) { bws_free(x->bws); x->bws_ready = false; }
    int32_t rc = bws_alloc(g->ctx, x->bws, T, g->d.n, x->bwd_chunks_per_target);
    if (rc) { bws_free(x->bws); return rc; }
    x->bws_ready = true;
    return PPR_OK;
}

extern "C" int32_t ppr_backward_search(ppr_graph* g, int32_t target, double alpha, double rmax, double* reserve_out,
                                       ppr_stats* stats) {
    int32_t rc = check_query(g, target, alpha, "ppr_backward_search");
    if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    if (!reserve_out) return ctx_fail(ctx, PPR_E_INVALID, "ppr_backward_search: reserve_out is NULL");
    cudaSetDevice(ctx->device);
    g->trace_active = false;
    rc = bwd_ensure(g, 1); if (rc) return rc;
    BwdWorkspace& b = extra_of(g)->bws;
    PushCtl h;
    CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    rc = bwd_run_batch(g, b, target, 1, alpha, rmax, &h); if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (h.err || h.err_a) return ctx_fail(ctx, PPR_E_CAPACITY, "ppr_backward_search: frontier buffers overflowed");
    // keep the raw vectors reachable through ppr_last_raw: copy the single-target state into workspace 0
    Workspace& w = g->ws[0];
    CUDA_TRY(cudaMemcpyAsync(w.reserve, b.reserve, (size_t)g->d.n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(w.r, b.r, (size_t)g->d.n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    g->last_ws = 0; g->last_kind = 1; extra_of(g)->last_divisor = 0.0; extra_of(g)->last_is_pm = false;
    rc = export_vec(g, w, w.reserve, 0.0, reserve_out, ctx->stream); if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->pushes = (i64)h.pushes; stats->edges = (i64)h.edges; stats->iters = (i64)h.iters;
        cudaEventElapsedTime(&stats->ms_push, ctx->ev[0], ctx->ev[1]);
        stats->ms_total = stats->ms_push; stats->launches = 2;
    }
    return PPR_OK;
}

static void apply_topk_rows(ppr_pairs* P, int32_t k);
static int32_t pairs_to_host(ppr_ctx* ctx, ppr_pairs* P);

// Launch configuration of k_bwd_cta: CTA size and CTAs per SM (PPR_BWD_BD / PPR_BWD_CPS / PPR_BWD_SLOTS override it for tuning runs).
template <int BD, int TEAM>
static cudaError_t launch_bwd_cta(const BwdCtaArgs& A, int slots, cudaStream_t st) {
    if (TEAM == 1) {
        k_bwd_cta<BD, 1><<<slots, BD, 0, st>>>(A);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)slots * TEAM); cfg.blockDim = dim3(BD); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = TEAM; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_bwd_cta<BD, TEAM>, A);
}
static cudaError_t launch_bwd_cta_bd(int bd, int team, const BwdCtaArgs& A, int slots, cudaStream_t st) {
    if (team == 8) return launch_bwd_cta<1024, 8>(A, slots, st);
    if (team == 4) return launch_bwd_cta<1024, 4>(A, slots, st);
    if (team == 2) return launch_bwd_cta<1024, 2>(A, slots, st);
    switch (bd) {
        case 64: return launch_bwd_cta<64, 1>(A, slots, st);
        case 128: return launch_bwd_cta<128, 1>(A, slots, st);
        case 256: return launch_bwd_cta<256, 1>(A, slots, st);
        case 512: return launch_bwd_cta<512, 1>(A, slots, st);
        default: return launch_bwd_cta<1024, 1>(A, slots, st);
    }
}
// searches a team launch can keep in flight (clusters resident at once)
static int bwd_team_slots(ppr_ctx* ctx, int team) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)ctx->sm_count / team * team); cfg.blockDim = dim3(1024);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = team; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    cudaError_t e = team == 8 ? cudaOccupancyMaxActiveClusters(&nc, k_bwd_cta<1024, 8>, &cfg)
                  : team == 4 ? cudaOccupancyMaxActiveClusters(&nc, k_bwd_cta<1024, 4>, &cfg)
                              : cudaOccupancyMaxActiveClusters(&nc, k_bwd_cta<1024, 2>, &cfg);
    if (e != cudaSuccess || nc < 1) { cudaGetLastError(); nc = std::max(1, ctx->sm_count / team / 2); }
    return nc;
}
static void bwd_cta_config(ppr_ctx* ctx, double edges_per_target, int* bd, int* cps) {
    // Measured on B200 (profiles/r02_bwd_cta_shapes.txt): one 1024-thread CTA per SM wins at every threshold of the reference
    // grid -- the fewer searches are in flight, the more of their sparse state stays in the L2 (148 slots: 60 / 101 / 1111 ms
    // at 1e-3 / 5e-4 / 5e-5; 1184 slots of 128 threads: 193 / 303 / 2555 ms).
    (void)edges_per_target;
    *bd = 1024; *cps = 1;
    const char* e;
    if ((e = getenv("PPR_BWD_BD"))) { int v = atoi(e); if (v == 64 || v == 128 || v == 256 || v == 512 || v == 1024) *bd = v; }
    if ((e = getenv("PPR_BWD_CPS"))) { int v = atoi(e); if (v >= 1 && v <= 32) *cps = v; }
    *cps = std::min(*cps, 2048 / *bd);
    int occ = 0;   // CTAs beyond what is resident would only queue (and cost 48 n bytes of slot state each)
    cudaError_t e2 = cudaSuccess;
    switch (*bd) {
        case 64: e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bwd_cta<64, 1>, 64, 0); break;
        case 128: e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bwd_cta<128, 1>, 128, 0); break;
        case 256: e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bwd_cta<256, 1>, 256, 0); break;
        case 512: e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bwd_cta<512, 1>, 512, 0); break;
        default: e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_bwd_cta<1024, 1>, 1024, 0); break;
    }
    if (e2 == cudaSuccess && occ >= 1) *cps = std::min(*cps, occ);
}
static int32_t bcw_ensure(ppr_graph* g, int slots, int bd) {
    GraphExtra* x = extra_of(g);
    if (x->bcw_ready && x->bcw.slots >= slots) { x->bcw.bd = bd; return PPR_OK; }
    u64* key[2] = {nullptr, nullptr}; u64* val[2] = {nullptr, nullptr}; i64 cap = 0; u32* hist = nullptr; size_t hist_len = 0;
    if (x->bcw_ready) {   // keep the pair buffers: they may already hold the pairs of earlier launches of this call
        for (int i = 0; i < 2; i++) { key[i] = x->bcw.key[i]; val[i] = x->bcw.val[i]; x->bcw.key[i] = nullptr; x->bcw.val[i] = nullptr; }
        cap = x->bcw.cap; hist = x->bcw.hist; hist_len = x->bcw.hist_len; x->bcw.hist = nullptr;
        bcw_free(x->bcw); x->bcw_ready = false;
    }
    int32_t rc = bcw_alloc(g->ctx, x->bcw, slots, bd, g->d.n);
    for (int i = 0; i < 2; i++) { x->bcw.key[i] = key[i]; x->bcw.val[i] = val[i]; }
    x->bcw.cap = cap; x->bcw.hist = hist; x->bcw.hist_len = hist_len;
    if (rc) { bcw_free(x->bcw); return rc; }
    x->bcw_ready = true;
    return PPR_OK;
}

// (source, target) order of the pair buffer: LSD radix sort over the 2 * key_bits key bits, then targets + row pointers
static int32_t sort_pairs(ppr_ctx* ctx, BwdCtaWorkspace& b, i64 count, int key_bits, i64 n, i32* d_tgt, i64* d_ptr, int* result_buf) {
    int cur = 0;
    if (count > 0) {
        const u32 nb = (u32)((count + RS_TILE - 1) / RS_TILE);
        const size_t need = (size_t)256 * nb;
        if (b.hist_len < need) { cudaFree(b.hist); b.hist = nullptr; b.hist_len = 0; CUDA_TRY(dalloc(&b.hist, need)); b.hist_len = need; }
        for (int shift = 0; shift < 2 * key_bits; shift += 8) {
            k_rs_hist<<<nb, RS_THREADS, 0, ctx->stream>>>(b.key[cur], count, shift, b.hist, nb);
            k_rs_scan<<<1, 1024, 0, ctx->stream>>>(b.hist, (i64)need);
            k_rs_scatter<<<nb, RS_THREADS, 0, ctx->stream>>>(b.key[cur], b.val[cur], count, shift, b.hist, nb, b.key[cur ^ 1], b.val[cur ^ 1]);
            CUDA_TRY(cudaGetLastError());
            cur ^= 1;
        }
        k_rs_rows<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(b.key[cur], count, key_bits, n, d_tgt, d_ptr);
        CUDA_TRY(cudaGetLastError());
    } else {
        CUDA_TRY(cudaMemsetAsync(d_ptr, 0, ((size_t)n + 1) * 8, ctx->stream));
    }
    *result_buf = cur;
    return PPR_OK;
}

extern "C" int32_t ppr_backward_all_pairs(ppr_graph* g, double alpha, double threshold, int32_t k, int32_t t_lo, int32_t t_hi,
                                          ppr_pairs** out, ppr_stats* stats) {
    if (!g) return PPR_E_INVALID;
    ppr_ctx* ctx = g->ctx;
    if (!out) return ctx_fail(ctx, PPR_E_INVALID, "ppr_backward_all_pairs: out is NULL");
    *out = nullptr;
    if (!(alpha > 0 && alpha < 1)) return ctx_fail(ctx, PPR_E_INVALID, "ppr_backward_all_pairs: alpha must be in (0, 1)");
    const i64 n = g->d.n;
    if (t_lo < 0 || t_hi > n || t_lo > t_hi) return ctx_fail(ctx, PPR_E_INVALID, "ppr_backward_all_pairs: bad target range");
    cudaSetDevice(ctx->device);
    g->trace_active = false;
    StageTimer tm;
    GraphExtra* x = extra_of(g);
    int key_bits = 1;
    while (key_bits < 31 && ((i64)1 << key_bits) < n) key_bits++;

    BwdCtaArgs A;
    memset(&A, 0, sizeof A);
    A.n = n; A.in_vtx = g->d.in_vtx; A.row_idx = g->d.row_idx; A.invdeg = g->d.invdeg;
    if (!x->edge_inv && g->d.m > 0 && !getenv("PPR_BWD_NO_EDGE_INV")) {   // 8 bytes per edge, built once per graph: removes a dependent gather from every relaxed edge
        if (cudaMalloc((void**)&x->edge_inv, (size_t)g->d.m * 8) == cudaSuccess) {
            k_build_edge_inv<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(g->d.row_idx, g->d.invdeg, g->d.m, x->edge_inv);
            CUDA_TRY(cudaGetLastError());
        } else { cudaGetLastError(); x->edge_inv = nullptr; }   // no room: keep the gather
    }
    A.edge_inv = getenv("PPR_BWD_NO_EDGE_INV") ? nullptr : x->edge_inv;
    A.alpha_q64 = h_alpha_q64(alpha);
    // J/Backward_Search.java:89 (r > rmax) and J/Base_Whole_Graph.java:83 (pi >= threshold), in fixed point
    {
        const double fl = threshold > 0 ? std::floor(threshold * FX_SCALE) : 0.0, cl = threshold > 0 ? std::ceil(threshold * FX_SCALE) : 0.0;
        A.rmax_fx = fl >= 18446744073709551615.0 ? ~0ull : (u64)fl;
        A.rmax_fx = std::max<u64>(A.rmax_fx, ~0ull / A.alpha_q64 + 2);   // floor: a push must move >= 1 unit into the reserve
        A.thr_fx = cl >= 18446744073709551615.0 ? ~0ull : (u64)cl;
    }
    A.key_bits = key_bits;
    // dense iterations of a search (frontier owning >= n / 8 in-edges): plain reductions + a sweep of the row instead of returning atomics
    { const char* e = getenv("PPR_BWD_SCAN"); A.scan_edges = e ? (u64)atoll(e) : (u64)std::max<i64>(n / 8, 4096); }

    // per-slot state is 48 n bytes; stay within half of the free memory
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    if (x->bcw_ready) free_b += (size_t)x->bcw.slots * (size_t)n * 48;
    const i64 max_slots = (i64)std::max<size_t>(1, (free_b / 2) / ((size_t)48 * (size_t)n));
    i64 pushes = 0, edges = 0, iters = 0;
    int launches = 0;
    i64 cursor = 0;                     // pairs accepted so far (device buffer 0)
    double edges_per_target = -1.0;
    CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    const bool timing = getenv("PPR_BWD_TIMING") != nullptr;
    ull h_tim[6] = {0, 0, 0, 0, 0, 0};
    // Launch plan.  One CTA owns a search from start to end, and search sizes are heavy-tailed (at threshold 1e-3 the mean search
    // relaxes 12 K edges, the largest ~10 M: 35 ms on one CTA, half of the whole run), so:
    //   (A) the `heavy` targets of highest in-degree run first, as teams: thread-block clusters of 8 CTAs per search;
    //   (B) a probe launch, one CTA per search, over an evenly spaced sample of the remaining targets (in descending in-degree order):
    //       pairs and relaxed edges per target;
    //   (C) the rest, heaviest first (longest-processing-time-first); as teams of 2 CTAs when the probe's searches relax >= 64 K edges
    //       on average (fewer searches in flight keep their residue rows in the L2: 836 against 1 090 ms at threshold 5e-5).
    const i64 R = t_hi - t_lo;
    int team_heavy = 4, team_rest = 2;
    i64 heavy = -1;
    { const char* e;
      if ((e = getenv("PPR_BWD_TEAM"))) { int v = atoi(e); team_heavy = (v == 2 || v == 4 || v == 8) ? v : 1; }
      if ((e = getenv("PPR_BWD_TEAM_REST"))) { int v = atoi(e); team_rest = (v == 2 || v == 4 || v == 8) ? v : 1; }
      if ((e = getenv("PPR_BWD_HEAVY"))) heavy = atoll(e); }
    const bool force_team = getenv("PPR_BWD_FORCE_TEAM") != nullptr && team_heavy > 1;   // tests: every launch as teams of team_heavy
    const bool ordered = R > 0 && !getenv("PPR_BWD_NO_ORDER");
    struct Launch { i64 count; int team; bool probe; };
    std::vector<Launch> plan;
    std::vector<i32> order;
    try {
        if (ordered) {
            std::vector<i32> sorted((size_t)R);
            for (i64 i = 0; i < R; i++) sorted[(size_t)i] = (i32)(t_lo + i);
            const std::vector<u32>& deg = x->h_indeg;
            std::stable_sort(sorted.begin(), sorted.end(), [&](i32 a, i32 c) { return deg[(size_t)a] > deg[(size_t)c]; });
            i64 H = 0, probe = 0;
            if (ctx->bwd_batch <= 0 && !force_team) {
                if (team_heavy > 1) H = std::min<i64>(R, heavy >= 0 ? heavy : (i64)8 * bwd_team_slots(ctx, team_heavy));
                probe = std::min<i64>(R - H, std::max<i64>(1024, (i64)ctx->sm_count * 8));
            }
            // layout of the order: [probe sample of the non-heavy targets][the H heaviest][the other targets, heaviest first]
            order.reserve((size_t)R);
            const i64 rest = R - H;
            std::vector<char> taken((size_t)std::max<i64>(rest, 1), 0);
            if (probe > 0 && probe < rest) {
                const i64 stride = rest / probe;
                for (i64 i = 0; i < probe; i++) { order.push_back(sorted[(size_t)(H + i * stride + stride / 2)]); taken[(size_t)(i * stride + stride / 2)] = 1; }
            } else if (ctx->bwd_batch <= 0 && !force_team) {
                probe = 0;   // small ranges: no probe launch, the heavy block and the rest are everything
            }
            for (i64 i = 0; i < H; i++) order.push_back(sorted[(size_t)i]);
            for (i64 i = 0; i < rest; i++) if (!taken[(size_t)i]) order.push_back(sorted[(size_t)(H + i)]);
            if (ctx->bwd_batch > 0 || force_team) {
                const i64 step = ctx->bwd_batch > 0 ? ctx->bwd_batch : R;
                for (i64 i = 0; i < R; i += step) plan.push_back({std::min<i64>(step, R - i), force_team ? team_heavy : 1, false});
            } else {
                if (probe > 0) plan.push_back({probe, 1, true});
                if (H > 0) plan.push_back({H, -1 /* heavy block: team_heavy, or merged into the rest when every search is large */, false});
                if (rest - probe > 0) plan.push_back({rest - probe, 0 /* decided after the probe */, false});
            }
        } else if (R > 0) {
            const i64 step = ctx->bwd_batch > 0 ? ctx->bwd_batch : R;
            for (i64 i = 0; i < R; i += step) plan.push_back({std::min<i64>(step, R - i), 1, false});
        }
    } catch (...) { return ctx_fail(ctx, PPR_E_NOMEM, "ppr_backward_all_pairs: host allocation failed"); }
    if (!order.empty()) {
        if (x->bwd_order_cap < R) {
            cudaFree(x->bwd_order); x->bwd_order = nullptr; x->bwd_order_cap = 0;
            CUDA_TRY(dalloc(&x->bwd_order, (size_t)R));
            x->bwd_order_cap = R;
        }
        CUDA_TRY(cudaMemcpyAsync(x->bwd_order, order.data(), (size_t)R * sizeof(i32), cudaMemcpyHostToDevice, ctx->stream));
    }
    double pairs_per_target = 64.0;   // refined by the probe
    i64 t0 = t_lo;
    for (size_t li = 0; li < plan.size(); li++) {
        int bd, cps;
        bwd_cta_config(ctx, edges_per_target, &bd, &cps);
        i64 Tb = plan[li].count;
        int lteam = plan[li].team;
        const bool all_large = edges_per_target >= 65536.0;   // (from the probe; -1 before any launch)
        if (lteam == -1) {   // the heavy block
            if (all_large && li + 1 < plan.size() && plan[li + 1].team == 0) { Tb += plan[li + 1].count; lteam = team_rest; li++; }   // one launch with the rest, heaviest first
            else lteam = team_heavy;
        } else if (lteam == 0) lteam = all_large ? team_rest : 1;
        i64 want_slots = lteam > 1 ? (i64)bwd_team_slots(ctx, lteam) : (i64)ctx->sm_count * cps;
        { const char* e = getenv("PPR_BWD_SLOTS"); if (e && atoll(e) > 0 && lteam == 1) want_slots = atoll(e); }
        const int slots = (int)std::max<i64>(1, std::min<i64>(std::min<i64>(want_slots, max_slots), Tb));
        int32_t rc = bcw_ensure(g, std::max(slots, x->bcw_ready ? x->bcw.slots : 0), bd); if (rc) return rc;
        BwdCtaWorkspace& b = x->bcw;
        // room for this launch's pairs: estimate, retry with the exact need on overflow (the searches are deterministic)
        const i64 want = cursor + std::max<i64>((i64)1 << 22, (i64)(pairs_per_target * (double)Tb * 1.25) + 65536);
        rc = bcw_reserve_pairs(ctx, b, want, cursor); if (rc) return rc;
        const i64 edges_before = edges, cursor_before = cursor;
        for (;;) {
            ull h_cnt[16] = {(ull)cursor, (ull)pushes, (ull)edges, (ull)iters, h_tim[0], h_tim[1], h_tim[2], h_tim[3], h_tim[4], h_tim[5], 0, 0, 0, 0, 0, 0};
            CUDA_TRY(cudaMemcpyAsync(b.counters, h_cnt, sizeof h_cnt, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(cudaMemsetAsync(b.ticket, 0, 4, ctx->stream));
            A.r = b.r; A.reserve = b.reserve; A.front = b.front; A.pushed = b.pushed; A.touched = b.touched; A.st_pk = b.st_pk; A.st_y = b.st_y;
            A.ticket = b.ticket; A.counters = b.counters; A.t_lo = (i32)t0; A.t_hi = (i32)(t0 + Tb);
            A.order = order.empty() ? nullptr : x->bwd_order + (t0 - t_lo); A.n_order = (i32)Tb;   // launches walk through the order in slices
            A.out_key = b.key[0]; A.out_val = b.val[0]; A.out_cap = b.cap; A.timing = timing ? 1 : 0;
            { const char* e = getenv("PPR_BWD_DBG"); A.dbg = e ? atoi(e) : 0; }
            CUDA_TRY(launch_bwd_cta_bd(bd, lteam, A, slots, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(h_cnt, b.counters, sizeof h_cnt, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            launches++;
            if ((i64)h_cnt[0] > b.cap) {   // pair buffer too small: run the launch again with room
                rc = bcw_reserve_pairs(ctx, b, (i64)h_cnt[0] + 65536, cursor); if (rc) return rc;
                continue;
            }
            cursor = (i64)h_cnt[0]; pushes = (i64)h_cnt[1]; edges = (i64)h_cnt[2]; iters = (i64)h_cnt[3];
            for (int i = 0; i < 6; i++) h_tim[i] = h_cnt[4 + i];
            break;
        }
        tm.mark(lteam > 1 ? "reverse-push launch (teams)" : "reverse-push launch");
        t0 += Tb;
        if (plan[li].probe || (edges_per_target < 0 && lteam == 1)) {
            edges_per_target = (double)(edges - edges_before) / (double)Tb;
            pairs_per_target = std::max(1.0, (double)(cursor - cursor_before) / (double)Tb);
        }
    }
    if (timing && h_tim[4])
        fprintf(stderr, "[pprb200] k_bwd_cta per target (thread 0 clock, us): ticket+seed %.2f  phase A %.2f  phase B %.2f  output+clean %.2f  (%llu targets with in-edges, %.2f iterations each; longest target %.1f us)\n",
                h_tim[0] / 1e3 / h_tim[4], h_tim[1] / 1e3 / h_tim[4], h_tim[2] / 1e3 / h_tim[4], h_tim[3] / 1e3 / h_tim[4], h_tim[4], (double)iters / (double)h_tim[4], h_tim[5] / 1e3);
    CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
    if (cursor > 0xffffffffll) return ctx_fail(ctx, PPR_E_CAPACITY, "ppr_backward_all_pairs: more than 2^32 pairs in one call; split the target range");
    // inverted index grouped by source v, targets ascending (J/Base_Whole_Graph.java:83-87): device radix sort + row pointers.
    // The index stays on the device; ppr_pairs_export copies it into the caller's arrays.
    ppr_pairs* P = new (std::nothrow) ppr_pairs();
    if (!P) return ctx_fail(ctx, PPR_E_NOMEM, "ppr_backward_all_pairs: host allocation failed");
    P->n = n; P->count = cursor; P->device = ctx->device; P->d_tgt = nullptr; P->d_val = nullptr; P->on_host = false;
    i64* d_ptr = nullptr;
    auto fail = [&](int32_t code) { cudaFree(d_ptr); ppr_pairs_destroy(P); return code; };
    try { P->src_ptr.assign((size_t)n + 1, 0); }
    catch (...) { return fail(ctx_fail(ctx, PPR_E_NOMEM, "ppr_backward_all_pairs: host allocation failed")); }
#define P_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return fail(ctx_fail_cuda(ctx, e__, #expr, __FILE__, __LINE__)); } while (0)
    if (!x->bcw_ready) { int32_t rc = bcw_ensure(g, 1, 1024); if (rc) return fail(rc); }   // empty target range: nothing was launched
    BwdCtaWorkspace& b = x->bcw;
    P_TRY(dalloc(&P->d_tgt, (size_t)cursor)); P_TRY(dalloc(&P->d_val, (size_t)cursor)); P_TRY(dalloc(&d_ptr, (size_t)n + 1));
    int buf = 0;
    { int32_t rc = sort_pairs(ctx, b, cursor, key_bits, n, P->d_tgt, d_ptr, &buf); if (rc) return fail(rc); }
    if (cursor) P_TRY(cudaMemcpyAsync(P->d_val, b.val[buf], (size_t)cursor * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    P_TRY(cudaEventRecord(ctx->ev[2], ctx->stream));
    P_TRY(cudaMemcpyAsync(P->src_ptr.data(), d_ptr, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    P_TRY(cudaStreamSynchronize(ctx->stream));
#undef P_TRY
    cudaFree(d_ptr); d_ptr = nullptr;
    tm.mark("sort + row pointers");
    if (k >= 0) {   // J/Base_Whole_Graph.java:133-148 on the host (rows are short after the cut; not the throughput path)
        int32_t rc = pairs_to_host(ctx, P); if (rc) { ppr_pairs_destroy(P); return rc; }
        try { apply_topk_rows(P, k); } catch (...) { ppr_pairs_destroy(P); return ctx_fail(ctx, PPR_E_NOMEM, "ppr_backward_all_pairs: host allocation failed"); }
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->pushes = pushes; stats->edges = edges; stats->iters = iters;
        cudaEventElapsedTime(&stats->ms_push, ctx->ev[0], ctx->ev[1]);
        float ms_sort = 0; cudaEventElapsedTime(&ms_sort, ctx->ev[1], ctx->ev[2]);
        stats->ms_select = ms_sort;
        stats->ms_total = stats->ms_push + ms_sort; stats->launches = launches; stats->rounds = launches;
    }
    *out = P;
    return PPR_OK;
}


// J/Base_Whole_Graph.java:133-148: per source keep the entries >= the k-th largest, sorted descending (stable: ties by ascending target)
static void apply_topk_rows(ppr_pairs* P, int32_t k) {
    const i64 n = P->n;
    std::vector<i64> nptr((size_t)n + 1, 0);
    std::vector<i32> ntgt; std::vector<u64> nval;
    std::vector<i64> ord; std::vector<u64> tmp;
    for (i64 v = 0; v < n; v++) {
        i64 lo = P->src_ptr[(size_t)v], hi = P->src_ptr[(size_t)v + 1], len = hi - lo;
        u64 kth = 0; bool have = false;
        if (k >= 1 && k <= len) {
            tmp.assign(P->val.begin() + lo, P->val.begin() + hi);
            std::nth_element(tmp.begin(), tmp.begin() + (k - 1), tmp.end(), [](u64 a, u64 c) { return a > c; });
            kth = tmp[(size_t)k - 1]; have = true;
        }
        ord.clear();
        for (i64 i = lo; i < hi; i++) if (!have || P->val[(size_t)i] >= kth) ord.push_back(i);
        std::stable_sort(ord.begin(), ord.end(), [&](i64 a, i64 c) { return P->val[(size_t)a] > P->val[(size_t)c]; });
        for (i64 i : ord) { ntgt.push_back(P->tgt[(size_t)i]); nval.push_back(P->val[(size_t)i]); }
        nptr[(size_t)v + 1] = (i64)ntgt.size();
    }
    P->src_ptr.swap(nptr); P->tgt.swap(ntgt); P->val.swap(nval);
    P->count = (i64)P->tgt.size();
}

// move a device-resident index into the host vectors of the handle
static int32_t pairs_to_host(ppr_ctx* ctx, ppr_pairs* P) {
    if (P->on_host) return PPR_OK;
    try { P->tgt.resize((size_t)P->count); P->val.resize((size_t)P->count); }
    catch (...) { return ctx_fail(ctx, PPR_E_NOMEM, "ppr_pairs: host allocation failed"); }
    cudaSetDevice(P->device);
    if (P->count) {
        CUDA_TRY(cudaMemcpy(P->tgt.data(), P->d_tgt, (size_t)P->count * 4, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(P->val.data(), P->d_val, (size_t)P->count * 8, cudaMemcpyDeviceToHost));
    }
    cudaFree(P->d_tgt); cudaFree(P->d_val); P->d_tgt = nullptr; P->d_val = nullptr;
    P->on_host = true;
    return PPR_OK;
}

// ---------------------------------------------------------------------------------------------- all pairs over the ranks
// Targets are block-partitioned (same split as the walk shards); no communication until the index is complete.  Gather:
// ncclAllGather of the per-source counts, then every rank's (target, value) arrays travel once (one broadcast per
// rank inside one NCCL group) into a rank-major device buffer, and a merge kernel writes row v of the result as the
// concatenation, in rank order = ascending targets, of the ranks' rows v.
__global__ void k_merge_rows(const i64* __restrict__ cnt_all /*[world][n]*/, const i64* __restrict__ rank_base /*[world]*/, int world, i64 n,
                             const i64* __restrict__ mptr /*[n+1]*/, const i32* __restrict__ tgt_in, const u64* __restrict__ val_in,
                             const i64* __restrict__ row_in /*[world][n]: start of row v inside rank r's slice*/,
                             i32* __restrict__ tgt_out, u64* __restrict__ val_out) {
    const i64 gw = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((i64)gridDim.x * blockDim.x) >> 5;
    const u32 lane = threadIdx.x & 31;
    for (i64 v = gw; v < n; v += nw) {
        i64 dst = mptr[v];
        for (int r = 0; r < world; r++) {
            const i64 c = cnt_all[(i64)r * n + v], src = rank_base[r] + row_in[(i64)r * n + v];
            for (i64 i = lane; i < c; i += 32) { tgt_out[dst + i] = tgt_in[src + i]; val_out[dst + i] = val_in[src + i]; }
            dst += c;
        }
    }
}
__global__ void k_rows_excl(const i64* __restrict__ cnt_all, int world, i64 n, i64* __restrict__ row_in) {
    // exclusive prefix of every rank's row counts (one CTA per rank; n is a few 10^5 for the all-pair configs)
    __shared__ i64 carry, part[1024];
    const int r = blockIdx.x;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (i64 base = 0; base < n; base += blockDim.x) {
        const i64 v = base + threadIdx.x;
        const i64 c = v < n ? cnt_all[(i64)r * n + v] : 0;
        part[threadIdx.x] = c;
        __syncthreads();
        for (unsigned o = 1; o < blockDim.x; o <<= 1) { i64 t = threadIdx.x >= o ? part[threadIdx.x - o] : 0; __syncthreads(); part[threadIdx.x] += t; __syncthreads(); }
        if (v < n) row_in[(i64)r * n + v] = carry + part[threadIdx.x] - c;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += part[threadIdx.x];
        __syncthreads();
    }
}

extern "C" int32_t ppr_backward_all_pairs_sharded(ppr_graph* g, double alpha, double threshold, int32_t k, ppr_pairs** out, ppr_stats* stats) {
    int32_t rc = require_comm(g, "ppr_backward_all_pairs_sharded"); if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    if (!out) return ctx_fail(ctx, PPR_E_INVALID, "ppr_backward_all_pairs_sharded: out is NULL");
    const i64 n = g->d.n;
    const int world = ctx->world, rank = ctx->rank;
    const i32 lo = (i32)((n * rank) / world), hi = (i32)((n * (rank + 1)) / world);
    ppr_pairs* L = nullptr;
    rc = ppr_backward_all_pairs(g, alpha, threshold, world > 1 ? -1 : k, lo, hi, &L, stats); if (rc) return rc;
    if (world == 1) { *out = L; return PPR_OK; }
    NcclApi* api = nccl_api(nullptr);
    cudaSetDevice(ctx->device);
    i64 *d_cnt = nullptr, *d_all = nullptr, *d_rows = nullptr, *d_base = nullptr, *d_mptr = nullptr;
    i32 *d_tin = nullptr, *d_tout = nullptr; u64 *d_vin = nullptr, *d_vout = nullptr;
    ppr_pairs* P = nullptr;
    auto done = [&]() { cudaFree(d_cnt); cudaFree(d_all); cudaFree(d_rows); cudaFree(d_base); cudaFree(d_mptr); cudaFree(d_tin); cudaFree(d_vin); };
    auto fail = [&](int32_t code) { done(); cudaFree(d_tout); cudaFree(d_vout); ppr_pairs_destroy(L); if (P) { P->d_tgt = nullptr; P->d_val = nullptr; ppr_pairs_destroy(P); } return code; };
#define S_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return fail(ctx_fail_cuda(ctx, e__, #expr, __FILE__, __LINE__)); } while (0)
#define SN_TRY(expr) do { ncclResult_t r__ = (expr); if (r__ != ncclSuccess) return fail(ctx_fail(ctx, PPR_E_COMM, std::string("NCCL: ") + api->GetErrorString(r__) + " in " #expr)); } while (0)
    cudaEvent_t e0 = ctx->ev[3], e1 = ctx->ev[4];
    S_TRY(cudaEventRecord(e0, ctx->stream));
    std::vector<i64> cnt((size_t)n), all((size_t)n * world), base((size_t)world + 1, 0), mptr((size_t)n + 1, 0);
    for (i64 v = 0; v < n; v++) cnt[(size_t)v] = L->src_ptr[(size_t)v + 1] - L->src_ptr[(size_t)v];
    S_TRY(dalloc(&d_cnt, (size_t)n)); S_TRY(dalloc(&d_all, (size_t)n * world)); S_TRY(dalloc(&d_rows, (size_t)n * world));
    S_TRY(dalloc(&d_base, (size_t)world)); S_TRY(dalloc(&d_mptr, (size_t)n + 1));
    S_TRY(cudaMemcpyAsync(d_cnt, cnt.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    SN_TRY(api->AllGather(d_cnt, d_all, (size_t)n, ncclInt64, (ncclComm_t)ctx->comm, ctx->stream));
    S_TRY(cudaMemcpyAsync(all.data(), d_all, (size_t)n * world * 8, cudaMemcpyDeviceToHost, ctx->stream));
    S_TRY(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < world; r++) { i64 t = 0; for (i64 v = 0; v < n; v++) t += all[(size_t)r * n + v]; base[(size_t)r + 1] = base[(size_t)r] + t; }
    for (i64 v = 0; v < n; v++) { i64 t = 0; for (int r = 0; r < world; r++) t += all[(size_t)r * n + v]; mptr[(size_t)v + 1] = mptr[(size_t)v] + t; }
    const i64 total = base[(size_t)world];
    S_TRY(dalloc(&d_tin, (size_t)total)); S_TRY(dalloc(&d_vin, (size_t)total)); S_TRY(dalloc(&d_tout, (size_t)total)); S_TRY(dalloc(&d_vout, (size_t)total));
    const i64 mine = base[(size_t)rank + 1] - base[(size_t)rank];
    if (mine) {   // the local index is device resident (k < 0): no host round trip
        if (L->on_host) {
            S_TRY(cudaMemcpyAsync(d_tin + base[(size_t)rank], L->tgt.data(), (size_t)mine * 4, cudaMemcpyHostToDevice, ctx->stream));
            S_TRY(cudaMemcpyAsync(d_vin + base[(size_t)rank], L->val.data(), (size_t)mine * 8, cudaMemcpyHostToDevice, ctx->stream));
        } else {
            S_TRY(cudaMemcpyAsync(d_tin + base[(size_t)rank], L->d_tgt, (size_t)mine * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            S_TRY(cudaMemcpyAsync(d_vin + base[(size_t)rank], L->d_val, (size_t)mine * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        }
    }
    SN_TRY(api->GroupStart());
    for (int r = 0; r < world; r++) {
        const i64 c = base[(size_t)r + 1] - base[(size_t)r];
        if (!c) continue;
        SN_TRY(api->Broadcast(d_tin + base[(size_t)r], d_tin + base[(size_t)r], (size_t)c, ncclInt32, r, (ncclComm_t)ctx->comm, ctx->stream));
        SN_TRY(api->Broadcast(d_vin + base[(size_t)r], d_vin + base[(size_t)r], (size_t)c, ncclUint64, r, (ncclComm_t)ctx->comm, ctx->stream));
    }
    SN_TRY(api->GroupEnd());
    S_TRY(cudaMemcpyAsync(d_base, base.data(), (size_t)world * 8, cudaMemcpyHostToDevice, ctx->stream));
    S_TRY(cudaMemcpyAsync(d_mptr, mptr.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_rows_excl<<<world, 1024, 0, ctx->stream>>>(d_all, world, n, d_rows);
    k_merge_rows<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d_all, d_base, world, n, d_mptr, d_tin, d_vin, d_rows, d_tout, d_vout);
    S_TRY(cudaGetLastError());
    P = new (std::nothrow) ppr_pairs();
    if (!P) return fail(ctx_fail(ctx, PPR_E_NOMEM, "ppr_backward_all_pairs_sharded: host allocation failed"));
    P->n = n; P->count = total; P->device = ctx->device; P->d_tgt = nullptr; P->d_val = nullptr; P->on_host = false;
    try { P->src_ptr = mptr; }
    catch (...) { return fail(ctx_fail(ctx, PPR_E_NOMEM, "ppr_backward_all_pairs_sharded: host allocation failed")); }
    S_TRY(cudaEventRecord(e1, ctx->stream));
    S_TRY(cudaStreamSynchronize(ctx->stream));
#undef S_TRY
#undef SN_TRY
    P->d_tgt = d_tout; P->d_val = d_vout; d_tout = nullptr; d_vout = nullptr;   // the merged index stays on the device (ppr_pairs_export)
    if (stats) { cudaEventElapsedTime(&stats->ms_comm, e0, e1); stats->ms_total += stats->ms_comm; }
    done();
    ppr_pairs_destroy(L);
    if (k >= 0) {   // re-select per source over the union of the ranks' rows (J/Base_Whole_Graph.java:133-148)
        rc = pairs_to_host(ctx, P); if (rc) { ppr_pairs_destroy(P); return rc; }
        try { apply_topk_rows(P, k); } catch (...) { ppr_pairs_destroy(P); return ctx_fail(ctx, PPR_E_NOMEM, "ppr_backward_all_pairs_sharded: host allocation failed"); }
    }
    *out = P;
    return PPR_OK;
}

extern "C" int64_t ppr_pairs_count(const ppr_pairs* p) { return p ? (int64_t)p->count : -1; }
extern "C" int32_t ppr_pairs_export(const ppr_pairs* p, int64_t* src_ptr, int32_t* tgt, double* val, uint64_t* val_fx) {
    if (!p || !src_ptr) return PPR_E_INVALID;
    memcpy(src_ptr, p->src_ptr.data(), p->src_ptr.size() * sizeof(i64));
    if (p->count == 0) return PPR_OK;
    if (p->on_host) {
        if (tgt) memcpy(tgt, p->tgt.data(), p->tgt.size() * sizeof(i32));
        if (val) for (size_t i = 0; i < p->val.size(); i++) val[i] = (double)p->val[i] * FX_INV;
        if (val_fx) memcpy(val_fx, p->val.data(), p->val.size() * sizeof(u64));
        return PPR_OK;
    }
    // device-resident index: straight into the caller's arrays (doubles are converted on the device: value * 2^-62 is exact).
    // (A hand-rolled double-buffered copy through pinned staging was measured slower than cudaMemcpy into pageable memory here:
    // 1.3 s against 0.5 s for the 75 M-pair index, the pinned allocations alone cost ~0.1 s per call.)
    if (cudaSetDevice(p->device) != cudaSuccess) return PPR_E_CUDA;
    const size_t c = (size_t)p->count;
    if (tgt && cudaMemcpy(tgt, p->d_tgt, c * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return PPR_E_CUDA;
    if (val_fx && cudaMemcpy(val_fx, p->d_val, c * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return PPR_E_CUDA;
    if (val) {
        double* d = nullptr;
        if (cudaMalloc((void**)&d, c * 8) != cudaSuccess) return PPR_E_NOMEM;
        k_export_f64<<<(unsigned)std::min<size_t>((c + 255) / 256, 65535), 256>>>(p->d_val, d, (i64)c, 0.0);
        cudaError_t e = cudaMemcpy(val, d, c * 8, cudaMemcpyDeviceToHost);
        cudaFree(d);
        if (e != cudaSuccess) return PPR_E_CUDA;
    }
    return PPR_OK;
}
extern "C" int32_t ppr_pairs_destroy(ppr_pairs* p) {
    if (!p) return PPR_OK;
    if (p->d_tgt || p->d_val) { cudaSetDevice(p->device); cudaFree(p->d_tgt); cudaFree(p->d_val); }
    delete p;
    return PPR_OK;
}

// ============================================================================================= stepped push (trace)
extern "C" int32_t ppr_push_trace_begin(ppr_graph* g, int32_t src, double alpha, double rmax) {
    int32_t rc = check_query(g, src, alpha, "ppr_push_trace_begin");
    if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    if (!(rmax > 0)) return ctx_fail(ctx, PPR_E_INVALID, "ppr_push_trace_begin: rmax must be > 0");
    cudaSetDevice(ctx->device);
    Workspace& w = g->ws[0];
    PushArgs A;
    fill_push_args(A, g, w);
    A.src = src; A.alpha_q64 = h_alpha_q64(alpha); A.alpha = alpha; A.rmax0 = rmax; A.mode = PMODE_PUSH; A.init = 1;
    int blocks = ctx->sm_count * 4;
    k_step_init<<<blocks, PUSH_THREADS, 0, ctx->stream>>>(A);
    k_step_seed<<<1, 32, 0, ctx->stream>>>(A);
    k_step_scan<<<blocks, PUSH_THREADS, 0, ctx->stream>>>(A, 0);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    g->trace_active = true; g->trace_parity = 0; g->trace_src = src; g->trace_alpha_q64 = A.alpha_q64;
    g->last_ws = 0; g->last_kind = 1; extra_of(g)->last_divisor = 0.0; extra_of(g)->last_is_pm = false;
    return PPR_OK;
}

extern "C" int32_t ppr_push_trace_step(ppr_graph* g, int32_t* frontier_out, int64_t cap, int64_t* count_out) {
    if (!g || !count_out) return PPR_E_INVALID;
    ppr_ctx* ctx = g->ctx;
    if (!g->trace_active) return ctx_fail(ctx, PPR_E_STATE, "ppr_push_trace_step: call ppr_push_trace_begin first");
    cudaSetDevice(ctx->device);
    Workspace& w = g->ws[0];
    int par = g->trace_parity;
    int32_t rc = read_ctl(g, w); if (rc) return rc;
    i64 cnt = w.ctl_host->cnt[par];
    *count_out = cnt;
    if (cnt == 0) return PPR_OK;
    if (frontier_out) {
        if (cnt > cap) return ctx_fail(ctx, PPR_E_CAPACITY, "ppr_push_trace_step: frontier_out too small");
        CUDA_TRY(cudaMemcpyAsync(frontier_out, w.front[par], (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PushArgs A;
    fill_push_args(A, g, w);
    A.src = g->trace_src; A.alpha_q64 = g->trace_alpha_q64; A.mode = PMODE_PUSH;
    int blocks = ctx->sm_count * 4;
    k_step_a<<<blocks, PUSH_THREADS, 0, ctx->stream>>>(A, par);
    k_step_b<<<blocks, PUSH_THREADS, 0, ctx->stream>>>(A, par);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    g->trace_parity = par ^ 1;
    return PPR_OK;
}

extern "C" int32_t ppr_walk_terminals(ppr_graph* g, int32_t start, int32_t restart, uint64_t j0, int64_t count, uint64_t seed,
                                      uint64_t tag, double alpha, int32_t no_zero_hop, int32_t* terminals_out) {
    int32_t rc = check_query(g, start, alpha, "ppr_walk_terminals");
    if (rc) return rc;
    ppr_ctx* ctx = g->ctx;
    if (restart < 0 || restart >= g->d.n || count < 0 || !terminals_out) return ctx_fail(ctx, PPR_E_INVALID, "ppr_walk_terminals: bad arguments");
    if (count == 0) return PPR_OK;
    cudaSetDevice(ctx->device);
    i32* d_out = nullptr;
    CUDA_TRY(dalloc(&d_out, (size_t)count));
    u64 key = h_walk_key(seed, tag);
    k_walk_terminals<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(g->d.out_vtx, g->d.col_idx, (u32)start, (u32)restart, j0,
                                                                              count, (u32)key, (u32)(key >> 32),
                                                                              h_stop_threshold(alpha), no_zero_hop, d_out);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(terminals_out, d_out, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_out);
    CUDA_TRY(e);
    return PPR_OK;
}
