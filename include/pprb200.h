/* pprb200.h -- C ABI of libpprb200.so, the B200 (sm_100a) Personalized-PageRank engine.
 *
 * This is the drop-in boundary for the hot path of joezie/Personalized-PageRank-Algorithms-on-Neo4j.
 * The reference has no FFI layer: its operator API is three Java interfaces
 * (Whole_Graph_Util_Interface, Topk_Util_Interface, Preprocessing_Interface) implemented by one class
 * per algorithm.  Each entry point below names the reference method it replaces
 * (J/ = src/main/java/joezie/fora_neo4j/).  The Java classes keep their names and call these through
 * JNI or Panama; the Python mirror used by the tests binds them with ctypes.
 *
 * Conventions
 *   - every function returns an int32 status: 0 = OK, negative = error (PPR_E_*); the message is
 *     available from ppr_last_error().  Nothing throws or aborts.
 *   - handles are opaque, owned by the library until the matching destroy call.
 *   - host arrays passed in are borrowed for the duration of the call only; output buffers are
 *     caller-owned and caller-sized.  Pinned buffers from ppr_host_alloc() make the copies faster.
 *   - vertex ids are "mapped" ids in [0, n): the Neo4j original <-> mapped translation stays on the
 *     host side, exactly where Graph.toMappedNodeId / toOriginalNodeId sit in the reference.
 *   - a context is used by one thread at a time (the reference harness is single threaded).
 *   - there is NO CPU fallback: every query runs CUDA kernels on the context's device and fails with
 *     PPR_E_CUDA when no device is usable.
 *   - results are deterministic functions of (graph, arguments, seed) and do not depend on the number
 *     of walk shards: residues/reserves are accumulated in 2^-62 fixed point with integer atomics
 *     (DESIGN.md), walks use Philox4x32-10 keyed by (seed, start vertex, walk index).
 */
#ifndef PPRB200_H
#define PPRB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPR_OK 0
#define PPR_E_INVALID (-1) /* bad argument */
#define PPR_E_CUDA (-2)    /* CUDA runtime error / no device */
#define PPR_E_NOMEM (-3)   /* host or device allocation failed */
#define PPR_E_RANGE (-4)   /* graph exceeds supported ranges (n < 2^31, degree < 2^24, m < 2^40) */
#define PPR_E_CAPACITY (-5)/* caller buffer too small; required size reported through the count output */
#define PPR_E_STATE (-6)   /* call sequence error (e.g. export before any query) */
#define PPR_E_CALLBACK (-7)/* a caller-supplied hook (ppr_reduce_fn) reported failure */

/* flags */
#define PPR_RESTART_AT_WALK_START 0 /* default: reference behaviour, J/Monte_Carlo.java:87-90 */
#define PPR_RESTART_AT_SOURCE 1     /* dangling vertices send a walk back to the query source */
#define PPR_BALANCE_PER_SHARD 2     /* sharded FORA: balance push against walk_cost / nshards (walks are nshards times cheaper) */
#define PPR_E_COMM (-8)     /* NCCL unavailable or a collective failed */

#define PPR_FX_ONE (1ull << 62) /* fixed-point scale of the raw uint64 vectors */

typedef struct ppr_ctx ppr_ctx;
typedef struct ppr_graph ppr_graph;
typedef struct ppr_pairs ppr_pairs; /* result of the all-pair backward search */

/* Accuracy parameters the reference passes to the algorithm constructors (delta, pfail; min_delta for
 * FORA top-k).  Algo_Conf derives them from the graph (J/Algo_Conf.java:29-81); a NULL ppr_conf*, or a
 * field <= 0, selects exactly that derivation: delta = pfail = 1/n; top-k: delta0 = 1/k,
 * min_delta = 1/n, pfail = 1/n/n/ln(n / k) with integer n / k. */
typedef struct ppr_conf {
    double delta;
    double pfail;
    double min_delta;
} ppr_conf;

/* Work and timing counters of the last query (all phases run on the context's stream). */
typedef struct ppr_stats {
    int64_t pushes;        /* vertices pushed (V_p) */
    int64_t edges;         /* edges relaxed (E_p) */
    int64_t iters;         /* synchronous push iterations */
    int64_t rounds;        /* rmax-halving rounds (FORA) / delta rounds (FORA top-k) */
    int64_t walks;         /* random walks executed on this shard */
    int64_t walks_total;   /* random walks of the whole query (all shards) */
    int64_t steps;         /* random-walk moves executed on this shard */
    int64_t walk_sources;  /* vertices that started walks */
    uint64_t rmax_fx;      /* final push threshold, fixed point */
    uint64_t rsum_fx;      /* residue mass handed to the walk phase, fixed point */
    double omega;          /* omega of the query (J/Fora_Whole_Graph.java:87, J/Monte_Carlo.java:145) */
    float ms_push;         /* CUDA-event time of the push kernel(s) */
    float ms_walk;         /* CUDA-event time of the walk kernel(s) */
    float ms_select;       /* CUDA-event time of top-k selection */
    float ms_total;        /* CUDA-event time of the whole device part of the call */
    int32_t launches;      /* kernels launched by the call */
    float ms_comm;         /* CUDA-event time of the call's NCCL collectives (sharded entry points) */
} ppr_stats;

/* ---------------------------------------------------------------- lifecycle
 * Replaces the embedded-database lifecycle of the driver: PPR.createDb / registerShutdownHook /
 * shutDown (J/PPR.java:35-68).  `device` is the CUDA ordinal (one context per GPU, one process per
 * GPU in multi-GPU runs). */
int32_t ppr_init(int32_t device, ppr_ctx** out);
int32_t ppr_shutdown(ppr_ctx* ctx);
const char* ppr_last_error(ppr_ctx* ctx); /* ctx may be NULL: last error of a failed ppr_init */
int32_t ppr_device_info(ppr_ctx* ctx, int32_t* sm_count, int64_t* total_mem, char* name, int32_t name_cap);

/* Tunables that replace the wall-clock push/walk balancer of J/Fora_Whole_Graph.java:35,93-103:
 * the push keeps halving rmax while (vertices pushed + edges relaxed) < walk_cost * r_sum * omega and
 * rounds < max_rounds.  walk_cost < 0 (default) = automatic: 6 while the graph fits in the 126 MB L2, 24
 * beyond (measured B200 costs; the reference's 400 ns/walk against its Java push is about 32).
 * max_rounds default 64. */
int32_t ppr_set_balance(ppr_ctx* ctx, double walk_cost, int32_t max_rounds);

/* Frontier-mode switch of the push kernels: an iteration whose frontier has at least this many out-edges
 * scatters with plain reductions and the next frontier is found by a dense sweep; smaller frontiers use
 * returning atomics and exact threshold-crossing detection.  Results are bit-identical either way.
 * < 0 restores the default (n/4); 0 = always sweep; INT64_MAX = never sweep.
 * hot_min_indeg (>= 0; < 0 keeps the current value, default 64): minimum in-degree for a vertex to get one of
 * the 4096 hub slots whose incoming adds are staged in shared memory; applies to graphs created afterwards. */
int32_t ppr_set_tuning(ppr_ctx* ctx, int64_t scan_threshold_edges, int32_t hot_min_indeg);

/* ppr_fora_batch: 2 keeps two queries in flight on two streams (each cooperative kernel takes one CTA per SM), 1 runs
 * them back to back, 0 (default) picks 2 while the graph fits in the L2.  Results are identical. */
int32_t ppr_set_batch_streams(ppr_ctx* ctx, int32_t streams);

/* All-pair backward search: targets advanced together per kernel launch (0 = automatic: sized so that the
 * batch's residue matrix stays near the L2 capacity). */
int32_t ppr_set_bwd_batch(ppr_ctx* ctx, int32_t targets_per_batch);

/* Wait for everything queued on the context's streams. */
int32_t ppr_device_sync(ppr_ctx* ctx);

/* Pinned host memory for result buffers (optional). */
int32_t ppr_host_alloc(ppr_ctx* ctx, int64_t bytes, void** out);
int32_t ppr_host_free(ppr_ctx* ctx, void* p);

/* ---------------------------------------------------------------- graph export
 * Replaces PPR.setupAdjMatrix (J/PPR.java:136-152: GraphLoader ... load(HeavyGraphFactory)) and the
 * Graph accessors used on the hot path (nodeCount J/PPR.java:129, degree, getTarget,
 * forEachRelationship; call sites in SURVEY.md section 8b).  Input is the canonical CSR of the
 * topology: row_ptr[n+1] (int64), col_idx[m] (int32 mapped target ids, any order, parallel edges
 * kept).  The library copies it to the device and builds out-degrees, the packed vertex table, the
 * CSC and reciprocal degrees there. */
int32_t ppr_graph_create(ppr_ctx* ctx, int64_t n, int64_t m, const int64_t* row_ptr, const int32_t* col_idx,
                         ppr_graph** out);
int32_t ppr_graph_destroy(ppr_graph* g);
int64_t ppr_graph_node_count(const ppr_graph* g); /* Graph.nodeCount() */
int64_t ppr_graph_rel_count(const ppr_graph* g);  /* rel_amount, J/PPR.java:131 */
/* Copy the device-built CSC back (tests / export): col_ptr[n+1], row_idx[m], sources ascending per column. */
int32_t ppr_graph_export_csc(ppr_graph* g, int64_t* col_ptr, int32_t* row_idx);

/* ---------------------------------------------------------------- single-source queries
 * All outputs are dense length-n vectors of doubles unless stated; stats may be NULL. */

/* Forward_Push.computeWholeGraphPPR(src, rmax), J/Forward_Push.java:63-142.  residue_out and rsum_out
 * may be NULL.  rsum_out receives the residue mass left (getUpdatedRsum, J/Forward_Push.java:252). */
int32_t ppr_forward_push(ppr_graph* g, int32_t src, double alpha, double rmax, double* reserve_out,
                         double* residue_out, double* rsum_out, ppr_stats* stats);

/* Monte_Carlo.computeWholeGraphPPR(src, eps), J/Monte_Carlo.java:136-158 (delta = pfail = 1/n,
 * J/Algo_Conf.java:29-35). */
int32_t ppr_monte_carlo(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed, const ppr_conf* conf,
                        double* ppr_out, ppr_stats* stats);

/* Fora_Whole_Graph.computeWholeGraphPPR(src, eps), J/Fora_Whole_Graph.java:82-146 (delta = pfail = 1/n,
 * J/Algo_Conf.java:45-53). */
int32_t ppr_fora(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed, int32_t flags,
                 const ppr_conf* conf, double* ppr_out, ppr_stats* stats);

/* Fora_Topk.computeTopKPPR(src, _, eps) followed by getTopKNodeIds(k), J/Fora_Topk.java:102-184,82-99,
 * parameters per J/Algo_Conf.java:71-81.  Returns every vertex whose estimate is >= the k-th largest
 * (ties included, so *count_out may exceed k), sorted by descending estimate, ties by ascending id.
 * cap is the capacity of ids_out / vals_out; on PPR_E_CAPACITY *count_out holds the needed size. */
int32_t ppr_fora_topk(ppr_graph* g, int32_t src, double alpha, double eps, int32_t k, uint64_t seed, int32_t flags,
                      const ppr_conf* conf, int32_t* ids_out, double* vals_out, int32_t cap, int32_t* count_out, ppr_stats* stats);

/* Algo_Util.kth_ppr + retrieveTopK + getTopKNodeIds over an arbitrary dense vector
 * (J/Algo_Util.java:32-73, J/Forward_Push.java:394-429).  Only positive entries count as present. */
int32_t ppr_topk(ppr_ctx* ctx, const double* vec, int64_t n, int32_t k, int32_t* ids_out, double* vals_out,
                 int32_t cap, int32_t* count_out, double* kth_out);

/* Power_Method.computeWholeGraphPPR(src), J/Power_Method.java:44-101 (`iters` synchronous rounds; the
 * reference uses 100).  Double arithmetic, deterministic pull over the CSC. */
int32_t ppr_power_iteration(ppr_graph* g, int32_t src, double alpha, int32_t iters, double* out);

/* Gen_Util.computeError(algo, pm, err_type, k), J/Gen_Util.java:259-326, evaluated on the device.
 * ppr_metrics_keep snapshots the dense result of the LAST query of this graph (power iteration, FORA, forward push,
 * Monte-Carlo, FORA top-k estimate) as doubles in HBM: slot 0 = ground truth (pm), slot 1 = estimate (algo).
 * ppr_metrics_compute returns MAX_ERR (max |est - truth| over the vertices the ground truth reached, :306-321),
 * PRECISION (:273-281: hits / size of the ground-truth top-k, ties included on both sides) and NDCG (:282-300) for k;
 * any output pointer may be NULL; k = 0 computes MAX_ERR only. */
int32_t ppr_metrics_keep(ppr_graph* g, int32_t slot);
int32_t ppr_metrics_compute(ppr_graph* g, int32_t k, double* max_err, double* precision, double* ndcg);

/* Throughput path: nq FORA queries back to back, results written to ppr_out[q*n .. (q+1)*n) while
 * the next query computes.  Query q uses seed + q.  stats_out (nullable) is an array of nq entries.
 * ppr_out == NULL leaves the results on the device (throughput measurements; ppr_last_raw reads the last one). */
int32_t ppr_fora_batch(ppr_graph* g, const int32_t* srcs, int32_t nq, double alpha, double eps, uint64_t seed,
                       int32_t flags, const ppr_conf* conf, double* ppr_out, ppr_stats* stats_out);

/* ---------------------------------------------------------------- backward search / all pairs */

/* Backward_Search.backward_search_whole_graph(target), J/Backward_Search.java:38-100: reserve_out[v] =
 * estimate of pi(v, target). */
int32_t ppr_backward_search(ppr_graph* g, int32_t target, double alpha, double rmax, double* reserve_out,
                            ppr_stats* stats);

/* Base_Whole_Graph.preprocessing(threshold, k), J/Base_Whole_Graph.java:58-164, compute part: backward
 * search from every target in [t_lo, t_hi) (a rank's block in multi-GPU runs), keeping pi(v,t) >=
 * threshold, grouped by source v.  k < 0 keeps everything with targets ascending; k >= 0 keeps the
 * entries >= the k-th largest per source, sorted by descending value (ties: ascending target). */
int32_t ppr_backward_all_pairs(ppr_graph* g, double alpha, double threshold, int32_t k, int32_t t_lo, int32_t t_hi,
                               ppr_pairs** out, ppr_stats* stats);
int64_t ppr_pairs_count(const ppr_pairs* p);
/* src_ptr[n+1], tgt[count], val[count] (val_fx may be NULL; raw fixed-point values otherwise). */
int32_t ppr_pairs_export(const ppr_pairs* p, int64_t* src_ptr, int32_t* tgt, double* val, uint64_t* val_fx);
int32_t ppr_pairs_destroy(ppr_pairs* p);

/* ---------------------------------------------------------------- sharded execution (multi-GPU)
 * Walks shard over GPUs (one process and one context per GPU, graph replicated).  Each rank runs the
 * (deterministic) push, then only the walks whose global index falls in
 * [W*shard/nshards, W*(shard+1)/nshards), and leaves its raw fixed-point partial vector on the
 * device: rank 0's partial carries the push reserve, the others only walk contributions, so an
 * integer all-reduce(sum) over ranks (NCCL, ncclUint64 / int64 bit pattern) gives exactly the
 * single-GPU vector.  ppr_result_devptr exposes the device buffer for that collective;
 * ppr_result_export converts it to doubles afterwards. */
int32_t ppr_fora_shard(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed, int32_t flags,
                       const ppr_conf* conf, int32_t shard, int32_t nshards, ppr_stats* stats);
int32_t ppr_monte_carlo_shard(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed,
                              const ppr_conf* conf, int32_t shard, int32_t nshards, ppr_stats* stats);

/* FORA top-k with the walks of every delta-round sharded (BASELINE configs[2]: "walks sharded 1/2/4/8 B200";
 * the walk loop of J/Fora_Topk.java:161-167).  Every rank replays the deterministic incremental push, runs its
 * slice of the round's walks into its estimate vector (shard 0's carries the push reserve, J/Fora_Topk.java:143),
 * and then calls `reduce` ONCE PER ROUND: the hook must sum the n raw uint64 (2^-62 fixed point) values at device
 * address dev_ptr element-wise over all shards, in place (two's-complement int64 sum: ncclInt64 / ncclSum), and
 * return 0 when the sum is complete on the device (non-zero aborts the query with PPR_E_CALLBACK).  The library's
 * stream is idle when the hook runs.  Selection and the stop test (J/Fora_Topk.java:173-178) then run on every
 * rank on identical data, so all ranks take the same number of rounds and return the same ids and values --
 * bit-identical to ppr_fora_topk.  reduce may be NULL only when nshards == 1. */
typedef int32_t (*ppr_reduce_fn)(void* user, uint64_t dev_ptr, int64_t n);
int32_t ppr_fora_topk_shard(ppr_graph* g, int32_t src, double alpha, double eps, int32_t k, uint64_t seed, int32_t flags,
                            const ppr_conf* conf, int32_t shard, int32_t nshards, ppr_reduce_fn reduce, void* user,
                            int32_t* ids_out, double* vals_out, int32_t cap, int32_t* count_out, ppr_stats* stats);
/* ---------------------------------------------------------------- multi-GPU: NCCL behind the ABI
 * One process (or thread) and one context per GPU, graph replicated.  The library owns the communicator and
 * issues its collectives on its own stream; NCCL is bound at run time (an NCCL the host process has already
 * loaded, else $PPR_NCCL_LIB, else libnccl.so.2), so single-GPU use needs no NCCL at all.
 *   rank 0:    ppr_comm_unique_id(id)  ->  ship the 128 bytes to every rank (any side channel)
 *   all ranks: ppr_comm_init(ctx, id, rank, world)  (collective)
 * The *_sharded entry points below are collective calls: every rank makes the same call with the same
 * arguments.  Walks shard by global walk index, results are summed with ncclAllReduce(ncclUint64, ncclSum) --
 * integer sums, so every rank ends with the same bits, identical to the single-GPU result unless
 * PPR_BALANCE_PER_SHARD is set. */
int32_t ppr_comm_unique_id(void* id_out /* 128 bytes */);
int32_t ppr_comm_init(ppr_ctx* ctx, const void* id /* 128 bytes */, int32_t rank, int32_t world);
int32_t ppr_comm_destroy(ppr_ctx* ctx);
int32_t ppr_comm_info(ppr_ctx* ctx, int32_t* rank, int32_t* world, int32_t* nccl_version);
/* Barrier + element-wise MAX of `count` doubles over the ranks (timing plumbing for multi-rank harnesses). */
int32_t ppr_comm_max_f64(ppr_ctx* ctx, double* inout, int32_t count);

/* Fora_Whole_Graph.computeWholeGraphPPR with the walk loop (J/Fora_Whole_Graph.java:119-140) sharded over the
 * communicator's ranks and one all-reduce of the reserve vector.  ppr_out may be NULL (result stays on the device). */
int32_t ppr_fora_sharded(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed, int32_t flags,
                         const ppr_conf* conf, double* ppr_out, ppr_stats* stats);
/* Monte_Carlo.computeWholeGraphPPR with the walk loop (J/Monte_Carlo.java:149-154) sharded + one all-reduce of the counts. */
int32_t ppr_monte_carlo_sharded(ppr_graph* g, int32_t src, double alpha, double eps, uint64_t seed, const ppr_conf* conf,
                                double* ppr_out, ppr_stats* stats);
/* ppr_fora_topk_shard with the library's own all-reduce as the per-round reduction (J/Fora_Topk.java:155-168). */
int32_t ppr_fora_topk_sharded(ppr_graph* g, int32_t src, double alpha, double eps, int32_t k, uint64_t seed, int32_t flags,
                              const ppr_conf* conf, int32_t* ids_out, double* vals_out, int32_t cap, int32_t* count_out, ppr_stats* stats);
/* Base_Whole_Graph.preprocessing with the target loop (J/Base_Whole_Graph.java:76-92) block-partitioned over the
 * ranks: no communication until the final gather (ncclAllGather of the per-source counts, then one variable-length
 * exchange straight into the grouped layout).  Every rank receives the complete index. */
int32_t ppr_backward_all_pairs_sharded(ppr_graph* g, double alpha, double threshold, int32_t k, ppr_pairs** out, ppr_stats* stats);

/* Device pointer and length (elements) of the raw uint64 result vector of the last query on g. */
int32_t ppr_result_devptr(ppr_graph* g, uint64_t* devptr_out, int64_t* n_out);
/* Convert the raw result vector to doubles (after an external all-reduce if sharded).
 * divisor = 0: value * 2^-62 (push / FORA); divisor > 0: count / divisor (Monte-Carlo omega). */
int32_t ppr_result_export(ppr_graph* g, double divisor, double* out);

/* Terminal vertex of walk `j` started at `start` (Monte_Carlo.random_walk / random_walk_no_zero_hop,
 * J/Monte_Carlo.java:60-133), for `count` consecutive walk indices starting at j0; runs the walk
 * kernel's device function one walk per thread. */
int32_t ppr_walk_terminals(ppr_graph* g, int32_t start, int32_t restart, uint64_t j0, int64_t count, uint64_t seed,
                           uint64_t tag, double alpha, int32_t no_zero_hop, int32_t* terminals_out);

const char* ppr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PPRB200_H */
