/* pprb200_debug.h -- measurement and test hooks of libpprb200.so.
 *
 * NOT part of the reference-facing boundary (include/pprb200.h): nothing here corresponds to a method of the
 * reference.  bench.py uses the timers and the L2 flush; the parity tests use the raw-state readers, the stepped push
 * and the walk-variant switch; profiling scripts use the timeline.  Same conventions as pprb200.h.
 */
#ifndef PPRB200_DEBUG_H
#define PPRB200_DEBUG_H
#include "pprb200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Measurement helpers (bench.py), per context: CUDA events on the context's compute stream, the stream every kernel
 * of this library is launched on (after joining the download stream), and an L2 flush (a write of a
 * 256 MiB scratch buffer, larger than the 126 MB L2).  slot is 0..3. */
int32_t ppr_timer_record(ppr_ctx* ctx, int32_t slot);
int32_t ppr_timer_elapsed_ms(ppr_ctx* ctx, int32_t slot_a, int32_t slot_b, float* ms_out); /* synchronises */
int32_t ppr_flush_l2(ppr_ctx* ctx);


/* Walk kernel variant: -1 automatic (beyond the L2: variant 2), 0 plain loads, 1 L2 eviction-priority hints, 2 hints + the vertex
 * record read from the half-size table of 32-bit row pointers (falls back to 1 when m >= 2^32); 3 / 4 / 5 = as 2 with the neighbour
 * sectors at normal priority / an address-hashed 25 % / 15 % of them evict-last and the rest evict-first (experiments: +2.5 % on
 * FORA's 1.4 ms walk launches at R-MAT scale 22, -23 % on the 100 ms Monte-Carlo launch, so not the default).
 * Results are identical; the parity tests force each variant on graphs of any size. */
int32_t ppr_set_walk_variant(ppr_ctx* ctx, int32_t hint);

/* Direction switch of the forward push: a dense iteration whose frontier owns at least enter_share * m edges is followed
 * by pull-direction passes (every in-edge read once, no atomics) until the frontier owns less than exit_share * m.
 * enter_share < 0 switches the pull direction off; 0 / 0 forces it after every dense iteration.  Results are identical. */
int32_t ppr_set_pull(ppr_ctx* ctx, double enter_share, double exit_share);

/* ---------------------------------------------------------------- raw state of the last query (parity tests)
 * Fixed-point vectors exactly as the kernels left them: reserve (or FORA estimate / Monte-Carlo
 * counts) and residue.  Either pointer may be NULL. */
int32_t ppr_last_raw(ppr_graph* g, uint64_t* reserve_fx, uint64_t* residue_fx);

/* Stepped forward push for schedule tests: same device code as ppr_forward_push, one synchronous
 * iteration per call.  begin() seeds the frontier for (src, rmax); each step() returns the frontier
 * it pushed (ids in frontier_out, unordered; count in *count_out; 0 = converged). */
int32_t ppr_push_trace_begin(ppr_graph* g, int32_t src, double alpha, double rmax);
int32_t ppr_push_trace_step(ppr_graph* g, int32_t* frontier_out, int64_t cap, int64_t* count_out);

/* Per-iteration timeline of the last push kernel (profiling aid): 6 uint64 per synchronous iteration =
 * {t_start, t_after_phase_A, t_after_phase_B (GPU globaltimer, ns), frontier size, cumulative edges relaxed, mode}. */
int32_t ppr_debug_timeline(ppr_graph* g, uint64_t* out, int64_t cap_entries, int64_t* count_out);


#ifdef __cplusplus
}
#endif
#endif /* PPRB200_DEBUG_H */
