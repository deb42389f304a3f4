/* pprgen.h -- C ABI of libpprgen.so: synthetic inputs for bench.py and the tests (host only, no CUDA).
 *
 * Kept OUT of libpprb200.so so that a process that only needs a graph (bench.py --impl reference, the CPU
 * oracle tests) never maps the product library.  Not part of the reference-facing boundary: the reference
 * reads its graph from a Neo4j store (J/PPR.java:136-152); BASELINE.json's configs name R-MAT shapes.
 */
#ifndef PPRGEN_H
#define PPRGEN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Seeded R-MAT generator (a, b, c; d = 1-a-b-c), `m` directed edges over `n` vertices (n need not be
 * a power of two: out-of-range endpoints are redrawn), self loops and duplicates kept, optional
 * random relabelling and symmetrisation (each generated edge also inserted reversed; m counts both).
 * Deterministic for a given argument list and independent of thread count.
 * row_ptr[n+1] and col_idx[m] are caller-allocated; rows come out in ascending target order.
 * Returns 0, or -1 on a bad argument. */
int32_t ppr_rmat_generate(int64_t n, int64_t m, double a, double b, double c, uint64_t seed, int32_t permute,
                          int32_t symmetric, int64_t* row_ptr, int32_t* col_idx);

#ifdef __cplusplus
}
#endif
#endif /* PPRGEN_H */
