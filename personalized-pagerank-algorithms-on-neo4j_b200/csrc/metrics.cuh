// metrics.cuh -- Gen_Util.computeError on the device (J/Gen_Util.java:259-326): max |est - truth|, precision@k, NDCG@k of a
// result vector against a ground-truth vector, both resident in HBM -- so that accuracy can be reported beside the
// throughput at sizes where downloading and ranking 4 M doubles per query on the host would dominate.  Also the ordering
// of top-k survivors (value descending, id ascending), used by every top-k entry point instead of a host sort.
#pragma once
#include "common.cuh"

// MAX_ERR branch (J/Gen_Util.java:306-321): the reference walks the ENTRIES of the ground-truth map, i.e. the vertices
// the power method reached (truth != 0); a vertex missing from the estimate counts as 0.
// Non-negative doubles order like their bit patterns, so the maximum is an integer atomicMax.
__global__ void k_max_err(const double* __restrict__ truth, const double* __restrict__ est, i64 n, ull* __restrict__ out_bits) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    double m = 0.0;
    for (i64 v = gtid; v < n; v += gsize) {
        const double t = truth[v];
        if (t != 0.0) m = fmax(m, fabs(est[v] - t));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out_bits, (ull)__double_as_longlong(m));
}

// Survivors of the radix select (<= k + ties of them) into value-descending, id-ascending order: rank sort.
__global__ void k_rank_sort(const i32* __restrict__ ids, const u64* __restrict__ vals, i64 c, i32* __restrict__ out_ids, u64* __restrict__ out_vals) {
    const i64 gtid = (i64)blockIdx.x * blockDim.x + threadIdx.x, gsize = (i64)gridDim.x * blockDim.x;
    for (i64 i = gtid; i < c; i += gsize) {
        const u64 x = vals[i];
        const i32 id = ids[i];
        i64 rank = 0;
        for (i64 j = 0; j < c; j++) {
            const u64 y = __ldg(&vals[j]);
            rank += (y > x || (y == x && __ldg(&ids[j]) < id)) ? 1 : 0;
        }
        out_ids[rank] = id; out_vals[rank] = x;
    }
}

// PRECISION and NDCG branches (J/Gen_Util.java:264-300) from the two ordered top-k lists.  gnd_vals are the order-preserving
// bit images of the ground-truth doubles (k_import_f64).  One CTA; the sums run in list order like the reference's loops.
// out[0] = precision = |algo ids that are ground-truth top-k| / |ground-truth top-k|; out[1] = DCG / Zk.
__global__ void __launch_bounds__(256) k_topk_metrics(const i32* __restrict__ gnd_ids, const u64* __restrict__ gnd_vals, i64 cg_,
                                                      const i32* __restrict__ algo_ids, i64 ca, double* __restrict__ gain /*[ca] scratch*/, double* __restrict__ out) {
    __shared__ u32 s_hits;
    if (threadIdx.x == 0) s_hits = 0;
    __syncthreads();
    u32 hits = 0;
    for (i64 i = threadIdx.x; i < ca; i += blockDim.x) {
        const i32 id = algo_ids[i];
        double g = 0.0;      // ground-truth value of the algorithm's i-th vertex, 0 when it is not a ground-truth top-k vertex (:291-293)
        for (i64 j = 0; j < cg_; j++) if (__ldg(&gnd_ids[j]) == id) { g = __longlong_as_double((long long)__ldg(&gnd_vals[j])); hits++; break; }
        gain[i] = g;
    }
    if (hits) atomicAdd(&s_hits, hits);
    __syncthreads();
    if (threadIdx.x == 0) {
        const double ln2 = log(2.0);
        double zk = 0.0, dcg = 0.0;
        for (i64 i = 1; i <= cg_; i++) zk += (pow(2.0, __longlong_as_double((long long)gnd_vals[i - 1])) - 1.0) / log((double)i + 1.0) / ln2;
        for (i64 i = 1; i <= ca; i++) dcg += (pow(2.0, gain[i - 1]) - 1.0) / log((double)i + 1.0) / ln2;
        out[0] = cg_ ? (double)s_hits / (double)cg_ : 0.0;
        out[1] = zk != 0.0 ? dcg / zk : 0.0;
    }
}
