// Microbenchmark for the dense sweep of push phase A: what does each ingredient cost at n = 1 M / 4 M?
//   V0 read r[]                       V1 + vertex records
//   V2 + taker updates (r store, reserve load+store)      V3 + warp-aggregated compaction (1 atomic / 128 items) + staging stores
//   V4 compaction inside the tile's own slots (no atomic)  V5 = V4 + one result-less atomic per tile
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sweep sweep.cu ; run: ./sweep
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long ull;

template <int V>
__global__ void __launch_bounds__(512, 2) k_sweep(const ull* __restrict__ r_in, ull* r, const ull* __restrict__ vtx, ull* reserve,
                                                  ull* s_pack, ull* s_q, ull* counter, ull* sink, long n, ull thr) {
    const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gsize = (long)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31;
    const long gwarp = gtid >> 5, nwarps = gsize >> 5;
    ull acc = 0;
    for (long tile = gwarp * 128; tile < n; tile += nwarps * 128) {
        ull x[4], pk[4];
        bool take[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { long i = tile + k * 32 + lane; x[k] = i < n ? r[i] : 0; }
#pragma unroll
        for (int k = 0; k < 4; k++) { long i = tile + k * 32 + lane; pk[k] = (V >= 1 && x[k]) ? vtx[i] : 1; }
        unsigned ms[4], tot = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            take[k] = x[k] >= thr * (pk[k] & 0xff);
            acc += x[k] ^ pk[k];
            if (V >= 2 && take[k]) {
                long i = tile + k * 32 + lane;
                ull a = x[k] / 7;
                r[i] = x[k] - a;   // (value stays above thr so every repetition does the same work)
                reserve[i] += a;
            }
            ms[k] = __ballot_sync(0xffffffffu, take[k]);
            tot += __popc(ms[k]);
        }
        if (V >= 3 && tot) {
            ull base = 0;
            if (V == 3) {
                if (lane == 0) base = atomicAdd(counter, (ull)tot);
                base = __shfl_sync(0xffffffffu, base, 0);
                base %= (ull)n;       // the counter runs on across repetitions
            } else {
                base = (ull)tile;     // V4: slots inside the tile's own 128 entries, no atomic
                if (V == 5 && lane == 0) atomicAdd(counter, (ull)tot);   // V5: V4 + a result-less atomic (RED) per tile
            }
            unsigned off = (unsigned)base;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (take[k]) { unsigned slot = off + __popc(ms[k] & ((1u << lane) - 1)); s_pack[slot] = pk[k]; s_q[slot] = x[k]; }
                off += __popc(ms[k]);
            }
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

// Double-buffered variant (design probe): read r_cur[v] and the vertex record, zero r_cur[v], carry the residue
// (leftover of a taker, everything of a non-taker) to r_next[v] with one coalesced RED; takers also update reserve.
// No staging, no compaction: in the real kernel the taker would relax its edges right here.
template <int V>
__global__ void __launch_bounds__(512, 2) k_sweep2(ull* r_cur, ull* r_next, const ull* __restrict__ vtx, ull* reserve, ull* sink, long n, ull thr) {
    const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gsize = (long)gridDim.x * blockDim.x;
    const unsigned lane = threadIdx.x & 31;
    const long gwarp = gtid >> 5, nwarps = gsize >> 5;
    ull acc = 0;
    for (long tile = gwarp * 128; tile < n; tile += nwarps * 128) {
        ull x[4], pk[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { long i = tile + k * 32 + lane; x[k] = i < n ? r_cur[i] : 0; }
#pragma unroll
        for (int k = 0; k < 4; k++) { long i = tile + k * 32 + lane; pk[k] = x[k] ? vtx[i] : 1; }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            long i = tile + k * 32 + lane;
            if (!x[k]) continue;
            bool take = x[k] >= thr * (pk[k] & 0xff);
            ull carry = x[k];
            if (take) { ull a = x[k] / 7; if (V >= 1) reserve[i] += a; carry = 5; acc += a; }
            r_cur[i] = 0;
            atomicAdd(&r_next[i], carry);
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

__global__ void k_fill(ull* r, ull* vtx, long n, ull thr, unsigned pct) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned h = (unsigned)(i * 2654435761u) >> 8;
    unsigned d = 1 + (h % 30);
    vtx[i] = ((ull)i << 8) | d;
    r[i] = (h % 100) < pct ? thr * d * 1000000ull + 5 : (h % 3 ? thr / 2 : 0);
}

template <int V>
static void run(const char* name, long n, unsigned pct, ull* r, ull* vtx, ull* reserve, ull* sp, ull* sq, ull* counter, ull* sink) {
    const ull thr = 1000;
    k_fill<<<(unsigned)((n + 255) / 256), 256>>>(r, vtx, n, thr, pct);
    cudaMemset(reserve, 0, n * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 50;
    for (int w = 0; w < 3; w++) { cudaMemset(counter, 0, 8); k_sweep<V><<<296, 512>>>(r, r, vtx, reserve, sp, sq, counter, sink, n, thr); }
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) { k_sweep<V><<<296, 512>>>(r, r, vtx, reserve, sp, sq, counter, sink, n, thr); }
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("n %8ld takers %2u%%  %-44s %7.2f us per sweep\n", n, pct, name, ms * 1000 / reps);
}

template <int V>
static void run2(const char* name, long n, unsigned pct, ull* r, ull* r2, ull* vtx, ull* reserve, ull* sink) {
    const ull thr = 1000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 20;
    float tot = 0;
    for (int i = 0; i < reps + 3; i++) {     // refill every time: the sweep consumes r_cur
        k_fill<<<(unsigned)((n + 255) / 256), 256>>>(r, vtx, n, thr, pct);
        cudaMemset(r2, 0, n * 8);
        cudaEventRecord(e0);
        k_sweep2<V><<<296, 512>>>(r, r2, vtx, reserve, sink, n, thr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 3) tot += ms;
    }
    printf("n %8ld takers %2u%%  %-44s %7.2f us per sweep\n", n, pct, name, tot * 1000 / reps);
}

int main() {
    const long nmax = 1L << 24;
    ull *r, *vtx, *reserve, *sp, *sq, *counter, *sink;
    cudaMalloc(&r, nmax * 8); cudaMalloc(&vtx, nmax * 8); cudaMalloc(&reserve, nmax * 8);
    cudaMalloc(&sp, nmax * 8 * 60); cudaMalloc(&sq, 8); cudaFree(sq); cudaFree(sp);
    cudaMalloc(&sp, (nmax + 256) * 8); cudaMalloc(&sq, (nmax + 256) * 8); cudaMalloc(&counter, 8); cudaMalloc(&sink, 8);
    for (long n : {1L << 20, 1L << 22, 1L << 24}) {
        for (unsigned pct : {0u, 5u, 30u}) {
            run<0>("V0 read r", n, pct, r, vtx, reserve, sp, sq, counter, sink);
            run<1>("V1 + vertex records", n, pct, r, vtx, reserve, sp, sq, counter, sink);
            run<2>("V2 + taker updates (r, reserve)", n, pct, r, vtx, reserve, sp, sq, counter, sink);
            run<3>("V3 + compaction (1 ATOM / tile) + staging", n, pct, r, vtx, reserve, sp, sq, counter, sink);
            run<4>("V4 + compaction inside the tile, no atomic", n, pct, r, vtx, reserve, sp, sq, counter, sink);
            run<5>("V5 = V4 + one RED / tile", n, pct, r, vtx, reserve, sp, sq, counter, sink);
            run2<0>("D0 double-buffered: zero + carry RED", n, pct, r, sp, vtx, reserve, sink);
            run2<1>("D1 = D0 + reserve update of takers", n, pct, r, sp, vtx, reserve, sink);
        }
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
