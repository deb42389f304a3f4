// Microbenchmark (design input, not product code): random RED.64 / ATOM.64 / LDG.64 throughput on B200 for
// footprints of 8 MB .. 1 GB, with the index stream read coalesced like col_idx in the push kernel.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long ull;
__global__ void k_init(uint32_t* idx, size_t m, uint32_t mask, uint64_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t z = (i + seed) * 0x9E3779B97F4A7C15ull; z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; z ^= z >> 31;
        idx[i] = (uint32_t)z & mask;
    }
}
__global__ void k_hot(uint32_t* idx, size_t m, int h) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t z = (i + 12345) * 0xD6E8FEB86659FD93ull; z ^= z >> 32; z *= 0xD6E8FEB86659FD93ull; z ^= z >> 32;
        if ((z & ((1ull << h) - 1)) == 0) idx[i] = 7;
    }
}
template <int MODE>
__global__ void __launch_bounds__(256) k_op(const uint32_t* __restrict__ idx, size_t m, ull* data, ull* sink) {
    ull acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < m; i += 4 * stride) {
        uint32_t u[4];
#pragma unroll
        for (int k = 0; k < 4; k++) u[k] = (i + k * stride < m) ? __ldg(&idx[i + k * stride]) : 0xffffffffu;
#pragma unroll
        for (int k = 0; k < 4; k++) if (u[k] != 0xffffffffu) {
            if (MODE == 0) atomicAdd(&data[u[k]], 3ull);                 // RED
            else if (MODE == 1) acc += atomicAdd(&data[u[k]], 3ull);     // ATOM with return
            else if (MODE == 2) acc += data[u[k]];                       // gather
            else if (MODE == 3) atomicAdd((unsigned*)&data[u[k]], 3u);   // RED.32
        }
    }
    if (acc == 0x1234567) *sink = acc;
}
int main() {
    const size_t m = (size_t)1 << 27;  // 134M ops
    uint32_t* idx; ull* data; ull* sink;
    cudaMalloc(&idx, m * 4); cudaMalloc(&data, (size_t)1 << 30); cudaMalloc(&sink, 8);
    cudaMemset(data, 0, (size_t)1 << 30);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const char* names[4] = {"RED.64", "ATOM.64+ret", "LDG.64 gather", "RED.32"};
    for (int lg = 20; lg <= 27; lg += (lg < 24 ? 2 : 1)) {  // 2^lg elements of 8 B
        k_init<<<1184, 256>>>(idx, m, (1u << lg) - 1, 77);
        for (int mode = 0; mode < 4; mode++) {
            for (int blocks : {148 * 4, 148 * 8}) {
                float best = 1e9;
                for (int rep = 0; rep < 3; rep++) {
                    cudaEventRecord(a);
                    if (mode == 0) k_op<0><<<blocks, 256>>>(idx, m, data, sink);
                    if (mode == 1) k_op<1><<<blocks, 256>>>(idx, m, data, sink);
                    if (mode == 2) k_op<2><<<blocks, 256>>>(idx, m, data, sink);
                    if (mode == 3) k_op<3><<<blocks, 256>>>(idx, m, data, sink);
                    cudaEventRecord(b); cudaEventSynchronize(b);
                    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
                }
                printf("footprint %5zu MB  %-14s blocks/SM %d : %7.3f ms  %6.1f Gop/s\n", ((size_t)8 << lg) >> 20, names[mode], blocks / 148, best, m / best / 1e6);
            }
        }
    }
    // hot spots: a fraction 1/2^h of the ops goes to ONE address (a hub), the rest is random over 8 MB
    for (int h : {4, 6, 8, 10, 12}) {
        k_init<<<1184, 256>>>(idx, m, (1u << 20) - 1, 99);
        k_hot<<<1184, 256>>>(idx, m, h);
        for (int mode : {0, 2}) {
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(a);
                if (mode == 0) k_op<0><<<148 * 4, 256>>>(idx, m, data, sink);
                if (mode == 2) k_op<2><<<148 * 4, 256>>>(idx, m, data, sink);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            }
            printf("hot 1/%d to one address, 8 MB footprint  %-14s : %7.3f ms  %6.1f Gop/s   (%.2f ns per hot op)\n", 1 << h, names[mode], best, m / best / 1e6, best * 1e6 / (m >> h));
        }
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
