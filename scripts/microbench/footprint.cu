// Random 4-byte gathers and 8-byte REDs against the footprint of the table, in 8 MB steps around the L2 size:
// how much randomly accessed data does the 126 MB L2 of a B200 really keep?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o footprint footprint.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long ull;
__device__ __forceinline__ uint32_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return (uint32_t)x; }
template <int MODE>
__global__ void __launch_bounds__(256) k_op(const uint32_t* __restrict__ t32, ull* t64, size_t ops, uint32_t count, ull* sink) {
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gsize = (size_t)gridDim.x * blockDim.x;
    ull acc = 0;
    for (size_t i = gtid; i < ops; i += 4 * gsize) {
        uint32_t u[4];
#pragma unroll
        for (int k = 0; k < 4; k++) u[k] = (uint32_t)(((uint64_t)mix(i + k * gsize) * count) >> 32);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (i + k * gsize >= ops) break;
            if (MODE == 0) acc += t32[u[k]];
            else atomicAdd(&t64[u[k]], 3ull);
        }
    }
    if (acc == 0x1234567) *sink = acc;
}
int main() {
    const size_t ops = (size_t)1 << 28;
    uint32_t* t32; ull* t64; ull* sink;
    cudaMalloc(&t32, (size_t)512 << 20); cudaMalloc(&t64, (size_t)512 << 20); cudaMalloc(&sink, 8);
    cudaMemset(t32, 0, (size_t)512 << 20); cudaMemset(t64, 0, (size_t)512 << 20);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int mb : {16, 32, 40, 48, 56, 64, 72, 80, 96, 112, 128, 160, 192, 256}) {
        for (int mode = 0; mode < 2; mode++) {
            uint32_t count = (uint32_t)(((size_t)mb << 20) / (mode == 0 ? 4 : 8));
            float best = 1e9;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(a);
                if (mode == 0) k_op<0><<<148 * 8, 256>>>(t32, t64, ops, count, sink);
                else k_op<1><<<148 * 8, 256>>>(t32, t64, ops, count, sink);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            }
            printf("footprint %4d MB  %-16s %7.3f ms  %6.1f Gop/s\n", mb, mode == 0 ? "LDG.32 gather" : "RED.64", best, ops / best / 1e6);
        }
    }
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
