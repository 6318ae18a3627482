// Developer micro-benchmark (not a pytest): throughput of the tile-counter atomics of the binning pass.
// N atomics on T counters spaced `stride` words apart, random counter per atomic (as Gaussians hit tiles):
//   red  : result unused (RED)            ret : result used (returning ATOM, as scatter_keys needs)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tests/micro/atomic_bench.bin tests/micro/atomic_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <bool RET, int PER>
__global__ void k(uint32_t* ctr, int T, int stride, uint32_t* sink, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const uint32_t t = hash(i * PER + j) % (uint32_t)T;
        if (RET) acc += atomicAdd(&ctr[(size_t)t * stride], 1u);
        else atomicAdd(&ctr[(size_t)t * stride], 1u);
    }
    if (RET) sink[i] = acc;
}

template <bool RET, int PER>
float run(uint32_t* ctr, int T, int stride, uint32_t* sink, int n) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<RET, PER><<<(n + 255) / 256, 256>>>(ctr, T, stride, sink, n);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<RET, PER><<<(n + 255) / 256, 256>>>(ctr, T, stride, sink, n);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}

int main() {
    const int n = 2000000;          // threads (Gaussians), PER atomics each
    uint32_t *ctr, *sink;
    cudaMalloc(&ctr, (size_t)32 * 8 * 6700 * 4); cudaMalloc(&sink, n * 4);
    cudaMemset(ctr, 0, (size_t)32 * 8 * 6700 * 4);
    const int Ts[] = {6700, 6700 * 4, 6700 * 8};
    const int strides[] = {32, 8, 1};
    for (int T : Ts)
        for (int st : strides) {
            if ((size_t)T * st > (size_t)32 * 8 * 6700) continue;
            const float a = run<false, 3>(ctr, T, st, sink, n), b = run<true, 3>(ctr, T, st, sink, n);
            printf("T=%6d stride=%2d words: 6M red %.1f us (%.1f G/s)   6M returning %.1f us (%.1f G/s)\n", T, st, a * 1e3,
                   6e6 / a / 1e6, b * 1e3, 6e6 / b / 1e6);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
