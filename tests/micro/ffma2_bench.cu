// Developer micro-benchmark (not a pytest): issue cost of packed FFMA2 against scalar FFMA on sm_100a.
// Each thread runs NCHAIN independent dependency chains; a warp-instruction count per second is printed for
//   (a) scalar FFMA, (b) FFMA2 (two FMAs per instruction), (c) a 1:1 mix of FFMA2 and integer ALU work.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tests/micro/ffma2_bench.bin tests/micro/ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
constexpr int NCHAIN = 8, ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
    float2 v[NCHAIN];
    unsigned u[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) { v[i] = make_float2(threadIdx.x + i, threadIdx.x - i); u[i] = threadIdx.x * 7 + i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NCHAIN; i++) {
            if (MODE == 0) { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b); }       // 2 FFMA
            if (MODE == 1) v[i] = ffma2(v[i], make_float2(a, a), make_float2(b, b));           // 1 FFMA2
            if (MODE == 2) { v[i] = ffma2(v[i], make_float2(a, a), make_float2(b, b)); u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e37u; }
            if (MODE == 3) { v[i].x = fmaf(v[i].x, a, b); v[i].y = fmaf(v[i].y, a, b); u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e37u; }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) s += v[i].x + v[i].y + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double fma_per_thread_iter, float* out) {
    const int grid = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<MODE><<<grid, 256>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double fmas = (double)grid * 256 * ITERS * NCHAIN * fma_per_thread_iter;
    printf("%-28s %.3f ms  %.1f TFMA/s (%.1f TFLOP/s)\n", name, ms, fmas / ms / 1e9, 2 * fmas / ms / 1e9);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("scalar FFMA x2", 2, out);
    run<1>("FFMA2", 2, out);
    run<3>("scalar FFMA x2 + 2 ALU", 2, out);
    run<2>("FFMA2 + 2 ALU", 2, out);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
