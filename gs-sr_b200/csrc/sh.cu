// sh.cu -- spherical-harmonics colour (forward) and its gradient (backward) as warp-cooperative,
// fully coalesced kernels shared by all three rasterizers, for sm_100a.
//
// Result contract = computeColorFromSH (S/cuda_rasterizer/forward.cu:20-71, identical in G/ and L/) and its
// backward (S/cuda_rasterizer/backward.cu:20-139): basis, +0.5, clamp at 0 with per-channel `clamped` flags that
// zero the gradient, dL/dmean through the normalised view direction.  The math lives in sh.cuh.
//
// Why separate kernels: the coefficient rows are 3*M floats (192 B at degree 3) per Gaussian.  A
// thread-per-Gaussian kernel reads / writes them with a 192-byte stride between lanes (48 partially used
// sectors per warp instruction; the backward with SH took 1.16 ms at P = 2 M against 0.13 ms without).
// Here a warp owns 32 consecutive Gaussians = ONE contiguous 32*3M-float run in HBM: it is staged into
// shared memory with 128-byte coalesced loads (row stride 3M+1 floats -> conflict-free per-lane access),
// lane g works on Gaussian g, the backward overwrites the staged coefficients with dL/dsh in place and the
// run is written back with coalesced stores.  Rows of culled Gaussians and coefficients above the active
// degree are written as zeros, so the output needs no memset.  HBM traffic = 3M*4 B read (+ 3M*4 B written in
// the backward) per Gaussian, each byte touched once.
#include "common.cuh"
#include "sh.cuh"

namespace gsr {

constexpr int SH_WARPS = 4;
constexpr int SH_MAX_M = 16;
constexpr int SH_ROW = 3 * SH_MAX_M + 1;

// rows * row consecutive floats -> buf[g * (row + 1) + j].  ALL global loads are issued before the first
// shared store (12 independent 128-bit loads per lane at degree 3), so a warp pays one HBM latency, not 48.
__device__ __forceinline__ void sh_stage_in(float* buf, const float* __restrict__ src, int row, int rows, int lane) {
    if (row == 48) {
        const float4* src4 = reinterpret_cast<const float4*>(src);     // run starts at base*192 B: 16-byte aligned
        const int total4 = rows * 12;
        float4 v[12];
#pragma unroll
        for (int t = 0; t < 12; t++) {
            const int i4 = lane + 32 * t;
            v[t] = i4 < total4 ? __ldg(src4 + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int t = 0; t < 12; t++) {
            const int i4 = lane + 32 * t;
            const int g = i4 / 12, j = (i4 - g * 12) * 4;
            float* d = buf + g * 49 + j;
            d[0] = v[t].x; d[1] = v[t].y; d[2] = v[t].z; d[3] = v[t].w;
        }
        return;
    }
    const int total = rows * row;
    int g = lane / row, j = lane - g * row;              // (g, j) of element i, advanced by 32 per step
    for (int i = lane; i < total; i += 32) {
        buf[g * (row + 1) + j] = __ldg(src + i);
        j += 32;
        while (j >= row) { j -= row; g++; }
    }
}

__device__ __forceinline__ void sh_stage_out(const float* buf, float* __restrict__ dst, int row, int rows, int lane) {
    if (row == 48) {
        float4* dst4 = reinterpret_cast<float4*>(dst);
        const int total4 = rows * 12;
#pragma unroll
        for (int t = 0; t < 12; t++) {
            const int i4 = lane + 32 * t;
            const int g = i4 / 12, j = (i4 - g * 12) * 4;
            const float* d = buf + g * 49 + j;
            if (i4 < total4) dst4[i4] = make_float4(d[0], d[1], d[2], d[3]);
        }
        return;
    }
    const int total = rows * row;
    int g = lane / row, j = lane - g * row;
    for (int i = lane; i < total; i += 32) {
        dst[i] = buf[g * (row + 1) + j];
        j += 32;
        while (j >= row) { j -= row; g++; }
    }
}

__global__ void __launch_bounds__(SH_WARPS * 32)
sh_forward_kernel(int P, int D, int M, const float* __restrict__ means3D, const float* __restrict__ campos,
                  const float* __restrict__ shs, const int* __restrict__ radii, float* __restrict__ rgb,
                  uint8_t* __restrict__ clamped) {
    __shared__ float sbuf[SH_WARPS][32 * SH_ROW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int base = (blockIdx.x * SH_WARPS + warp) * 32;
    if (base >= P) return;
    const int rows = min(32, P - base), row = 3 * M;
    const int idx = base + lane;
    const bool visible = lane < rows && radii[idx] > 0;
    if (!__any_sync(0xffffffffu, visible)) return;
    float* buf = sbuf[warp];
    sh_stage_in(buf, shs + (size_t)base * row, row, rows, lane);
    __syncwarp();
    if (!visible) return;
    const float3 p = make_float3(__ldg(means3D + 3 * (size_t)idx), __ldg(means3D + 3 * (size_t)idx + 1), __ldg(means3D + 3 * (size_t)idx + 2));
    const float3 cp = make_float3(__ldg(campos), __ldg(campos + 1), __ldg(campos + 2));
    uint8_t cl[3];
    const float3 c = sh_to_rgb(D, M, p, cp, buf + lane * (row + 1), cl);
    rgb[3 * (size_t)idx + 0] = c.x; rgb[3 * (size_t)idx + 1] = c.y; rgb[3 * (size_t)idx + 2] = c.z;
    clamped[3 * (size_t)idx + 0] = cl[0]; clamped[3 * (size_t)idx + 1] = cl[1]; clamped[3 * (size_t)idx + 2] = cl[2];
}

// dL_dcolor: (P,3) per-Gaussian colour gradient (already reduced); dL_dmean3D is ACCUMULATED into.
__global__ void __launch_bounds__(SH_WARPS * 32)
sh_backward_kernel(int P, int D, int M, const float* __restrict__ means3D, const float* __restrict__ campos,
                   const float* __restrict__ shs, const uint8_t* __restrict__ clamped, const int* __restrict__ radii,
                   const float* __restrict__ dL_dcolor, float* __restrict__ dL_dsh, float* __restrict__ dL_dmean3D) {
    __shared__ float sbuf[SH_WARPS][32 * SH_ROW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int base = (blockIdx.x * SH_WARPS + warp) * 32;
    if (base >= P) return;
    const int rows = min(32, P - base), row = 3 * M;
    const int idx = base + lane;
    const bool visible = lane < rows && radii[idx] > 0;
    float* buf = sbuf[warp];
    float* out = dL_dsh + (size_t)base * row;
    const int total = rows * row;
    if (!__any_sync(0xffffffffu, visible)) {          // nothing visible in this run: zero rows, no read
        if ((total & 3) == 0) {
            float4* o4 = reinterpret_cast<float4*>(out);
            for (int i = lane; i < (total >> 2); i += 32) o4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int i = lane; i < total; i += 32) out[i] = 0.f;
        }
        return;
    }
    sh_stage_in(buf, shs + (size_t)base * row, row, rows, lane);
    __syncwarp();
    if (lane < rows) {
        float* mine = buf + lane * (row + 1);
        if (visible) {
            const float3 p = make_float3(__ldg(means3D + 3 * (size_t)idx), __ldg(means3D + 3 * (size_t)idx + 1), __ldg(means3D + 3 * (size_t)idx + 2));
            const float3 cp = make_float3(__ldg(campos), __ldg(campos + 1), __ldg(campos + 2));
            const float3 dcol = make_float3(dL_dcolor[3 * (size_t)idx], dL_dcolor[3 * (size_t)idx + 1], dL_dcolor[3 * (size_t)idx + 2]);
            // in place: sh_to_rgb_bwd reads coefficient k before it writes dL/dsh k
            const float3 dm = sh_to_rgb_bwd_inplace(D, M, p, cp, mine, clamped + 3 * (size_t)idx, dcol);
            dL_dmean3D[3 * (size_t)idx + 0] += dm.x;
            dL_dmean3D[3 * (size_t)idx + 1] += dm.y;
            dL_dmean3D[3 * (size_t)idx + 2] += dm.z;
        } else {
            for (int j = 0; j < row; j++) mine[j] = 0.f;
        }
    }
    __syncwarp();
    sh_stage_out(buf, out, row, rows, lane);
}

cudaError_t launch_sh_forward(int P, int D, int M, const float* means3D, const float* campos, const float* shs,
                              const int* radii, float* rgb, uint8_t* clamped, cudaStream_t s) {
    if (P <= 0) return cudaSuccess;
    const int per_block = SH_WARPS * 32;
    sh_forward_kernel<<<(P + per_block - 1) / per_block, per_block, 0, s>>>(P, D, M, means3D, campos, shs, radii, rgb, clamped);
    return cudaGetLastError();
}

cudaError_t launch_sh_backward(int P, int D, int M, const float* means3D, const float* campos, const float* shs,
                               const uint8_t* clamped, const int* radii, const float* dL_dcolor, float* dL_dsh,
                               float* dL_dmean3D, cudaStream_t s) {
    if (P <= 0) return cudaSuccess;
    const int per_block = SH_WARPS * 32;
    sh_backward_kernel<<<(P + per_block - 1) / per_block, per_block, 0, s>>>(P, D, M, means3D, campos, shs, clamped, radii,
                                                                             dL_dcolor, dL_dsh, dL_dmean3D);
    return cudaGetLastError();
}

}  // namespace gsr
