// pinned.cuh -- float32 operation sequences pinned to the reference build (shared by the surfel and
// the EWA preprocess kernels).
#pragma once
#include "common.cuh"

namespace gsr {

__device__ __forceinline__ void load16(const float* __restrict__ src, float* dst) {
#pragma unroll
    for (int i = 0; i < 16; i++) dst[i] = __ldg(src + i);
}

// ---- pinned float32 sequences ------------------------------------------------------------
// radius = ceil(sqrt(p^2 - f.(T o T))) cancels ~4 digits in global pixel coordinates and the
// sort key is the raw bits of the view depth, so radii / tile lists / ordering only match the
// reference if T, the AABB and p_view are rounded EXACTLY as its build rounds them.  The
// sequences below were read off the SASS nvcc 12.9 emits for the unmodified reference
// (cuobjdump of oracle/_ref/libref_surfel.so): every  a*x + b*y + c*z (+ d)  is evaluated as
//     fma(c, z, fma(a, x, rn(b*y)))  (+ d with a separate add),
// and explicit intrinsics keep the compiler from re-contracting them here.
__device__ __forceinline__ float dot_yxz(float a, float x, float b, float y, float c, float z) {
    return __fmaf_rn(c, z, __fmaf_rn(a, x, __fmul_rn(b, y)));
}
__device__ __forceinline__ float3 xform43_pinned(const float* __restrict__ m, float3 p) {
    return make_float3(__fadd_rn(dot_yxz(p.x, m[0], p.y, m[4], p.z, m[8]), m[12]),
                       __fadd_rn(dot_yxz(p.x, m[1], p.y, m[5], p.z, m[9]), m[13]),
                       __fadd_rn(dot_yxz(p.x, m[2], p.y, m[6], p.z, m[10]), m[14]));
}
__device__ __forceinline__ float3 xformvec43_pinned(const float* __restrict__ m, float3 p) {
    return make_float3(dot_yxz(p.x, m[0], p.y, m[4], p.z, m[8]), dot_yxz(p.x, m[1], p.y, m[5], p.z, m[9]),
                       dot_yxz(p.x, m[2], p.y, m[6], p.z, m[10]));
}
}  // namespace gsr
