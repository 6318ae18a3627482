// render_common.cuh -- pieces shared by the forward and backward surfel render kernels:
// batch staging, the warp-block cull of one record, and the per-(pixel, splat) evaluation.
#pragma once
#include "common.cuh"
#include "async_copy.cuh"
#include "cull.cuh"

namespace gsr {

#ifndef GSR_RBATCH
#define GSR_RBATCH 224                            // measured at cfg-B: 64 -> 1132 us, 128 -> 1097, 192 -> 1086, 224 -> 1080 (43 KB of static smem)
#endif
constexpr int RBATCH = GSR_RBATCH;                // record entries per shared-memory stage of the forward kernels (multiple of 32)
static_assert(RBATCH % 32 == 0 && 2 * 6 * RBATCH * 16 <= 47 * 1024, "forward stage size");
constexpr uint32_t FULLMASK = 0xffffffffu;
constexpr float MSCALE = FAR_N / (FAR_N - NEAR_N);              // S/forward.cu:396
constexpr float DMD = (FAR_N * NEAR_N) / (FAR_N - NEAR_N);      // S/backward.cu:348
constexpr float NEG_HALF_LOG2E = -0.72134752044448170368f;      // exp(-rho/2) = 2^(rho * this)

__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Packed FP32 pairs (PTX *.f32x2, SASS FFMA2 / FMUL2 / FADD2 on sm_100a): one issue slot for two IEEE round-to-nearest
// operations, each half bit-identical to the scalar fmaf / mul / add.  The render kernels are issue-bound (75-83 % of the
// issue slots, FMA pipe < 40 % busy), so halving the instruction count of the two-wide pieces is what counts.  The
// (x, y) / (z, w) halves of a 128-bit shared-memory load are already aligned register pairs; a scalar second operand
// uses the instruction's broadcast form (no MOV).  GSR_FFMA2=0 builds the scalar sequences for A/B runs.
#ifndef GSR_FFMA2
#define GSR_FFMA2 1
#endif
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
#if GSR_FFMA2
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
#if GSR_FFMA2
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
#if GSR_FFMA2
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\t"
        "mov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__device__ __forceinline__ float2 ffma2s(float2 a, float s, float2 c) { return ffma2(a, make_float2(s, s), c); }
__device__ __forceinline__ float2 fmul2s(float2 a, float s) { return fmul2(a, make_float2(s, s)); }

// One thread: arm the stage's mbarrier and launch the six plane copies of entries
// [first, first+count) of this tile's list.
template <int NB>
__device__ __forceinline__ void issue_batch(float4 (*dst)[NB], const float4* __restrict__ src, size_t pstride,
                                            int first, int count, uint64_t* bar) {
    const uint32_t bytes = (uint32_t)count * 16u;
    mbar_expect_tx(bar, bytes * REC_PLANES);
#pragma unroll
    for (int pl = 0; pl < REC_PLANES; pl++) bulk_g2s(&dst[pl][0], src + pl * pstride + first, bytes, bar);
}

// Can record (qa,qb,qc,qd) reach alpha >= 1/255 inside the (widened) block rectangle?
__device__ __forceinline__ bool entry_hits_block(float4 qa, float4 qb, float4 qc, float4 qd, float x0, float x1,
                                                 float y0, float y1) {
    if (__float_as_uint(qd.w) & REC_FLAG_ALWAYS) return true;
    const float tau = qd.y;
    if (disc_hits_rect(qa.w, qb.w, 0.5f * tau, x0, x1, y0, y1)) return true;
    const Quadric q = make_quadric(make_float3(qa.x, qa.y, qa.z), make_float3(qb.x, qb.y, qb.z),
                                   make_float3(qc.x, qc.y, qc.z), tau);
    const float det = q.xx * q.yy - q.xy * q.xy;
    if (!(q.xx > 0.f && q.yy > 0.f && det > 0.f)) return true;   // numerically not an ellipse in this frame
    return ellipse_hits_rect(q, x0, x1, y0, y1);
}

struct PairEval {
    bool valid, ray;
    float alpha, depth, G, s0, s1, ip, d0, d1;
};

// Per-(pixel, splat) evaluation up to alpha; (fx, fy) are tile-local pixel coordinates.
// Follows S/forward.cu:351-383 with p = a x + b y + c and depth = det(T)/p.z.
__device__ __forceinline__ PairEval eval_pair(float4 qa, float4 qb, float4 qc, float4 qd, float fx, float fy) {
    PairEval e;
    const float2 p01 = ffma2s(make_float2(qa.x, qa.y), fx, ffma2s(make_float2(qb.x, qb.y), fy, make_float2(qc.x, qc.y)));
    const float p2 = fmaf(qa.z, fx, fmaf(qb.z, fy, qc.z));
    e.ip = fast_rcp(p2);
    const float2 s01 = fmul2s(p01, e.ip);
    e.s0 = s01.x;
    e.s1 = s01.y;
    const float rho3d = e.s0 * e.s0 + e.s1 * e.s1;
    e.d0 = qa.w - fx;
    e.d1 = qb.w - fy;
    const float rho2d = FILTER_INV_SQUARE * (e.d0 * e.d0 + e.d1 * e.d1);
    const float rho = fminf(rho3d, rho2d);
    e.ray = rho3d <= rho2d;
    e.depth = e.ray ? qd.x * e.ip : qd.z;
    e.G = fast_ex2(rho * NEG_HALF_LOG2E);
    e.alpha = fminf(ALPHA_MAX, qc.w * e.G);
    e.valid = (p2 != 0.0f) && !(e.depth < NEAR_N) && !(e.alpha < ALPHA_MIN);
    return e;
}

}  // namespace gsr
