// audit.cu -- verification hook: how close did every blend decision of a forward pass come to its threshold?
//
// The blend loop of all three rasterizers takes hard decisions per (pixel, splat) pair (S/forward.cu:368-389,
// G/forward.cu:336-356): skip if alpha < 1/255, skip if depth < 0.2, stop if T (1 - alpha) < 1e-4, "median" while
// T > 0.5, ray-splat vs low-pass branch.  Two correct float32 implementations can disagree on a decision whose
// operand lies within rounding of its threshold, and such a flip moves ONE pixel by up to ~4e-3.  Parity tests must
// not hide real errors behind a blanket outlier trim, so this kernel re-walks the record stream of a finished forward
// with exactly the arithmetic of the render kernels (same eval_pair / ewa_eval, same update order) and reports, per
// pixel, the smallest RELATIVE distance of any decision to its threshold.  A test then accepts a deviating pixel only
// if one of its decisions was demonstrably marginal.  The kernel also checks that it reproduces the forward's
// final_T / last_contributor bit for bit (`mismatches` counts the pixels where it does not).
//
// One thread per pixel, whole list from global memory: speed is irrelevant here (debug entry point).
#include "common.cuh"
#include "async_copy.cuh"
#include "cull.cuh"
#include "render_common.cuh"
#include "ewa_common.cuh"

namespace gsr {

struct EwaEval {
    bool valid;
    float alpha, G, dx, dy;
};
__device__ EwaEval ewa_eval_audit(float4 p0, float4 p1, float fx, float fy, float& power_out) {
    // identical expression order to ewa_eval (ewa_render.cu)
    EwaEval e;
    e.dx = p0.x - fx;
    e.dy = p0.y - fy;
    // the operation sequence nvcc emits for the reference's  -0.5f * (a dx dx + c dy dy) - b dx dy  (G/forward.cu:339, read
    // off oracle/_ref/libref_gaussian.so): fma(dx, dx a, dy (dy c)) ; fma(that, -0.5, -(dy (dx b))).  Spelled out so that
    // `power` -- and with it the power > 0 test -- is bit-identical to the reference whatever code surrounds this function
    // (left to the compiler, a refactoring of the forward changed the contraction and moved an Octree-PGSR iteration's
    // agreement with the reference kernels from 4.8e-4 to 1.1e-3).  dx, dy are exact in both (tile-local vs global pixels).
    const float quad = __fmaf_rn(e.dx, __fmul_rn(e.dx, p0.z), __fmul_rn(e.dy, __fmul_rn(e.dy, p1.x)));
    const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(e.dy, __fmul_rn(e.dx, p0.w)));
    e.G = fast_ex2(power * 1.44269504088896340736f);
    e.alpha = fminf(ALPHA_MAX, p1.y * e.G);
    e.valid = !(power > 0.0f) && !(e.alpha < ALPHA_MIN);
    power_out = power;
    return e;
}

__device__ __forceinline__ float rel_gap(float v, float thr) { return fabsf(v / thr - 1.0f); }

// margins (5,H,W): alpha vs 1/255 | T(1-alpha) vs 1e-4 | T vs 0.5 | depth vs 0.2 | rho3d vs rho2d;  info (2,H,W): blended count, last gid
__global__ void __launch_bounds__(TILE_PIX)
surfel_audit_kernel(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ planes, size_t pstride, int W, int H, int gx,
                    const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib, uint32_t idx_mask,
                    float* __restrict__ margins, int* __restrict__ info, int* __restrict__ mismatches) {
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int lx = threadIdx.x % TILE, ly = threadIdx.x / TILE;
    const int px = tx * TILE + lx, py = ty * TILE + ly;
    if (px >= W || py >= H) return;
    const float fx = (float)lx, fy = (float)ly;
    const uint32_t range_x = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - range_x);
    const float4* src = planes + range_x;
    float T = 1.0f;
    float m_alpha = 1e30f, m_T = 1e30f, m_med = 1e30f, m_near = 1e30f, m_branch = 1e30f;
    int blended = 0, last_gid = -1;
    uint32_t last_contrib = 0;
    for (int i = 0; i < n; i++) {
        const float4 qa = __ldg(src + i), qb = __ldg(src + pstride + i), qc = __ldg(src + 2 * pstride + i),
                     qd = __ldg(src + 3 * pstride + i);
        const PairEval ev = eval_pair(qa, qb, qc, qd, fx, fy);
        const float p2 = fmaf(qa.z, fx, fmaf(qb.z, fy, qc.z));
        const bool okp = p2 != 0.0f, okd = !(ev.depth < NEAR_N), oka = !(ev.alpha < ALPHA_MIN);
        if (okp && okd) m_alpha = fminf(m_alpha, rel_gap(ev.alpha, ALPHA_MIN));
        if (okp && oka) m_near = fminf(m_near, rel_gap(ev.depth, NEAR_N));
        if (!ev.valid) continue;
        const float test_T = T * (1.0f - ev.alpha);
        m_T = fminf(m_T, rel_gap(test_T, T_EPS));
        if (test_T < T_EPS) break;
        m_med = fminf(m_med, rel_gap(T, 0.5f));
        const float rho3d = ev.s0 * ev.s0 + ev.s1 * ev.s1, rho2d = FILTER_INV_SQUARE * (ev.d0 * ev.d0 + ev.d1 * ev.d1);
        m_branch = fminf(m_branch, fabsf(rho3d - rho2d) / fmaxf(fmaxf(rho3d, rho2d), 1e-30f));
        T = test_T;
        blended++;
        last_contrib = (uint32_t)(i + 1);
        last_gid = (int)(__float_as_uint(qd.w) & idx_mask);
    }
    const size_t N = (size_t)W * H, pid = (size_t)py * W + px;
    margins[pid] = m_alpha; margins[pid + N] = m_T; margins[pid + 2 * N] = m_med; margins[pid + 3 * N] = m_near;
    margins[pid + 4 * N] = m_branch;
    info[pid] = blended; info[pid + N] = last_gid;
    if (__float_as_uint(T) != __float_as_uint(final_T[pid]) || last_contrib != n_contrib[pid]) atomicAdd(mismatches, 1);
}

// margins (3,H,W): alpha vs 1/255 | T(1-alpha) vs 1e-4 | T vs 0.5 (out_observe);  info (2,H,W): blended count, last gid
__global__ void __launch_bounds__(TILE_PIX)
ewa_audit_kernel(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ planes, size_t pstride, int W, int H, int gx,
                 const float* __restrict__ final_T, const uint32_t* __restrict__ n_contrib, uint32_t idx_mask,
                 float* __restrict__ margins, int* __restrict__ info, int* __restrict__ mismatches) {
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int lx = threadIdx.x % TILE, ly = threadIdx.x / TILE;
    const int px = tx * TILE + lx, py = ty * TILE + ly;
    if (px >= W || py >= H) return;
    const float fx = (float)lx, fy = (float)ly;
    const uint32_t range_x = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - range_x);
    const float4* src = planes + range_x;
    float T = 1.0f;
    float m_alpha = 1e30f, m_T = 1e30f, m_med = 1e30f;
    int blended = 0, last_gid = -1;
    uint32_t last_contrib = 0;
    for (int i = 0; i < n; i++) {
        const float4 p0 = __ldg(src + i), p1 = __ldg(src + pstride + i);
        float power;
        const EwaEval ev = ewa_eval_audit(p0, p1, fx, fy, power);
        if (!(power > 0.0f)) m_alpha = fminf(m_alpha, rel_gap(ev.alpha, ALPHA_MIN));
        if (!ev.valid) continue;
        const float test_T = T * (1.0f - ev.alpha);
        m_T = fminf(m_T, rel_gap(test_T, T_EPS));
        if (test_T < T_EPS) break;
        m_med = fminf(m_med, rel_gap(T, 0.5f));
        T = test_T;
        blended++;
        last_contrib = (uint32_t)(i + 1);
        last_gid = (int)(__float_as_uint(p1.w) & idx_mask);
    }
    const size_t N = (size_t)W * H, pid = (size_t)py * W + px;
    margins[pid] = m_alpha; margins[pid + N] = m_T; margins[pid + 2 * N] = m_med;
    info[pid] = blended; info[pid + N] = last_gid;
    if (__float_as_uint(T) != __float_as_uint(final_T[pid]) || last_contrib != n_contrib[pid]) atomicAdd(mismatches, 1);
}

cudaError_t launch_surfel_audit(int ntiles, const uint32_t* tile_offset, const float4* planes, size_t pstride, int W, int H, int gx,
                                const float* final_T, const uint32_t* n_contrib, uint32_t idx_mask, float* margins, int* info,
                                int* mismatches, cudaStream_t s) {
    surfel_audit_kernel<<<ntiles, TILE_PIX, 0, s>>>(tile_offset, planes, pstride, W, H, gx, final_T, n_contrib, idx_mask, margins, info,
                                                     mismatches);
    return cudaGetLastError();
}
cudaError_t launch_ewa_audit(int ntiles, const uint32_t* tile_offset, const float4* planes, size_t pstride, int W, int H, int gx,
                             const float* final_T, const uint32_t* n_contrib, uint32_t idx_mask, float* margins, int* info,
                             int* mismatches, cudaStream_t s) {
    ewa_audit_kernel<<<ntiles, TILE_PIX, 0, s>>>(tile_offset, planes, pstride, W, H, gx, final_T, n_contrib, idx_mask, margins, info,
                                                  mismatches);
    return cudaGetLastError();
}

}  // namespace gsr
