// ewa_preprocess.cu -- per-Gaussian forward / backward preprocess of the EWA-splat rasterizers
// (3DGS and PGSR plane) and scaffold_filter's visible_filter, for sm_100a.
//
// Semantics follow the reference kernels
//   forward : G/cuda_rasterizer/forward.cu:155-256 (+ computeCov3D :115-151, computeCov2D :74-112,
//             computeColorFromSH :20-71); L/ is identical; F/cuda_rasterizer/forward.cu:267-342 is the
//             same sequence stopping at the radius (visible_filter)
//   backward: G/cuda_rasterizer/backward.cu:144-274 (computeCov2DCUDA), :278-343 (computeCov3D),
//             :346-396 (preprocessCUDA)
// The data layout is ours: the forward emits one 32-byte EwaGeom per Gaussian plus the exact
// contribution ellipse used to cull (Gaussian, tile) pairs without changing any result; the backward
// is ONE kernel (the reference launches two and round-trips dL_dcov3D / dL_dmean3D through HBM), reads
// the reduced 64-byte accumulator and fully writes every output row, so no output needs a memset, and
// recomputes cov3D from scale/rotation instead of storing 24 bytes per Gaussian in the forward.
#include "common.cuh"
#include "sh.cuh"
#include "cull.cuh"
#include "pinned.cuh"
#include "ewa_common.cuh"
#include "../../include/gsr_b200.h"

namespace gsr {

// Rotation from the quaternion (w,x,y,z) AS GIVEN -- the reference does not normalise it
// (G/forward.cu:127).  Rs[r][c] is the usual row-major rotation matrix.
__device__ __forceinline__ void ewa_rotation(float4 q, float Rs[3][3]) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    Rs[0][0] = 1.f - 2.f * (y * y + z * z); Rs[0][1] = 2.f * (x * y - r * z); Rs[0][2] = 2.f * (x * z + r * y);
    Rs[1][0] = 2.f * (x * y + r * z); Rs[1][1] = 1.f - 2.f * (x * x + z * z); Rs[1][2] = 2.f * (y * z - r * x);
    Rs[2][0] = 2.f * (x * z - r * y); Rs[2][1] = 2.f * (y * z + r * x); Rs[2][2] = 1.f - 2.f * (x * x + y * y);
}

// Sigma = R S S R^T, six upper entries (xx xy xz yy yz zz), G/forward.cu:115-151.
// With Mk[i] = s_k Rs[i][k] (row k of the reference's M = S R^T): Sigma_ij = sum_k Mk[i] Mk[j].
__device__ __forceinline__ void ewa_cov3d(float3 s, float4 q, float* cov) {
    float Rs[3][3];
    ewa_rotation(q, Rs);
    const float sk[3] = {s.x, s.y, s.z};
    float Mk[3][3];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int i = 0; i < 3; i++) Mk[k][i] = sk[k] * Rs[i][k];
    int o = 0;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = i; j < 3; j++) cov[o++] = Mk[0][i] * Mk[0][j] + Mk[1][i] * Mk[1][j] + Mk[2][i] * Mk[2][j];
}

// Pieces of the EWA projection shared by the forward (computeCov2D, G/forward.cu:74-112) and the
// backward (G/backward.cu:160-206): clamped view-space mean t, the two rows A0, A1 of J W (the
// reference's T = W J, read as T[i][j] = Ai[j]), u_i = Vrk A_i, and the clamp gradient multipliers.
struct EwaProj {
    float3 t;
    float A0[3], A1[3], u0[3], u1[3];
    float gmx, gmy;
};
__device__ __forceinline__ EwaProj ewa_project(float3 pv, float fx, float fy, float tan_fovx, float tan_fovy,
                                               const float* __restrict__ view, const float* __restrict__ cov) {
    EwaProj e;
    const float limx = 1.3f * tan_fovx, limy = 1.3f * tan_fovy;
    const float txtz = pv.x / pv.z, tytz = pv.y / pv.z;
    e.t = make_float3(fminf(limx, fmaxf(-limx, txtz)) * pv.z, fminf(limy, fmaxf(-limy, tytz)) * pv.z, pv.z);
    e.gmx = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
    e.gmy = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
    const float j00 = fx / e.t.z, j02 = -(fx * e.t.x) / (e.t.z * e.t.z);
    const float j11 = fy / e.t.z, j12 = -(fy * e.t.y) / (e.t.z * e.t.z);
#pragma unroll
    for (int r = 0; r < 3; r++) {
        e.A0[r] = view[4 * r] * j00 + view[4 * r + 2] * j02;
        e.A1[r] = view[4 * r + 1] * j11 + view[4 * r + 2] * j12;
    }
    const float V[3][3] = {{cov[0], cov[1], cov[2]}, {cov[1], cov[3], cov[4]}, {cov[2], cov[4], cov[5]}};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        e.u0[k] = e.A0[0] * V[k][0] + e.A0[1] * V[k][1] + e.A0[2] * V[k][2];
        e.u1[k] = e.A1[0] * V[k][0] + e.A1[1] * V[k][1] + e.A1[2] * V[k][2];
    }
    return e;
}

// kRadiiOnly: scaffold_filter.visible_filter (F/forward.cu:267-342) -- stops at the radius.
template <bool kRadiiOnly>
__global__ void __launch_bounds__(256, 5)
ewa_preprocess_fwd(int P, int D, int M, const float* __restrict__ means3D, const float* __restrict__ scales,
                   const float4* __restrict__ rotations, const float* __restrict__ opacities,
                   const float* __restrict__ shs, const float* __restrict__ cov3D_precomp, const bool has_colors,
                   const ViewParams vc, const float focal_x, const float focal_y, const float tan_fovx,
                   const float tan_fovy, const bool prefiltered, const bool no_cull, int* __restrict__ radii,
                   EwaGeom* __restrict__ geom, CullRec* __restrict__ cull, float* __restrict__ depths,
                   uint32_t* __restrict__ masks, uint32_t* __restrict__ tile_count, float* __restrict__ rgb,
                   uint8_t* __restrict__ clamped, int* __restrict__ flags) {
    const int idx = kRadiiOnly ? (int)(blockIdx.x * blockDim.x + threadIdx.x) : spread_gaussian_index();
    __shared__ float4 s_rec[kRadiiOnly ? 1 : 8][4][32];      // warp_count_tiles (cull.cuh); unused by the radii-only filter
    __shared__ uint32_t s_mask[kRadiiOnly ? 1 : 8][32];
    __shared__ int s_prefix[kRadiiOnly ? 1 : 257];       // cta_count_big_tiles
    int radius_out = 0;
    float view[16];
    load16(vc.view, view);
    CullRec cr_t;
    cr_t.q0 = cr_t.q1 = cr_t.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
    float cx_t = 0.f, cy_t = 0.f;
    int x0_t = 0, y0_t = 0, w_t = 1, area_t = 0;
    if (idx < P) do {
        const float3 p = make_float3(__ldg(means3D + 3 * (size_t)idx), __ldg(means3D + 3 * (size_t)idx + 1),
                                     __ldg(means3D + 3 * (size_t)idx + 2));
        const float3 pv = xform43_pinned(view, p);
        if (pv.z <= 0.2f) {   // in_frustum, G/auxiliary.h:139-163
            if (prefiltered) atomicExch(flags, 1);
            break;
        }
        float cov3D[6];
        if (cov3D_precomp != nullptr) {
#pragma unroll
            for (int i = 0; i < 6; i++) cov3D[i] = __ldg(cov3D_precomp + 6 * (size_t)idx + i);
        } else {
            const float m = vc.scale_modifier;
            const float3 s = make_float3(m * __ldg(scales + 3 * (size_t)idx), m * __ldg(scales + 3 * (size_t)idx + 1),
                                         m * __ldg(scales + 3 * (size_t)idx + 2));
            ewa_cov3d(s, __ldg(rotations + idx), cov3D);
        }
        const EwaProj e = ewa_project(pv, focal_x, focal_y, tan_fovx, tan_fovy, view, cov3D);
        // cov = T^T Vrk^T T + 0.3 I; stored entries [0][0], [0][1] (= u1.A0), [1][1]
        const float cxx = e.u0[0] * e.A0[0] + e.u0[1] * e.A0[1] + e.u0[2] * e.A0[2] + 0.3f;
        const float cxy = e.u1[0] * e.A0[0] + e.u1[1] * e.A0[1] + e.u1[2] * e.A0[2];
        const float cyy = e.u1[0] * e.A1[0] + e.u1[1] * e.A1[1] + e.u1[2] * e.A1[2] + 0.3f;
        const float det = cxx * cyy - cxy * cxy;
        if (det == 0.0f) break;
        const float det_inv = 1.f / det;
        const float3 conic = make_float3(cyy * det_inv, -cxy * det_inv, cxx * det_inv);
        const float mid = 0.5f * (cxx + cyy);
        const float root = sqrtf(fmaxf(0.1f, mid * mid - det));
        const float radius = ceilf(3.f * sqrtf(fmaxf(mid + root, mid - root)));
        // screen position: p_hom, p_w = 1/(w + 1e-7), ndc2Pix in double (G/auxiliary.h:41-44)
        float pm[16];
        load16(vc.proj, pm);
        const float hx = __fadd_rn(dot_yxz(p.x, pm[0], p.y, pm[4], p.z, pm[8]), pm[12]);
        const float hy = __fadd_rn(dot_yxz(p.x, pm[1], p.y, pm[5], p.z, pm[9]), pm[13]);
        const float hw = __fadd_rn(dot_yxz(p.x, pm[3], p.y, pm[7], p.z, pm[11]), pm[15]);
        const float p_w = 1.0f / (hw + 0.0000001f);
        const float px = (float)((((double)(hx * p_w) + 1.0) * vc.W - 1.0) * 0.5);
        const float py = (float)((((double)(hy * p_w) + 1.0) * vc.H - 1.0) * 0.5);
        const int ri = (int)radius;
        // A covariance that overflowed (scales ~1e18 and beyond) has a NaN radius -> 0.  The reference then still counts
        // the one tile under the centre but writes no key for radii == 0 (G/rasterizer_impl.cu duplicateWithKeys): an
        // uninitialised slot in its lists.  Here such a Gaussian is simply invisible.
        if (ri <= 0) break;
        int x0, y0, x1, y1;
        get_rect(px, py, ri, vc.gx, vc.gy, x0, y0, x1, y1);
        if ((x1 - x0) * (y1 - y0) == 0) break;
        radius_out = ri;
        if (kRadiiOnly) break;

        // SH -> RGB is evaluated by sh_forward_kernel (sh.cu) on the Gaussians this kernel keeps (radii > 0)
        (void)has_colors; (void)shs; (void)rgb; (void)clamped; (void)D; (void)M;
        const float opa = __ldg(opacities + idx);
        // contribution ellipse {conic.x dx^2 + 2 conic.y dx dy + conic.z dy^2 <= tau}, tau = 2 ln(255 o):
        // alpha >= 1/255 only inside it (G/forward.cu:336-347)
        CullRec cr;
        float tau = contribution_tau(opa);
        int mode = CULL_EXACT;
        Quadric q = {conic.x, conic.y, conic.z, 0.f, 0.f, 0.f};
        if (no_cull) { mode = CULL_ALWAYS; tau = fmaxf(tau, 0.f); }
        else if (tau < 0.f) { mode = CULL_NEVER; tau = 0.f; }
        else if (tau == TAU_ALWAYS) mode = CULL_ALWAYS;
        else {
            const bool finite = fabsf(q.xx) < 3e38f && fabsf(q.yy) < 3e38f && fabsf(q.xy) < 3e38f &&
                                fabsf(px) < 1e7f && fabsf(py) < 1e7f;
            if (!finite || !quadric_is_ellipse(q)) mode = CULL_ALWAYS;
        }
        cr.q0 = make_float4(q.xx, q.xy, q.yy, 0.f);
        cr.q1 = make_float4(0.f, -tau, -1.0f /* no low-pass disc */, tau);
        cr.q2 = make_float4(px, py, (float)mode, 0.f);
        cull[idx] = cr;
        EwaGeom g;
        g.a = make_float4(px, py, conic.x, conic.y);
        g.b = make_float4(conic.z, opa, tau, (float)mode);
        geom[idx] = g;
        depths[idx] = pv.z;
        cr_t = cr; cx_t = px; cy_t = py;
        x0_t = x0; y0_t = y0; w_t = x1 - x0; area_t = w_t * (y1 - y0);
    } while (0);
    if (kRadiiOnly) {
        if (idx < P) radii[idx] = radius_out;
        return;
    }
    // tiles of the reference rect (G/auxiliary.h:46-56) the splat can actually reach, counted by the whole warp over
    // the flattened (Gaussian, tile) list (cull.cuh)
    const int wic = threadIdx.x >> 5;
    const bool big = area_t > WARP_AREA_MAX;            // screen-filling rectangles: flattened over the CTA instead
    const uint32_t m = warp_count_tiles(cr_t, cx_t, cy_t, x0_t, y0_t, w_t, big ? 0 : area_t, vc.gx, tile_count,
                                        s_rec[kRadiiOnly ? 0 : wic], s_mask[kRadiiOnly ? 0 : wic]);
    if (!kRadiiOnly) cta_count_big_tiles(big ? area_t : 0, vc.gx, tile_count, s_rec, s_prefix);
    if (idx < P) {
        radii[idx] = radius_out;
        masks[idx] = area_t == 0 ? 0u : (area_t <= 32 ? m : MASK_RETEST);
    }
}

template __global__ void ewa_preprocess_fwd<false>(int, int, int, const float*, const float*, const float4*, const float*,
                                                   const float*, const float*, const bool, const ViewParams, const float,
                                                   const float, const float, const float, const bool, const bool, int*,
                                                   EwaGeom*, CullRec*, float*, uint32_t*, uint32_t*, float*, uint8_t*, int*);
template __global__ void ewa_preprocess_fwd<true>(int, int, int, const float*, const float*, const float4*, const float*,
                                                  const float*, const float*, const bool, const ViewParams, const float,
                                                  const float, const float, const float, const bool, const bool, int*,
                                                  EwaGeom*, CullRec*, float*, uint32_t*, uint32_t*, float*, uint8_t*, int*);

// One thread per Gaussian; every output row is written (zeros when radii == 0).
__global__ void __launch_bounds__(256)
ewa_preprocess_bwd(int P, int D, int M, const float* __restrict__ means3D, const float* __restrict__ scales,
                   const float4* __restrict__ rotations, const float* __restrict__ shs,
                   const float* __restrict__ cov3D_precomp, const ViewParams vc, const float focal_x,
                   const float focal_y, const float tan_fovx, const float tan_fovy, const int* __restrict__ radii,
                   const uint8_t* __restrict__ clamped, const float* __restrict__ gacc,
                   float* __restrict__ dL_dmean2D, float* __restrict__ dL_dmean2D_abs, float* __restrict__ dL_dconic,
                   float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor, float* __restrict__ dL_dmean3D,
                   float* __restrict__ dL_dcov3D, float* __restrict__ dL_dsh, float* __restrict__ dL_dscale,
                   float* __restrict__ dL_drot, float* __restrict__ dL_dall_map) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float4* ga = reinterpret_cast<const float4*>(gacc + (size_t)idx * EWA_GACC);
    const float4 a0 = ga[0], a1 = ga[1], a2 = ga[2], a3 = ga[3];
    const float dm2x = a0.x, dm2y = a0.y;
    const float dcon_x = a0.z, dcon_y = a0.w, dcon_w = a1.x;
    const float dopa = a1.y;
    const float3 dcol = make_float3(a1.z, a1.w, a2.x);
    float3 dmean = make_float3(0.f, 0.f, 0.f);
    float dcov[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float3 dscale = make_float3(0.f, 0.f, 0.f);
    float4 drot = make_float4(0.f, 0.f, 0.f, 0.f);
    (void)dL_dsh; (void)shs; (void)clamped; (void)D; (void)M;   // SH backward lives in sh.cu
    if (radii[idx] > 0) {
        float view[16];
        load16(vc.view, view);
        const float3 p = make_float3(__ldg(means3D + 3 * (size_t)idx), __ldg(means3D + 3 * (size_t)idx + 1),
                                     __ldg(means3D + 3 * (size_t)idx + 2));
        float cov3D[6];
        float3 s = make_float3(0.f, 0.f, 0.f);
        float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
        if (cov3D_precomp != nullptr) {
#pragma unroll
            for (int i = 0; i < 6; i++) cov3D[i] = __ldg(cov3D_precomp + 6 * (size_t)idx + i);
        } else {
            const float m = vc.scale_modifier;
            s = make_float3(m * __ldg(scales + 3 * (size_t)idx), m * __ldg(scales + 3 * (size_t)idx + 1),
                            m * __ldg(scales + 3 * (size_t)idx + 2));
            q = __ldg(rotations + idx);
            ewa_cov3d(s, q, cov3D);
        }
        // ---- conic -> cov2D -> cov3D and view-space mean (G/backward.cu:144-274)
        const float3 pv = xform43_pinned(view, p);
        const EwaProj e = ewa_project(pv, focal_x, focal_y, tan_fovx, tan_fovy, view, cov3D);
        const float a = e.u0[0] * e.A0[0] + e.u0[1] * e.A0[1] + e.u0[2] * e.A0[2] + 0.3f;
        const float b = e.u1[0] * e.A0[0] + e.u1[1] * e.A0[1] + e.u1[2] * e.A0[2];
        const float c = e.u1[0] * e.A1[0] + e.u1[1] * e.A1[1] + e.u1[2] * e.A1[2] + 0.3f;
        const float denom = a * c - b * b;
        const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float dL_da = 0.f, dL_db = 0.f, dL_dc = 0.f;
        if (denom2inv != 0.f) {
            dL_da = denom2inv * (-c * c * dcon_x + 2 * b * c * dcon_y + (denom - a * c) * dcon_w);
            dL_dc = denom2inv * (-a * a * dcon_w + 2 * a * b * dcon_y + (denom - a * c) * dcon_x);
            dL_db = denom2inv * 2 * (b * c * dcon_x - (denom + 2 * b * b) * dcon_y + a * b * dcon_w);
            const float* A0 = e.A0;
            const float* A1 = e.A1;
            dcov[0] = A0[0] * A0[0] * dL_da + A0[0] * A1[0] * dL_db + A1[0] * A1[0] * dL_dc;
            dcov[3] = A0[1] * A0[1] * dL_da + A0[1] * A1[1] * dL_db + A1[1] * A1[1] * dL_dc;
            dcov[5] = A0[2] * A0[2] * dL_da + A0[2] * A1[2] * dL_db + A1[2] * A1[2] * dL_dc;
            dcov[1] = 2 * A0[0] * A0[1] * dL_da + (A0[0] * A1[1] + A0[1] * A1[0]) * dL_db + 2 * A1[0] * A1[1] * dL_dc;
            dcov[2] = 2 * A0[0] * A0[2] * dL_da + (A0[0] * A1[2] + A0[2] * A1[0]) * dL_db + 2 * A1[0] * A1[2] * dL_dc;
            dcov[4] = 2 * A0[2] * A0[1] * dL_da + (A0[1] * A1[2] + A0[2] * A1[1]) * dL_db + 2 * A1[1] * A1[2] * dL_dc;
        }
        float dT0[3], dT1[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dT0[k] = 2 * e.u0[k] * dL_da + e.u1[k] * dL_db;
            dT1[k] = 2 * e.u1[k] * dL_dc + e.u0[k] * dL_db;
        }
        const float dJ00 = view[0] * dT0[0] + view[4] * dT0[1] + view[8] * dT0[2];
        const float dJ02 = view[2] * dT0[0] + view[6] * dT0[1] + view[10] * dT0[2];
        const float dJ11 = view[1] * dT1[0] + view[5] * dT1[1] + view[9] * dT1[2];
        const float dJ12 = view[2] * dT1[0] + view[6] * dT1[1] + view[10] * dT1[2];
        const float tz = 1.f / e.t.z, tz2 = tz * tz, tz3 = tz2 * tz;
        const float3 dt = make_float3(
            e.gmx * -focal_x * tz2 * dJ02, e.gmy * -focal_y * tz2 * dJ12,
            -focal_x * tz2 * dJ00 - focal_y * tz2 * dJ11 + (2 * focal_x * e.t.x) * tz3 * dJ02 + (2 * focal_y * e.t.y) * tz3 * dJ12);
        dmean = xformvec43T(view, dt);
        // ---- screen-space mean -> 3D mean through the projection (G/backward.cu:365-383)
        float pm[16];
        load16(vc.proj, pm);
        const float m_w = 1.0f / ((pm[3] * p.x + pm[7] * p.y + pm[11] * p.z + pm[15]) + 0.0000001f);
        const float mul1 = (pm[0] * p.x + pm[4] * p.y + pm[8] * p.z + pm[12]) * m_w * m_w;
        const float mul2 = (pm[1] * p.x + pm[5] * p.y + pm[9] * p.z + pm[13]) * m_w * m_w;
        dmean.x += (pm[0] * m_w - pm[3] * mul1) * dm2x + (pm[1] * m_w - pm[3] * mul2) * dm2y;
        dmean.y += (pm[4] * m_w - pm[7] * mul1) * dm2x + (pm[5] * m_w - pm[7] * mul2) * dm2y;
        dmean.z += (pm[8] * m_w - pm[11] * mul1) * dm2x + (pm[9] * m_w - pm[11] * mul2) * dm2y;
        // SH backward: sh_backward_kernel (sh.cu) runs after this kernel and accumulates into dL_dmean3D
        // ---- cov3D -> scale, quaternion (G/backward.cu:278-343): M = S R^T, dL/dM = 2 M dSigma
        if (cov3D_precomp == nullptr) {
            float Rs[3][3];
            ewa_rotation(q, Rs);
            const float sk[3] = {s.x, s.y, s.z};
            const float Dm[3][3] = {{dcov[0], 0.5f * dcov[1], 0.5f * dcov[2]},
                                    {0.5f * dcov[1], dcov[3], 0.5f * dcov[4]},
                                    {0.5f * dcov[2], 0.5f * dcov[4], dcov[5]}};
            float dM[3][3];   // dM[k][r] = dL/dM(row k, col r)
            float ds[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {
#pragma unroll
                for (int r = 0; r < 3; r++)
                    dM[k][r] = 2.0f * (sk[k] * Rs[0][k] * Dm[0][r] + sk[k] * Rs[1][k] * Dm[1][r] + sk[k] * Rs[2][k] * Dm[2][r]);
                ds[k] = Rs[0][k] * dM[k][0] + Rs[1][k] * dM[k][1] + Rs[2][k] * dM[k][2];
#pragma unroll
                for (int r = 0; r < 3; r++) dM[k][r] *= sk[k];
            }
            dscale = make_float3(ds[0], ds[1], ds[2]);
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            drot.x = 2 * z * (dM[0][1] - dM[1][0]) + 2 * y * (dM[2][0] - dM[0][2]) + 2 * x * (dM[1][2] - dM[2][1]);
            drot.y = 2 * y * (dM[1][0] + dM[0][1]) + 2 * z * (dM[2][0] + dM[0][2]) + 2 * r * (dM[1][2] - dM[2][1]) - 4 * x * (dM[2][2] + dM[1][1]);
            drot.z = 2 * x * (dM[1][0] + dM[0][1]) + 2 * r * (dM[2][0] - dM[0][2]) + 2 * z * (dM[1][2] + dM[2][1]) - 4 * y * (dM[2][2] + dM[0][0]);
            drot.w = 2 * r * (dM[0][1] - dM[1][0]) + 2 * x * (dM[2][0] + dM[0][2]) + 2 * y * (dM[1][2] + dM[2][1]) - 4 * z * (dM[1][1] + dM[0][0]);
        }
    }
    dL_dmean2D[3 * (size_t)idx + 0] = dm2x; dL_dmean2D[3 * (size_t)idx + 1] = dm2y; dL_dmean2D[3 * (size_t)idx + 2] = 0.f;
    if (dL_dmean2D_abs) {
        dL_dmean2D_abs[3 * (size_t)idx + 0] = a2.y; dL_dmean2D_abs[3 * (size_t)idx + 1] = a2.z;
        dL_dmean2D_abs[3 * (size_t)idx + 2] = 0.f;
    }
    if (dL_dconic) reinterpret_cast<float4*>(dL_dconic)[idx] = make_float4(dcon_x, dcon_y, 0.f, dcon_w);
    dL_dopacity[idx] = dopa;
    dL_dcolor[3 * (size_t)idx + 0] = dcol.x; dL_dcolor[3 * (size_t)idx + 1] = dcol.y; dL_dcolor[3 * (size_t)idx + 2] = dcol.z;
    dL_dmean3D[3 * (size_t)idx + 0] = dmean.x; dL_dmean3D[3 * (size_t)idx + 1] = dmean.y; dL_dmean3D[3 * (size_t)idx + 2] = dmean.z;
#pragma unroll
    for (int i = 0; i < 6; i++) dL_dcov3D[6 * (size_t)idx + i] = dcov[i];
    if (dL_dscale) { dL_dscale[3 * (size_t)idx] = dscale.x; dL_dscale[3 * (size_t)idx + 1] = dscale.y; dL_dscale[3 * (size_t)idx + 2] = dscale.z; }
    if (dL_drot) reinterpret_cast<float4*>(dL_drot)[idx] = drot;
    if (dL_dall_map) {
        float* o = dL_dall_map + NUM_ALL_MAP * (size_t)idx;
        o[0] = a2.w; o[1] = a3.x; o[2] = a3.y; o[3] = a3.z; o[4] = a3.w;
    }
}

}  // namespace gsr
