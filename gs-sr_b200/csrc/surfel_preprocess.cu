// surfel_preprocess.cu -- per-Gaussian forward / backward preprocess of the 2DGS
// surfel rasterizer for sm_100a.
//
// Semantics follow the reference kernels
//   forward : S/cuda_rasterizer/forward.cu:148-251 (+ compute_transmat :75-115,
//             compute_aabb :119-145, computeColorFromSH :20-71)
//   backward: S/cuda_rasterizer/backward.cu:582-637 (+ compute_transmat_aabb :450-580,
//             computeColorFromSH :20-139)
// but the data layout is ours: the forward emits one 64-byte GeomRec per
// Gaussian (four 128-bit stores) plus a conservative contribution box used to
// cull (pixel, splat) pairs without changing any result; the backward reads
// the reduced 80-byte gradient accumulator and fully writes every output
// (zeros for culled Gaussians), so no output needs a memset.
#include "common.cuh"
#include "sh.cuh"
#include "cull.cuh"
#include "pinned.cuh"
#include "../../include/gsr_b200.h"

namespace gsr {

// quat (w,x,y,z) -> rotation columns, normalised in-kernel (S/auxiliary.h:215-237)
__device__ __forceinline__ void quat_to_rot(float4 q, float3& c0, float3& c1, float3& c2) {
    float s = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    float w = q.x * s, x = q.y * s, y = q.z * s, z = q.w * s;
    c0 = make_float3(1.f - 2.f * (y * y + z * z), 2.f * (x * y + w * z), 2.f * (x * z - w * y));
    c1 = make_float3(2.f * (x * y - w * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z + w * x));
    c2 = make_float3(2.f * (x * z + w * y), 2.f * (y * z - w * x), 1.f - 2.f * (x * x + y * y));
}

// quat (w,x,y,z) -> rotation columns, S/auxiliary.h:215-237, in the reference build's rounding
__device__ __forceinline__ void quat_to_rot_pinned(float4 q, float3& c0, float3& c1, float3& c2) {
    const float sum = __fmaf_rn(q.z, q.z, __fmaf_rn(q.y, q.y, __fmaf_rn(q.x, q.x, __fmul_rn(q.w, q.w))));
    const float s = rsqrtf(sum);
    const float w = __fmul_rn(q.x, s), x = __fmul_rn(q.y, s), y = __fmul_rn(q.z, s), z = __fmul_rn(q.w, s);
    const float wz = __fmul_rn(w, z), wx = __fmul_rn(w, x), wy = __fmul_rn(w, y);
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z);
    const float t00 = __fadd_rn(yy, zz), t11 = __fmaf_rn(x, x, zz), t22 = __fmaf_rn(x, x, yy);
    const float xy_p = __fmaf_rn(x, y, wz), xy_m = __fmaf_rn(x, y, -wz);
    const float yz_p = __fmaf_rn(y, z, wx), yz_m = __fmaf_rn(y, z, -wx);
    const float xz_p = __fmaf_rn(x, z, wy), xz_m = __fmaf_rn(x, z, -wy);
    c0 = make_float3(__fadd_rn(1.f, -__fadd_rn(t00, t00)), __fadd_rn(xy_p, xy_p), __fadd_rn(xz_m, xz_m));
    c1 = make_float3(__fadd_rn(xy_m, xy_m), __fadd_rn(1.f, -__fadd_rn(t11, t11)), __fadd_rn(yz_p, yz_p));
    c2 = make_float3(__fadd_rn(xz_p, xz_p), __fadd_rn(yz_m, yz_m), __fadd_rn(1.f, -__fadd_rn(t22, t22)));
}
// one row of compute_aabb: centre c = f.(T o Tw), h^2 = c^2 - f.(T o T)   (f = (9,9,-1)/d)
__device__ __forceinline__ void aabb_row_pinned(float3 T, float3 Tw, float f0, float invd, float& c, float& h2) {
    const float mx = __fmul_rn(Tw.x, T.x), my = __fmul_rn(Tw.y, T.y), mz = __fmul_rn(Tw.z, T.z);
    c = __fmaf_rn(mz, -invd, __fmaf_rn(f0, mx, __fmul_rn(f0, my)));
    const float qx = __fmul_rn(T.x, T.x), qy = __fmul_rn(T.y, T.y), qz = __fmul_rn(T.z, T.z);
    const float nd = __fmaf_rn(qz, invd, -__fmaf_rn(f0, qx, __fmul_rn(f0, qy)));
    h2 = __fmaf_rn(c, c, nd);
}

#ifndef GSR_PREFWD_MINB
#define GSR_PREFWD_MINB 5      // 48 registers: 155 -> 149 us at cfg-B
#endif
__global__ void __launch_bounds__(256, GSR_PREFWD_MINB)
surfel_preprocess_fwd(int P, int D, int M, const float* __restrict__ means3D,
                      const float2* __restrict__ scales, const float4* __restrict__ rotations,
                      const float* __restrict__ opacities, const float* __restrict__ shs,
                      const float* __restrict__ transMat_precomp, const bool has_colors,
                      const ViewParams vc, const bool prefiltered, const bool no_cull,
                      int* __restrict__ radii, GeomRec* __restrict__ geom, CullRec* __restrict__ cull, float* __restrict__ depths,
                      uint32_t* __restrict__ masks, uint32_t* __restrict__ tile_count, float* __restrict__ rgb,
                      uint8_t* __restrict__ clamped, int* __restrict__ flags) {
    __shared__ float4 s_rec[8][4][32];        // warp_count_tiles: the CullRec + rectangle of each lane, per warp
    __shared__ uint32_t s_mask[8][32];
    __shared__ int s_prefix[257];             // cta_count_big_tiles
    const int idx = spread_gaussian_index();
    int radius_out = 0;
    float view[16];
    load16(vc.view, view);
    // what the warp-cooperative tile count needs of this thread's Gaussian (area 0: culled or out of range)
    CullRec cr_t;
    cr_t.q0 = cr_t.q1 = cr_t.q2 = make_float4(0.f, 0.f, 0.f, 0.f);
    float cx_t = 0.f, cy_t = 0.f;
    int x0_t = 0, y0_t = 0, w_t = 1, area_t = 0;
    if (idx < P) do {
        float3 p = make_float3(__ldg(means3D + 3 * idx), __ldg(means3D + 3 * idx + 1), __ldg(means3D + 3 * idx + 2));
        float3 pv = xform43_pinned(view, p);
        if (pv.z <= 0.2f) {  // in_frustum, S/auxiliary.h:187-212
            if (prefiltered) atomicExch(flags, 1);
            break;
        }
        float3 Tu, Tv, Tw, normal;
        if (transMat_precomp == nullptr) {
            float3 c0, c1, c2;
            quat_to_rot_pinned(__ldg(rotations + idx), c0, c1, c2);
            float2 sc = __ldg(scales + idx);
            const float su = __fmul_rn(sc.x, vc.scale_modifier), sv = __fmul_rn(sc.y, vc.scale_modifier);
            const float3 L0 = make_float3(__fmul_rn(su, c0.x), __fmul_rn(su, c0.y), __fmul_rn(su, c0.z));
            const float3 L1 = make_float3(__fmul_rn(sv, c1.x), __fmul_rn(sv, c1.y), __fmul_rn(sv, c1.z));
            // T = (splat2world^T * world2ndc) * ndc2pix, S/forward.cu:93-112
            float pm[16];
            load16(vc.proj, pm);
            const float hw = 0.5f * (float)vc.W, hwm = 0.5f * (float)(vc.W - 1);
            const float hh = 0.5f * (float)vc.H, hhm = 0.5f * (float)(vc.H - 1);
            float cu[4], cv[4], cc[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                cu[c] = dot_yxz(L0.x, pm[c], L0.y, pm[4 + c], L0.z, pm[8 + c]);
                cv[c] = dot_yxz(L1.x, pm[c], L1.y, pm[4 + c], L1.z, pm[8 + c]);
                cc[c] = __fadd_rn(pm[12 + c], dot_yxz(p.x, pm[c], p.y, pm[4 + c], p.z, pm[8 + c]));
            }
            Tu = make_float3(__fmaf_rn(hwm, cu[3], __fmul_rn(hw, cu[0])), __fmaf_rn(hwm, cv[3], __fmul_rn(hw, cv[0])),
                             __fmaf_rn(hwm, cc[3], __fmul_rn(hw, cc[0])));
            Tv = make_float3(__fmaf_rn(hhm, cu[3], __fmul_rn(hh, cu[1])), __fmaf_rn(hhm, cv[3], __fmul_rn(hh, cv[1])),
                             __fmaf_rn(hhm, cc[3], __fmul_rn(hh, cc[1])));
            Tw = make_float3(cu[3], cv[3], cc[3]);
            normal = xformvec43_pinned(view, c2);
        } else {
            const float* t = transMat_precomp + 9 * (size_t)idx;
            Tu = make_float3(__ldg(t + 0), __ldg(t + 1), __ldg(t + 2));
            Tv = make_float3(__ldg(t + 3), __ldg(t + 4), __ldg(t + 5));
            Tw = make_float3(__ldg(t + 6), __ldg(t + 7), __ldg(t + 8));
            normal = make_float3(0.f, 0.f, 1.f);
        }
        // Domain limit of the record form: the per-tile records hold the adjugate rows and det T of the tile-local T
        // (~|T|^2, ~1e4 |T|^3), which must stay finite.  |T| ~ scale * focal + depth * pixel: ~1e5 in real scenes; beyond
        // 1e10 (world scales ~1e7, or NaN / Inf parameters, which the reference drops at its rectangle test) the surfel
        // is invisible here -- the reference still renders finite ones up to ~1e16.  Contained: no NaN reaches a pixel or
        // another Gaussian's gradient.
        const float tsum = fabsf(Tu.x) + fabsf(Tu.y) + fabsf(Tu.z) + fabsf(Tv.x) + fabsf(Tv.y) + fabsf(Tv.z) + fabsf(Tw.x) +
                           fabsf(Tw.y) + fabsf(Tw.z);
        if (!(tsum < 1e10f)) break;
        // DUAL_VISIABLE, S/forward.cu:209-214
        const float cosv = -dot_yxz(pv.x, normal.x, pv.y, normal.y, pv.z, normal.z);
        if (cosv == 0.f) break;
        const float mult = cosv > 0.f ? 1.f : -1.f;
        normal = make_float3(mult * normal.x, mult * normal.y, mult * normal.z);

        // compute_aabb with cutoff 3, S/forward.cu:119-145,223-231 (pinned sequences above)
        const float d = __fmaf_rn(-Tw.z, Tw.z, __fmaf_rn(__fmul_rn(Tw.x, Tw.x), 9.0f, __fmul_rn(__fmul_rn(Tw.y, Tw.y), 9.0f)));
        if (d == 0.0f) break;
        const float invd = __fdiv_rn(1.0f, d);
        const float f0 = __fmul_rn(invd, 9.0f);
        float cx, cy, h0, h1;
        aabb_row_pinned(Tu, Tw, f0, invd, cx, h0);
        aabb_row_pinned(Tv, Tw, f0, invd, cy, h1);
        const float ex = sqrtf(fmaxf(1e-4f, h0)), ey = sqrtf(fmaxf(1e-4f, h1));
        float radius = ceilf(fmaxf(fmaxf(ex, ey), 3.0f * FILTER_SIZE));
        int ri = (int)radius;
        int x0, y0, x1, y1;
        get_rect(cx, cy, ri, vc.gx, vc.gy, x0, y0, x1, y1);
        if ((x1 - x0) * (y1 - y0) == 0) break;

        // SH -> RGB is evaluated by sh_forward_kernel (sh.cu) on the Gaussians this kernel keeps (radii > 0)
        (void)has_colors; (void)shs; (void)rgb; (void)clamped; (void)D; (void)M;
        float opa = __ldg(opacities + idx);
        GeomRec g;
        g.tu = make_float4(Tu.x, Tu.y, Tu.z, cx);
        g.tv = make_float4(Tv.x, Tv.y, Tv.z, cy);
        g.tw = make_float4(Tw.x, Tw.y, Tw.z, opa);
        const CullRec cr = make_cull_rec(Tu, Tv, Tw, cx, cy, opa, no_cull);
        cull[idx] = cr;
        g.nd = make_float4(normal.x, normal.y, normal.z, ((int)cr.q2.z == CULL_EXACT) ? cr.q1.w : -(cr.q1.w + 1.f));
        geom[idx] = g;
        depths[idx] = pv.z;
        radius_out = ri;
        cr_t = cr; cx_t = cx; cy_t = cy;
        x0_t = x0; y0_t = y0; w_t = x1 - x0; area_t = w_t * (y1 - y0);
    } while (0);
    // tiles of the reference rect (S/auxiliary.h:69-79) that the splat can actually reach: counted per tile (bucket
    // sizes) and remembered as a bit mask -- by the whole warp over the flattened (Gaussian, tile) list (cull.cuh)
    const int wic = threadIdx.x >> 5;
    const bool big = area_t > WARP_AREA_MAX;            // screen-filling rectangles: flattened over the CTA instead
    const uint32_t m = warp_count_tiles(cr_t, cx_t, cy_t, x0_t, y0_t, w_t, big ? 0 : area_t, vc.gx, tile_count, s_rec[wic],
                                        s_mask[wic]);
    cta_count_big_tiles(big ? area_t : 0, vc.gx, tile_count, s_rec, s_prefix);
    if (idx < P) {
        radii[idx] = radius_out;
        masks[idx] = area_t == 0 ? 0u : (area_t <= 32 ? m : MASK_RETEST);
    }
}

// quat_to_rotmat_vjp, S/auxiliary.h:240-284. vR columns c0,c1,c2 (column-major).
__device__ __forceinline__ float4 quat_vjp(float4 q, float3 v0, float3 v1, float3 v2) {
    float s = rsqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    float w = q.x * s, x = q.y * s, y = q.z * s, z = q.w * s;
    float4 r;
    r.x = 2.f * (x * (v1.z - v2.y) + y * (v2.x - v0.z) + z * (v0.y - v1.x));
    r.y = 2.f * (-2.f * x * (v1.y + v2.z) + y * (v0.y + v1.x) + z * (v0.z + v2.x) + w * (v1.z - v2.y));
    r.z = 2.f * (x * (v0.y + v1.x) - 2.f * y * (v0.x + v2.z) + z * (v1.z + v2.y) + w * (v2.x - v0.z));
    r.w = 2.f * (x * (v0.z + v2.x) + y * (v1.z + v2.y) - 2.f * z * (v0.x + v1.y) + w * (v0.y - v1.x));
    return r;
}

// One thread per Gaussian; every output row is written (zeros when radii == 0).
#ifndef GSR_PREBWD_MINB
#define GSR_PREBWD_MINB 2      // with every load issued up front: 2 CTAs/SM (no spills) 117 us, 3 (108 B of spills) 120 us at cfg-B
#endif
__global__ void __launch_bounds__(256, GSR_PREBWD_MINB)
surfel_preprocess_bwd(int P, int D, int M, const float* __restrict__ means3D,
                      const float2* __restrict__ scales, const float4* __restrict__ rotations,
                      const float* __restrict__ shs, const bool precomp, const ViewParams vc,
                      const int Wb, const int Hb, const int* __restrict__ radii,
                      const GeomRec* __restrict__ geom, const uint8_t* __restrict__ clamped,
                      const float* __restrict__ gacc, float* __restrict__ dL_dmean2D,
                      float* __restrict__ dL_dnormal, float* __restrict__ dL_dopacity,
                      float* __restrict__ dL_dcolor, float* __restrict__ dL_dmean3D,
                      float* __restrict__ dL_dtransMat, float* __restrict__ dL_dsh,
                      float* __restrict__ dL_dscale, float* __restrict__ dL_drot) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    const float4* ga = reinterpret_cast<const float4*>(gacc + (size_t)idx * GACC_STRIDE);
    const float4 a0 = ga[0], a1 = ga[1], a2 = ga[2], a3 = ga[3], a4 = ga[4];
    const float3 M0 = make_float3(a0.x, a0.y, a0.z), MX = make_float3(a0.w, a1.x, a1.y),
                 MY = make_float3(a1.z, a1.w, a2.x);
    const float dDet = a2.y;
    float3 dcol = make_float3(a2.z, a2.w, a3.x);
    float3 dnrm = make_float3(a3.y, a3.z, a3.w);
    float dopa = a4.x;
    const float dTwz_lowpass = a4.y;
    float dm2x = a4.z, dm2y = a4.w;
    float dT[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float3 dmean = make_float3(0.f, 0.f, 0.f);
    float2 dscale = make_float2(0.f, 0.f);
    float4 drot = make_float4(0.f, 0.f, 0.f, 0.f);
    float m2x = dm2x, m2y = dm2y;
    float dTout[9];
#pragma unroll
    for (int i = 0; i < 9; i++) dTout[i] = dT[i];
    const bool visible = radii[idx] > 0;
    // every per-Gaussian load of the kernel is issued here, before the first use and regardless of visibility (rows of
    // culled Gaussians are in bounds and ignored): one memory round trip instead of three dependent ones
    const GeomRec g = geom[idx];
    float3 p = make_float3(0.f, 0.f, 0.f);
    float2 sc = make_float2(0.f, 0.f);
    float4 q = make_float4(1.f, 0.f, 0.f, 0.f);
    if (!precomp) {
        p = make_float3(__ldg(means3D + 3 * idx), __ldg(means3D + 3 * idx + 1), __ldg(means3D + 3 * idx + 2));
        q = __ldg(rotations + idx);
        sc = __ldg(scales + idx);
    }
    float view[16];
    load16(vc.view, view);
    // dL/dsh and the view-direction term of dL/dmean3D are produced by sh_backward_kernel (sh.cu), which runs
    // after this kernel and accumulates into dL_dmean3D
    (void)dL_dsh; (void)shs; (void)clamped; (void)D; (void)M;
    if (visible) {
        {   // moments of dL/dp -> dL/dT (the linear part of S/backward.cu:413-421, once per Gaussian)
            const float sx = moment_origin(g.tu.w, vc.W), sy = moment_origin(g.tv.w, vc.H);
            const float3 tw = make_float3(g.tw.x, g.tw.y, g.tw.z);
            const float3 tu = make_float3(fmaf(-sx, tw.x, g.tu.x), fmaf(-sx, tw.y, g.tu.y), fmaf(-sx, tw.z, g.tu.z));
            const float3 tv = make_float3(fmaf(-sy, tw.x, g.tv.x), fmaf(-sy, tw.y, g.tv.y), fmaf(-sy, tw.z, g.tv.z));
            const float3 A = cross3(tv, tw), B = cross3(tw, tu), C = cross3(tu, tv);   // cofactors = d det / dT
            const float3 u1 = cross3(MY, tw), u2 = cross3(tv, M0);
            const float3 v1 = cross3(tw, MX), v2 = cross3(M0, tu);
            const float3 w1 = cross3(MX, tv), w2 = cross3(tu, MY);
            const float3 dtu = make_float3(u1.x + u2.x + dDet * A.x, u1.y + u2.y + dDet * A.y, u1.z + u2.z + dDet * A.z);
            const float3 dtv = make_float3(v1.x + v2.x + dDet * B.x, v1.y + v2.y + dDet * B.y, v1.z + v2.z + dDet * B.z);
            const float3 dtw = make_float3(w1.x + w2.x + dDet * C.x, w1.y + w2.y + dDet * C.y,
                                           w1.z + w2.z + dDet * C.z + dTwz_lowpass);
            // Tu~ = Tu - sx Tw, Tv~ = Tv - sy Tw  =>  dTw = dTw~ - sx dTu~ - sy dTv~
            dT[0] = dtu.x; dT[1] = dtu.y; dT[2] = dtu.z;
            dT[3] = dtv.x; dT[4] = dtv.y; dT[5] = dtv.z;
            dT[6] = dtw.x - sx * dtu.x - sy * dtv.x;
            dT[7] = dtw.y - sx * dtu.y - sy * dtv.y;
            dT[8] = dtw.z - sx * dtu.z - sy * dtv.z;
#pragma unroll
            for (int i = 0; i < 9; i++) dTout[i] = dT[i];
        }
        float3 Tu, Tv, Tw, c0, c1, c2, normal = make_float3(0.f, 0.f, 0.f);
        float Pm[3][4];
        if (precomp) {
            Tu = make_float3(g.tu.x, g.tu.y, g.tu.z);
            Tv = make_float3(g.tv.x, g.tv.y, g.tv.z);
            Tw = make_float3(g.tw.x, g.tw.y, g.tw.z);
        } else {
            // re-evaluated with scale_modifier ignored and Wb/Hb from float truncation
            // (S/backward.cu:484-512,614-615)
            quat_to_rot(q, c0, c1, c2);
            float3 L0 = make_float3(c0.x * sc.x, c0.y * sc.x, c0.z * sc.x);
            float3 L1 = make_float3(c1.x * sc.y, c1.y * sc.y, c1.z * sc.y);
            // P = world2ndc * ndc2pix with the truncated Wb, Hb
#pragma unroll
            for (int r = 0; r < 4; r++) {
                float x = __ldg(vc.proj + 4 * r), y = __ldg(vc.proj + 4 * r + 1), w = __ldg(vc.proj + 4 * r + 3);
                Pm[0][r] = x * ((float)Wb / 2.0f) + w * ((float)(Wb - 1) / 2.0f);
                Pm[1][r] = y * ((float)Hb / 2.0f) + w * ((float)(Hb - 1) / 2.0f);
                Pm[2][r] = w;
            }
            float t[3][3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                t[i][0] = L0.x * Pm[i][0] + L0.y * Pm[i][1] + L0.z * Pm[i][2];
                t[i][1] = L1.x * Pm[i][0] + L1.y * Pm[i][1] + L1.z * Pm[i][2];
                t[i][2] = p.x * Pm[i][0] + p.y * Pm[i][1] + p.z * Pm[i][2] + Pm[i][3];
            }
            Tu = make_float3(t[0][0], t[0][1], t[0][2]);
            Tv = make_float3(t[1][0], t[1][1], t[1][2]);
            Tw = make_float3(t[2][0], t[2][1], t[2][2]);
            normal = xformvec43(view, c2);
        }
        const bool aabb_branch = (dm2x != 0.f || dm2y != 0.f);
        if (aabb_branch) {  // S/backward.cu:522-550
            float3 tv3 = make_float3(9.0f, 9.0f, -1.0f);
            float d = tv3.x * Tw.x * Tw.x + tv3.y * Tw.y * Tw.y + tv3.z * Tw.z * Tw.z;
            float id = 1.0f / d;
            float3 f = make_float3(tv3.x * id, tv3.y * id, tv3.z * id);
            float3 dT0 = make_float3(dm2x * f.x * Tw.x, dm2x * f.y * Tw.y, dm2x * f.z * Tw.z);
            float3 dT1 = make_float3(dm2y * f.x * Tw.x, dm2y * f.y * Tw.y, dm2y * f.z * Tw.z);
            float3 dT3 = make_float3(dm2x * f.x * Tu.x + dm2y * f.x * Tv.x, dm2x * f.y * Tu.y + dm2y * f.y * Tv.y,
                                     dm2x * f.z * Tu.z + dm2y * f.z * Tv.z);
            float3 dLdf = make_float3(dm2x * Tu.x * Tw.x + dm2y * Tv.x * Tw.x, dm2x * Tu.y * Tw.y + dm2y * Tv.y * Tw.y,
                                      dm2x * Tu.z * Tw.z + dm2y * Tv.z * Tw.z);
            float dLdd = (dLdf.x * f.x + dLdf.y * f.y + dLdf.z * f.z) * (-1.0f / d);
            dT3.x += dLdd * (tv3.x * Tw.x * 2.0f);
            dT3.y += dLdd * (tv3.y * Tw.y * 2.0f);
            dT3.z += dLdd * (tv3.z * Tw.z * 2.0f);
            dT[0] += dT0.x; dT[1] += dT0.y; dT[2] += dT0.z;
            dT[3] += dT1.x; dT[4] += dT1.y; dT[5] += dT1.z;
            dT[6] += dT3.x; dT[7] += dT3.y; dT[8] += dT3.z;
            if (precomp) {
#pragma unroll
                for (int i = 0; i < 9; i++) dTout[i] = dT[i];
            }
        }
        if (!precomp) {
            // dL_dM = P * transpose(dL_dT), S/backward.cu:555
            float dM[3][3];
#pragma unroll
            for (int j = 0; j < 3; j++)
#pragma unroll
                for (int r = 0; r < 3; r++)
                    dM[j][r] = Pm[0][r] * dT[j] + Pm[1][r] * dT[3 + j] + Pm[2][r] * dT[6 + j];
            float3 dtn = xformvec43T(view, dnrm);
            float3 pv = xform43(view, p);
            float cosv = -(pv.x * normal.x + pv.y * normal.y + pv.z * normal.z);
            float mult = cosv > 0.f ? 1.f : -1.f;
            dtn = make_float3(mult * dtn.x, mult * dtn.y, mult * dtn.z);
            float3 dRS0 = make_float3(dM[0][0], dM[0][1], dM[0][2]);
            float3 dRS1 = make_float3(dM[1][0], dM[1][1], dM[1][2]);
            drot = quat_vjp(q, make_float3(dRS0.x * sc.x, dRS0.y * sc.x, dRS0.z * sc.x),
                            make_float3(dRS1.x * sc.y, dRS1.y * sc.y, dRS1.z * sc.y), dtn);
            dscale = make_float2(dot3(dRS0, c0), dot3(dRS1, c1));
            dmean = make_float3(dM[2][0], dM[2][1], dM[2][2]);
        }
        // densification hack, S/backward.cu:633-636: uses the stored dL_dtransMat and T[8]
        float depth = g.tw.z;
        m2x = dTout[2] * depth * 0.5f * (float)Wb;
        m2y = dTout[5] * depth * 0.5f * (float)Hb;
    }
    dL_dmean2D[3 * (size_t)idx + 0] = m2x;
    dL_dmean2D[3 * (size_t)idx + 1] = m2y;
    dL_dmean2D[3 * (size_t)idx + 2] = 0.f;
    dL_dnormal[3 * (size_t)idx + 0] = dnrm.x; dL_dnormal[3 * (size_t)idx + 1] = dnrm.y; dL_dnormal[3 * (size_t)idx + 2] = dnrm.z;
    dL_dopacity[idx] = dopa;
    dL_dcolor[3 * (size_t)idx + 0] = dcol.x; dL_dcolor[3 * (size_t)idx + 1] = dcol.y; dL_dcolor[3 * (size_t)idx + 2] = dcol.z;
    dL_dmean3D[3 * (size_t)idx + 0] = dmean.x; dL_dmean3D[3 * (size_t)idx + 1] = dmean.y; dL_dmean3D[3 * (size_t)idx + 2] = dmean.z;
#pragma unroll
    for (int i = 0; i < 9; i++) dL_dtransMat[9 * (size_t)idx + i] = dTout[i];
    if (dL_dscale) { dL_dscale[2 * (size_t)idx] = dscale.x; dL_dscale[2 * (size_t)idx + 1] = dscale.y; }
    if (dL_drot) reinterpret_cast<float4*>(dL_drot)[idx] = drot;
}

__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const ViewParams vc,
                                    uint8_t* __restrict__ present) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float3 p = make_float3(means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]);
    float view[16];
    load16(vc.view, view);
    present[idx] = xform43(view, p).z > 0.2f;
}

}  // namespace gsr
