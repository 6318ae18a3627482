// cull.cuh -- conservative "can this splat reach alpha >= 1/255 anywhere in this pixel
// rectangle?" test, shared by tile binning (16x16 tiles) and the render kernels (8x4 warp blocks).
//
// A surfel contributes to pixel (x,y) only if (S/cuda_rasterizer/forward.cu:351-389)
//      alpha = min(.99, o * exp(-rho/2)) >= 1/255,  rho = min(rho3d, rho2d)
//   <=> rho <= tau := 2 ln(255 o)
//   <=> (x,y) in  E = { rho3d <= tau }  or  (x,y) in  Dsc = { |(x,y) - c|^2 <= tau/2 }.
// With p(x,y) = a x + b y + c0 (adjugate rows of the splat->pixel homography) the ray-splat
// intersection is s = p.xy / p.z, so  rho3d <= tau  <=>  q(x,y) = p.x^2 + p.y^2 - tau p.z^2 <= 0:
// a conic.  When its quadratic part is positive definite it is an ellipse and the test below is
// exact for the continuous rectangle (hence conservative for the pixel centres inside it);
// otherwise (the tau-disc of the splat reaches the camera plane) the caller must treat the
// splat as "always evaluate".  tau carries slack for exp/log rounding and the rectangle is
// widened by a quarter pixel.  That margin does NOT cover edge-on surfels: their conic is a sliver
// whose coefficients are ~1e6 while q is only ~0.1 deep inside it, below float32 resolution of the
// cancelling terms (found at cfg-B: one pair in 1e9, tests/test_fullsize_parity_gpu.py).  Every
// comparison therefore carries an explicit rounding bound QUADRIC_EPS x (sum of the magnitudes of
// the terms that cancel), so a decision that float32 cannot make reliably falls on "evaluate".
#pragma once
#include "common.cuh"

namespace gsr {

constexpr float CULL_MARGIN = 0.25f;
constexpr float QUADRIC_EPS = 1e-6f;     // ~16 ulp: rounding bound, relative to the magnitude of the cancelling terms

struct Quadric {           // q(x,y) = xx x^2 + 2 xy x y + yy y^2 + 2 bx x + 2 by y + c0
    float xx, xy, yy, bx, by, c0;
};

// tau (with slack) for an opacity; < 0 means the splat can never reach 1/255.
// A NaN opacity is not "never": the reference's alpha = min(0.99f, opacity * G) is fminf, which drops the NaN, so such a
// Gaussian blends with alpha 0.99 on every pixel of its rectangle (S/forward.cu:384, G/forward.cu:349).  TAU_ALWAYS tells
// the callers to skip the culling for it.
constexpr float TAU_ALWAYS = 3.0e38f;
__device__ __forceinline__ float contribution_tau(float opacity) {
    float a = 255.0f * opacity;
    if (a != a) return TAU_ALWAYS;
    if (!(a >= 1.0f)) return -1.0f;
    return 2.0f * __logf(a) * 1.001f + 2e-3f;
}

__device__ __forceinline__ Quadric make_quadric(float3 a, float3 b, float3 c, float tau) {
    Quadric q;
    q.xx = a.x * a.x + a.y * a.y - tau * a.z * a.z;
    q.xy = a.x * b.x + a.y * b.y - tau * a.z * b.z;
    q.yy = b.x * b.x + b.y * b.y - tau * b.z * b.z;
    q.bx = a.x * c.x + a.y * c.y - tau * a.z * c.z;
    q.by = b.x * c.x + b.y * c.y - tau * b.z * c.z;
    q.c0 = c.x * c.x + c.y * c.y - tau * c.z * c.z;
    return q;
}

// true when the quadratic part is safely positive definite (ellipse)
__device__ __forceinline__ bool quadric_is_ellipse(const Quadric& q) {
    float det = q.xx * q.yy - q.xy * q.xy;
    return (q.xx > 0.f) && (q.yy > 0.f) && (det > 1e-6f * q.xx * q.yy);
}

__device__ __forceinline__ float quadric_eval(const Quadric& q, float x, float y) {
    return (q.xx * x + 2.f * (q.xy * y + q.bx)) * x + (q.yy * y + 2.f * q.by) * y + q.c0;
}

// Ellipse {q <= 0} vs rectangle [x0,x1]x[y0,y1]; requires quadric_is_ellipse(q).  Conservative under float32
// rounding: E bounds the absolute error of any evaluation of q inside the rectangle.
__device__ __forceinline__ bool ellipse_hits_rect(const Quadric& q, float x0, float x1, float y0, float y1) {
    const float X = fmaxf(fabsf(x0), fabsf(x1)), Y = fmaxf(fabsf(y0), fabsf(y1));
    const float axy = fabsf(q.xy), abx = fabsf(q.bx), aby = fabsf(q.by);
    const float E = QUADRIC_EPS * ((q.xx * X + 2.f * (axy * Y + abx)) * X + (q.yy * Y + 2.f * aby) * Y + fabsf(q.c0));
    const float det = q.xx * q.yy - q.xy * q.xy;            // > 0
    // centre e solves [xx xy; xy yy] e = -(bx, by); compare without dividing
    const float ex = q.xy * q.by - q.yy * q.bx, ey = q.xy * q.bx - q.xx * q.by;   // = det * centre
    const float dm = QUADRIC_EPS * (q.xx * q.yy + q.xy * q.xy);                   // rounding of det (it cancels for rotated slivers)
    const float tx = QUADRIC_EPS * (axy * aby + q.yy * abx) + X * dm, ty = QUADRIC_EPS * (axy * abx + q.xx * aby) + Y * dm;
    if (ex >= x0 * det - tx && ex <= x1 * det + tx && ey >= y0 * det - ty && ey <= y1 * det + ty) return true;
    // corners
    if (quadric_eval(q, x0, y0) <= E || quadric_eval(q, x1, y0) <= E || quadric_eval(q, x0, y1) <= E ||
        quadric_eval(q, x1, y1) <= E) return true;
    // edges: interior minimum of the 1-D restriction  t -> A t^2 + 2 B t + C  is C - B^2/A at t* = -B/A
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const float yc = k ? y1 : y0;
        const float B = q.xy * yc + q.bx, C = (q.yy * yc + 2.f * q.by) * yc + q.c0;
        if (-B > x0 * q.xx && -B < x1 * q.xx && C * q.xx <= B * B + (E * q.xx + QUADRIC_EPS * B * B)) return true;
        const float xc = k ? x1 : x0;
        const float B2 = q.xy * xc + q.by, C2 = (q.xx * xc + 2.f * q.bx) * xc + q.c0;
        if (-B2 > y0 * q.yy && -B2 < y1 * q.yy && C2 * q.yy <= B2 * B2 + (E * q.yy + QUADRIC_EPS * B2 * B2)) return true;
    }
    return false;
}

// Disc of squared radius r2 around (cx,cy) vs rectangle.
__device__ __forceinline__ bool disc_hits_rect(float cx, float cy, float r2, float x0, float x1, float y0, float y1) {
    const float dx = fmaxf(fmaxf(x0 - cx, cx - x1), 0.f), dy = fmaxf(fmaxf(y0 - cy, cy - y1), 0.f);
    return dx * dx + dy * dy <= r2;
}

// Culling record of one Gaussian: conic of {rho3d <= tau} in a frame shifted to the rounded
// screen centre (keeps the float32 coefficients well conditioned), low-pass disc radius, mode.
__device__ __forceinline__ CullRec make_cull_rec(float3 Tu, float3 Tv, float3 Tw, float cx, float cy,
                                                 float opacity, bool no_cull) {
    CullRec r;
    const float sx = rintf(cx), sy = rintf(cy);
    float tau = contribution_tau(opacity);
    int mode = CULL_EXACT;
    Quadric q = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (no_cull) { mode = CULL_ALWAYS; tau = fmaxf(tau, 0.f); }
    else if (tau < 0.f) { mode = CULL_NEVER; tau = 0.f; }
    else if (tau == TAU_ALWAYS) mode = CULL_ALWAYS;
    else {
        const float3 tu = make_float3(fmaf(-sx, Tw.x, Tu.x), fmaf(-sx, Tw.y, Tu.y), fmaf(-sx, Tw.z, Tu.z));
        const float3 tv = make_float3(fmaf(-sy, Tw.x, Tv.x), fmaf(-sy, Tw.y, Tv.y), fmaf(-sy, Tw.z, Tv.z));
        q = make_quadric(cross3(tv, Tw), cross3(Tw, tu), cross3(tu, tv), tau);
        const bool finite = fabsf(q.xx) < 3e38f && fabsf(q.yy) < 3e38f && fabsf(q.c0) < 3e38f &&
                            fabsf(sx) < 1e7f && fabsf(sy) < 1e7f;
        if (!finite || !quadric_is_ellipse(q)) mode = CULL_ALWAYS;
    }
    r.q0 = make_float4(q.xx, q.xy, q.yy, q.bx);
    r.q1 = make_float4(q.by, q.c0, 0.5f * tau, tau);
    r.q2 = make_float4(sx, sy, (float)mode, 0.f);
    return r;
}

// ---- tile-level test -------------------------------------------------------------------
// The forward preprocess COUNTS tiles with tile_may_contribute and duplicate_with_keys EMITS
// with it; both must take bit-identical decisions although they are inlined into different
// kernels.  Every operation below is therefore an explicit round-to-nearest intrinsic
// (__fmul_rn/__fadd_rn/__fmaf_rn are never contracted or re-associated by the compiler).
__device__ __forceinline__ float det_eval_rn(float xx, float xy, float yy, float bx, float by, float c0, float x, float y) {
    // (xx x + 2 (xy y + bx)) x + (yy y + 2 by) y + c0
    const float t0 = __fmaf_rn(xx, x, __fmul_rn(2.f, __fmaf_rn(xy, y, bx)));
    const float t1 = __fmaf_rn(yy, y, __fmul_rn(2.f, by));
    return __fadd_rn(__fmaf_rn(t0, x, __fmul_rn(t1, y)), c0);
}

// Branch-free on purpose: the sub-tests (disc, centre inside, four corners, four edges) are all evaluated and OR-ed.  The
// warp-cooperative callers run 32 different (Gaussian, tile) items per instruction; with early returns the items of a
// round left the function at ten different points and the remaining code re-ran per lane subset (1-4 active threads,
// 350 instructions per round instead of ~130 -- ncu, profiles/r02_prof_binning_summary.txt).
__device__ __forceinline__ bool tile_may_contribute(const CullRec& r, float cx, float cy, int tx, int ty) {
    const int mode = (int)r.q2.z;
    const float gx0 = __fadd_rn((float)(tx * TILE), -CULL_MARGIN), gx1 = __fadd_rn((float)(tx * TILE + TILE - 1), CULL_MARGIN);
    const float gy0 = __fadd_rn((float)(ty * TILE), -CULL_MARGIN), gy1 = __fadd_rn((float)(ty * TILE + TILE - 1), CULL_MARGIN);
    // low-pass disc
    const float ddx = fmaxf(fmaxf(__fadd_rn(gx0, -cx), __fadd_rn(cx, -gx1)), 0.f);
    const float ddy = fmaxf(fmaxf(__fadd_rn(gy0, -cy), __fadd_rn(cy, -gy1)), 0.f);
    int hit = __fmaf_rn(ddx, ddx, __fmul_rn(ddy, ddy)) <= r.q1.z;
    const float xx = r.q0.x, xy = r.q0.y, yy = r.q0.z, bx = r.q0.w, by = r.q1.x, c0 = r.q1.y;
    const float x0 = __fadd_rn(gx0, -r.q2.x), x1 = __fadd_rn(gx1, -r.q2.x);
    const float y0 = __fadd_rn(gy0, -r.q2.y), y1 = __fadd_rn(gy1, -r.q2.y);
    // rounding bounds, as in ellipse_hits_rect (explicit intrinsics: the two callers must agree bit for bit)
    const float X = fmaxf(fabsf(x0), fabsf(x1)), Y = fmaxf(fabsf(y0), fabsf(y1));
    const float axy = fabsf(xy), abx = fabsf(bx), aby = fabsf(by);
    const float m0 = __fmaf_rn(xx, X, __fmul_rn(2.f, __fmaf_rn(axy, Y, abx)));
    const float m1 = __fmaf_rn(yy, Y, __fmul_rn(2.f, aby));
    const float E = __fmul_rn(QUADRIC_EPS, __fadd_rn(__fmaf_rn(m0, X, __fmul_rn(m1, Y)), fabsf(c0)));
    const float det = __fmaf_rn(xx, yy, -__fmul_rn(xy, xy));
    const float ex = __fmaf_rn(xy, by, -__fmul_rn(yy, bx)), ey = __fmaf_rn(xy, bx, -__fmul_rn(xx, by));
    const float dm = __fmul_rn(QUADRIC_EPS, __fmaf_rn(xx, yy, __fmul_rn(xy, xy)));
    const float tolx = __fmaf_rn(X, dm, __fmul_rn(QUADRIC_EPS, __fmaf_rn(axy, aby, __fmul_rn(yy, abx))));
    const float toly = __fmaf_rn(Y, dm, __fmul_rn(QUADRIC_EPS, __fmaf_rn(axy, abx, __fmul_rn(xx, aby))));
    // centre of the ellipse inside the rectangle
    hit |= (ex >= __fadd_rn(__fmul_rn(x0, det), -tolx)) & (ex <= __fadd_rn(__fmul_rn(x1, det), tolx)) &
           (ey >= __fadd_rn(__fmul_rn(y0, det), -toly)) & (ey <= __fadd_rn(__fmul_rn(y1, det), toly));
    // a corner inside the ellipse
    hit |= (det_eval_rn(xx, xy, yy, bx, by, c0, x0, y0) <= E) | (det_eval_rn(xx, xy, yy, bx, by, c0, x1, y0) <= E) |
           (det_eval_rn(xx, xy, yy, bx, by, c0, x0, y1) <= E) | (det_eval_rn(xx, xy, yy, bx, by, c0, x1, y1) <= E);
    // an edge crossing the ellipse: interior minimum of the 1-D restriction
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const float yc = k ? y1 : y0;
        const float B = __fmaf_rn(xy, yc, bx);
        const float C = __fadd_rn(__fmul_rn(__fmaf_rn(yy, yc, __fmul_rn(2.f, by)), yc), c0);
        const float BB = __fmul_rn(B, B);
        hit |= (-B > __fmul_rn(x0, xx)) & (-B < __fmul_rn(x1, xx)) &
               (__fmul_rn(C, xx) <= __fadd_rn(BB, __fmaf_rn(E, xx, __fmul_rn(QUADRIC_EPS, BB))));
        const float xc = k ? x1 : x0;
        const float B2 = __fmaf_rn(xy, xc, by);
        const float C2 = __fadd_rn(__fmul_rn(__fmaf_rn(xx, xc, __fmul_rn(2.f, bx)), xc), c0);
        const float BB2 = __fmul_rn(B2, B2);
        hit |= (-B2 > __fmul_rn(y0, yy)) & (-B2 < __fmul_rn(y1, yy)) &
               (__fmul_rn(C2, yy) <= __fadd_rn(BB2, __fmaf_rn(E, yy, __fmul_rn(QUADRIC_EPS, BB2))));
    }
    return mode == CULL_ALWAYS ? true : (mode == CULL_NEVER ? false : hit != 0);
}

// ---- warp-cooperative tile counting ---------------------------------------------------------
// The forward preprocess kernels (one thread per Gaussian) count, per tile of the Gaussian's getRect rectangle, the
// tiles the splat can reach.  Rectangles are 4.7 tiles on average but 14.5 for the largest of a warp's 32 Gaussians
// (cfg-B), so a per-thread loop over the rectangle ran with 2-10 of 32 lanes active and made up 82 % of the kernel's
// instructions (ncu, profiles/r02b_*).  Here the warp flattens its 32 rectangles into one list of (Gaussian, tile)
// items (prefix sum of the areas) and every lane tests item base + lane: the owner is found by a 5-step binary search
// over the prefix (shuffles), its CullRec comes out of shared memory, and the decision is the SAME
// tile_may_contribute() call with the same operands, so counts, masks and the scatter's re-test agree bit for bit.
// ALL 32 lanes must call it (area = 0: nothing to test).  Returns the lane's own hit mask over the first 32 tiles
// of its rectangle (row-major).  s_rec: [4][32] float4 of this warp, s_mask: [32] words of this warp.
__device__ __forceinline__ uint32_t warp_count_tiles(const CullRec& cr, float cx, float cy, int x0, int y0, int w, int area,
                                                     int gx, uint32_t* __restrict__ tile_count, float4 (*s_rec)[32],
                                                     uint32_t* s_mask) {
    const int lane = threadIdx.x & 31;
    int incl = area;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - area;
    s_rec[0][lane] = cr.q0;
    s_rec[1][lane] = cr.q1;
    s_rec[2][lane] = make_float4(cr.q2.x, cr.q2.y, cr.q2.z, cx);
    s_rec[3][lane] = make_float4(cy, __int_as_float(x0), __int_as_float(y0), __int_as_float(w));
    s_mask[lane] = 0u;
    __syncwarp();
    for (int base = 0; base < total; base += 32) {
        const int t = base + lane;
        // owner = the last lane whose exclusive prefix is <= t (lanes with area 0 share their successor's prefix and are
        // never the last one below an item that exists)
        int o = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int e = __shfl_sync(0xffffffffu, excl, (o + step) & 31);
            if (e <= t) o += step;
        }
        const int k = t - __shfl_sync(0xffffffffu, excl, o);
        if (t < total) {
            CullRec r;
            r.q0 = s_rec[0][o];
            r.q1 = s_rec[1][o];
            const float4 a = s_rec[2][o], b = s_rec[3][o];
            r.q2 = make_float4(a.x, a.y, a.z, 0.f);
            const int ow = __float_as_int(b.w);
            const int ry = k / ow, rx = k - ry * ow;
            const int tx = __float_as_int(b.y) + rx, ty = __float_as_int(b.z) + ry;
            if (tile_may_contribute(r, a.w, b.x, tx, ty)) {
                atomicAdd(&tile_count[(size_t)(ty * gx + tx) * TILE_CTR_STRIDE], 1u);
                if (k < 32) atomicOr(&s_mask[o], 1u << k);
            }
        }
    }
    __syncwarp();
    return s_mask[lane];
}

// ---- CTA-cooperative tile counting of BIG rectangles ---------------------------------------------
// A screen-filling background / sky splat has thousands of tiles in its rectangle; inside the warp's list it kept ONE
// warp busy for hundreds of rounds (50 such splats among 500 k: preprocess 56 us -> 1.7 ms).  Rectangles of more than
// WARP_AREA_MAX tiles are therefore left out of the warp's list (area 0 there) and flattened over the whole 256-thread
// CTA: block-wide prefix of their areas, item base + thread per round, owner by binary search over the prefix in shared
// memory, the owner's record out of the s_rec rows warp_count_tiles has already written.  They never have a tile mask
// (> 32 tiles), only counts.  All threads of the CTA must call it; CTAs without a big rectangle leave after one barrier.
constexpr int WARP_AREA_MAX = 64;
__device__ __forceinline__ void cta_count_big_tiles(int area_big, int gx, uint32_t* __restrict__ tile_count,
                                                    float4 (*s_rec)[4][32], int* s_prefix /* [257] */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!__syncthreads_or(area_big > 0)) return;          // (also orders every warp's s_rec writes before the reads below)
    int incl = area_big;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    __shared__ int s_wsum[8];
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) wbase += w < warp ? s_wsum[w] : 0;
    s_prefix[tid] = wbase + incl - area_big;
    if (tid == 255) s_prefix[256] = wbase + incl;
    __syncthreads();
    const int total = s_prefix[256];
    for (int base = 0; base < total; base += 256) {
        const int t = base + tid;
        if (t < total) {
            int o = 0;                                     // the last thread whose exclusive prefix is <= t
#pragma unroll
            for (int step = 128; step > 0; step >>= 1)
                if (s_prefix[o + step] <= t) o += step;
            const int k = t - s_prefix[o];
            const int ow_ = o >> 5, ol = o & 31;
            CullRec r;
            r.q0 = s_rec[ow_][0][ol];
            r.q1 = s_rec[ow_][1][ol];
            const float4 a = s_rec[ow_][2][ol], b = s_rec[ow_][3][ol];
            r.q2 = make_float4(a.x, a.y, a.z, 0.f);
            const int rw = __float_as_int(b.w);
            const int ry = k / rw, rx = k - ry * rw;
            const int tx = __float_as_int(b.y) + rx, ty = __float_as_int(b.z) + ry;
            if (tile_may_contribute(r, a.w, b.x, tx, ty))
                atomicAdd(&tile_count[(size_t)(ty * gx + tx) * TILE_CTR_STRIDE], 1u);
        }
    }
}

}  // namespace gsr
