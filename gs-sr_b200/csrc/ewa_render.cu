// ewa_render.cu -- record build, per-tile front-to-back blend and back-to-front gradient walk of the
// EWA-splat rasterizers (3DGS and its PGSR plane superset) for sm_100a.
//
// Result contract = the reference kernels
//   forward  renderCUDA  G/cuda_rasterizer/forward.cu:261-374, L/cuda_rasterizer/forward.cu:273-407
//            (L adds: all_map blend :376-379, out_observe count :381-384, plane depth :401-405)
//   backward renderCUDA  G/cuda_rasterizer/backward.cu:399-557, L/cuda_rasterizer/backward.cu:399-614
//            (L adds: plane-depth fold :469-481, all_map terms :563-579, |dL_dmean2D| :602-603)
//
// B200 design (shared with the surfel kernels, see surfel_render_fwd.cu / surfel_render_bwd.cu):
// one CTA per 16x16 tile, each warp owns an 8x4 pixel block; the tile's sorted record stream is 3
// (4 with render_geo) contiguous float4 planes pulled through a double-buffered cp.async.bulk ring;
// every warp culls 32 entries at a time with the exact contribution-ellipse-vs-block test and only
// evaluates survivors; MUFU ex2/rcp; per-warp early exit.  The backward reduces each splat's 9..16
// gradient sums over the warp's 32 pixels in registers (transposed shuffle network) and flushes them
// with one burst of red.global.add.f32 into the Gaussian's 64-byte accumulator (the reference issues
// 9..16 global atomics per (pixel, splat) pair); out_observe is one int atomic per (warp, splat).
#include "common.cuh"
#include "async_copy.cuh"
#include "cull.cuh"
#include "render_common.cuh"
#include "tile_sort.cuh"
#include "ewa_common.cuh"

namespace gsr {

constexpr float LOG2E = 1.44269504088896340736f;

// ---- per-tile sort fused with record materialisation -------------------------------------------
template <bool GEO>
__global__ void __launch_bounds__(256)
ewa_build_records(const uint32_t* __restrict__ offsets, uint64_t* __restrict__ keys, const EwaGeom* __restrict__ geom,
                  const float* __restrict__ colors, const float* __restrict__ all_map, int gx,
                  float4* __restrict__ planes, size_t pstride) {
    __shared__ uint64_t skeys[SORT_SMEM_CAP];
    const int tile = blockIdx.x;
    const uint32_t begin = offsets[tile], end = offsets[tile + 1];
    const int n = (int)(end - begin);
    if (n == 0) return;
    __shared__ BucketSortSmem bs;
    const uint64_t* sorted = sort_tile_bucket(keys + begin, n, skeys, bs);
    const float ox = (float)((tile % gx) * TILE), oy = (float)((tile / gx) * TILE);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t g = (uint32_t)(sorted[i] & 0xffffffffull);
        const float4* gp = reinterpret_cast<const float4*>(geom + g);
        const float4 ga = __ldg(gp), gb = __ldg(gp + 1);
        const float cr = __ldg(colors + 3 * (size_t)g), cg = __ldg(colors + 3 * (size_t)g + 1),
                    cb = __ldg(colors + 3 * (size_t)g + 2);
        const uint32_t flag = ((int)gb.w == CULL_EXACT) ? 0u : REC_FLAG_ALWAYS;
        const size_t o = (size_t)begin + i;
        planes[0 * pstride + o] = make_float4(ga.x - ox, ga.y - oy, ga.z, ga.w);
        planes[1 * pstride + o] = make_float4(gb.x, gb.y, gb.z, __uint_as_float(g | flag));
        float am4 = 0.f;
        if (GEO) {
            const float* am = all_map + NUM_ALL_MAP * (size_t)g;
            planes[3 * pstride + o] = make_float4(__ldg(am), __ldg(am + 1), __ldg(am + 2), __ldg(am + 3));
            am4 = __ldg(am + 4);
        }
        planes[2 * pstride + o] = make_float4(cr, cg, cb, am4);
    }
}
template __global__ void ewa_build_records<false>(const uint32_t*, uint64_t*, const EwaGeom*, const float*, const float*, int, float4*, size_t);
template __global__ void ewa_build_records<true>(const uint32_t*, uint64_t*, const EwaGeom*, const float*, const float*, int, float4*, size_t);

template <int NPL>
__device__ __forceinline__ void ewa_issue_batch(float4 (*dst)[RBATCH], const float4* __restrict__ src, size_t pstride,
                                                int first, int count, uint64_t* bar) {
    const uint32_t bytes = (uint32_t)count * 16u;
    mbar_expect_tx(bar, bytes * NPL);
#pragma unroll
    for (int pl = 0; pl < NPL; pl++) bulk_g2s(&dst[pl][0], src + pl * pstride + first, bytes, bar);
}

// Can record (p0, p1) reach alpha >= 1/255 inside the (widened) block rectangle?
__device__ __forceinline__ bool ewa_entry_hits_block(float4 p0, float4 p1, float x0, float x1, float y0, float y1) {
    if (__float_as_uint(p1.w) & REC_FLAG_ALWAYS) return true;
    const Quadric q = {p0.z, p0.w, p1.x, 0.f, 0.f, -p1.z};
    return ellipse_hits_rect(q, x0 - p0.x, x1 - p0.x, y0 - p0.y, y1 - p0.y);
}

struct EwaEval {
    bool valid;
    float alpha, G, dx, dy;
};
// G/forward.cu:331-347: power, alpha and the two `continue` tests.
__device__ __forceinline__ EwaEval ewa_eval(float4 p0, float4 p1, float fx, float fy) {
    EwaEval e;
    e.dx = p0.x - fx;
    e.dy = p0.y - fy;
    const float power = -0.5f * (p0.z * e.dx * e.dx + p1.x * e.dy * e.dy) - p0.w * e.dx * e.dy;
    e.G = fast_ex2(power * LOG2E);
    e.alpha = fminf(ALPHA_MAX, p1.y * e.G);
    e.valid = !(power > 0.0f) && !(e.alpha < ALPHA_MIN);
    return e;
}

// ---- forward ------------------------------------------------------------------------------------
template <bool GEO>
__global__ void __launch_bounds__(TILE_PIX)
ewa_render_fwd(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ planes, size_t pstride, int W, int H,
               int gx, const float* __restrict__ bg, float focal_x, float focal_y, float* __restrict__ final_T,
               uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, int* __restrict__ out_observe,
               float* __restrict__ out_all_map, float* __restrict__ out_plane_depth, float4* __restrict__ mark_plane) {
    // mark_plane != NULL (P < 2^23): bits 23..30 of the record's idx|flag word get "blended by warp block w" marks
    // for the backward (see common.cuh REC_USED_SHIFT and surfel_render_fwd.cu)
    constexpr int NPL = GEO ? EWA_PLANES_GEO : EWA_PLANES;
    __shared__ __align__(128) float4 sbuf[2][NPL][RBATCH];
    __shared__ __align__(8) uint64_t full_bar[2];

    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wx0 = (warp & 1) * 8, wy0 = (warp >> 1) * 4;
    const int lx = wx0 + (lane & 7), ly = wy0 + (lane >> 3);
    const int px = tx * TILE + lx, py = ty * TILE + ly;
    const bool inside = px < W && py < H;
    const float fx = (float)lx, fy = (float)ly;
    const float bx0 = (float)wx0 - CULL_MARGIN, bx1 = (float)(wx0 + 7) + CULL_MARGIN;
    const float by0 = (float)wy0 - CULL_MARGIN, by1 = (float)(wy0 + 3) + CULL_MARGIN;

    const uint32_t range_x = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - range_x);
    const int nb = (n + RBATCH - 1) / RBATCH;
    const float4* src = planes + range_x;
    const uint32_t idx_mask = mark_plane != nullptr ? REC_INDEX_MASK_USED : ~REC_FLAG_ALWAYS;

    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int b = 0; b < 2 && b < nb; b++)
            ewa_issue_batch<NPL>(sbuf[b], src, pstride, b * RBATCH, min(RBATCH, n - b * RBATCH), &full_bar[b]);

    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    float A0 = 0.f, A1 = 0.f, A2 = 0.f, A3 = 0.f, A4 = 0.f;
    uint32_t last_contrib = 0;
    bool done = !inside;
    bool warp_done = __all_sync(FULLMASK, done);

    for (int b = 0; b < nb; b++) {
        const int stage = b & 1;
        const uint32_t parity = (uint32_t)((b >> 1) & 1);
        const int cnt = min(RBATCH, n - b * RBATCH);
        if (!warp_done) {
            mbar_wait(&full_bar[stage], parity);
            const float4(*sb)[RBATCH] = sbuf[stage];
            for (int c0 = 0; c0 < cnt && !warp_done; c0 += 32) {
                const int e = c0 + lane;
                bool hit = false;
                if (e < cnt) hit = ewa_entry_hits_block(sb[0][e], sb[1][e], bx0, bx1, by0, by1);
                uint32_t m = __ballot_sync(FULLMASK, hit);
                uint32_t used = 0;
                while (m) {
                    const int bitpos = __ffs(m) - 1;
                    const int j = c0 + bitpos;
                    m &= m - 1;
                    const float4 p0 = sb[0][j], p1 = sb[1][j];
                    const EwaEval ev = ewa_eval(p0, p1, fx, fy);
                    bool valid = ev.valid && !done;
                    if (__any_sync(FULLMASK, valid)) {
                        const float test_T = T * (1.0f - ev.alpha);
                        if (valid && test_T < T_EPS) { done = true; valid = false; }
                        used |= 1u << bitpos;
                        if (valid) {
                            const float4 pc = sb[2][j];
                            const float w = ev.alpha * T;
                            C0 += pc.x * w; C1 += pc.y * w; C2 += pc.z * w;
                            if (GEO) {
                                const float4 pm = sb[3][j];
                                A0 += pm.x * w; A1 += pm.y * w; A2 += pm.z * w; A3 += pm.w * w; A4 += pc.w * w;
                            }
                        }
                        if (out_observe != nullptr) {   // L/forward.cu:381-384, T before the update
                            const uint32_t ob = __ballot_sync(FULLMASK, valid && T > 0.5f);
                            if (ob != 0u && lane == 0)
                                atomicAdd(out_observe + (__float_as_uint(p1.w) & idx_mask), __popc(ob));
                        }
                        if (valid) {
                            T = test_T;
                            last_contrib = (uint32_t)(b * RBATCH + j + 1);
                        }
                        if (__all_sync(FULLMASK, done)) { warp_done = true; break; }
                    }
                }
                if (mark_plane != nullptr && ((used >> lane) & 1u))
                    atomicOr(reinterpret_cast<uint32_t*>(mark_plane + range_x + b * RBATCH + c0 + lane) + 3,
                             1u << (REC_USED_SHIFT + warp));
            }
        }
        const int all_done = __syncthreads_and(warp_done);
        if (all_done) {
            if (threadIdx.x == 0 && b + 1 < nb) mbar_wait(&full_bar[(b + 1) & 1], (uint32_t)(((b + 1) >> 1) & 1));
            break;
        }
        if (threadIdx.x == 0 && b + 2 < nb) {
            fence_proxy_async();
            ewa_issue_batch<NPL>(sbuf[stage], src, pstride, (b + 2) * RBATCH, min(RBATCH, n - (b + 2) * RBATCH), &full_bar[stage]);
        }
    }

    if (inside) {
        const size_t N = (size_t)W * H;
        const size_t pid = (size_t)py * W + px;
        final_T[pid] = T;
        n_contrib[pid] = last_contrib;
        out_color[pid] = C0 + T * __ldg(bg);
        out_color[pid + N] = C1 + T * __ldg(bg + 1);
        out_color[pid + 2 * N] = C2 + T * __ldg(bg + 2);
        if (GEO) {
            out_all_map[pid] = A0; out_all_map[pid + N] = A1; out_all_map[pid + 2 * N] = A2;
            out_all_map[pid + 3 * N] = A3; out_all_map[pid + 4 * N] = A4;
            // L/forward.cu:304,404: ray through the pixel, cx = W/2, cy = H/2; the 1e-8 is a double
            const float rx = ((float)px - (float)W * 0.5f) / focal_x, ry = ((float)py - (float)H * 0.5f) / focal_y;
            out_plane_depth[pid] = (float)((double)A4 / -((double)(A0 * rx + A1 * ry + A2) + 1.0e-8));
        }
    }
}
template __global__ void ewa_render_fwd<false>(const uint32_t*, const float4*, size_t, int, int, int, const float*, float, float,
                                               float*, uint32_t*, float*, int*, float*, float*, float4*);
template __global__ void ewa_render_fwd<true>(const uint32_t*, const float4*, size_t, int, int, int, const float*, float, float,
                                              float*, uint32_t*, float*, int*, float*, float*, float4*);

// ---- backward -----------------------------------------------------------------------------------
// Transposed warp reduction of 16 per-lane values (see surfel_render_bwd.cu): afterwards lane L holds
// the 32-lane sum of v[idx], idx = ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1).
__device__ __forceinline__ float ewa_reduce16(float (&v)[16], int lane) {
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const float send = hi ? v[i] : v[i + 8];
            const float keep = hi ? v[i + 8] : v[i];
            v[i] = keep + __shfl_xor_sync(FULLMASK, send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float send = hi ? v[i] : v[i + 4];
            const float keep = hi ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(FULLMASK, send, 8);
        }
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float send = hi ? v[i] : v[i + 2];
            const float keep = hi ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(FULLMASK, send, 4);
        }
    }
    {
        const bool hi = lane & 2;
        const float send = hi ? v[0] : v[1];
        const float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULLMASK, send, 2);
    }
    v[0] += __shfl_xor_sync(FULLMASK, v[0], 1);
    return v[0];
}

// MODE 0: 3DGS (9 sums)   1: plane without render_geo (+ |dL_dmean2D|, 11 sums)   2: plane with render_geo (16 sums)
// CTA = 4 warps = half a tile (16x8 px), 3-stage record ring, warps decoupled (the last warp to arrive on a stage
// refills it; no __syncthreads() in the walk) -- same scheme as surfel_render_bwd.cu.
constexpr int EWA_BWD_WARPS = 4, EWA_BWD_SPLIT = 8 / EWA_BWD_WARPS, EWA_BWD_STAGES = 3;
int ewa_bwd_ctas_per_tile() { return EWA_BWD_SPLIT; }

// USED: walk the forward's "blended" marks in the record word instead of repeating the cull test (P < 2^23)
template <int MODE, bool USED>
__global__ void __launch_bounds__(EWA_BWD_WARPS * 32, 6)
ewa_render_bwd(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ planes, size_t pstride, int W, int H,
               int gx, const float* __restrict__ bg, float focal_x, float focal_y, const float* __restrict__ final_T,
               const uint32_t* __restrict__ n_contrib, const float* __restrict__ all_map_pixels,
               const float* __restrict__ dL_dpix, const float* __restrict__ dL_dout_all_map,
               const float* __restrict__ dL_dout_plane_depth, float* __restrict__ gacc) {
    constexpr bool GEO = MODE == 2;
    constexpr int NPL = GEO ? EWA_PLANES_GEO : EWA_PLANES;
    constexpr int NV = MODE == 0 ? 9 : (MODE == 1 ? 11 : 16);
    __shared__ __align__(128) float4 sbuf[EWA_BWD_STAGES][NPL][RBATCH];
    __shared__ __align__(8) uint64_t full_bar[EWA_BWD_STAGES];
    __shared__ int s_arrive[EWA_BWD_STAGES];
    __shared__ int s_maxlast;

    const int tile = blockIdx.x / EWA_BWD_SPLIT;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = (threadIdx.x >> 5) + (blockIdx.x % EWA_BWD_SPLIT) * EWA_BWD_WARPS, lane = threadIdx.x & 31;
    const int wx0 = (warp & 1) * 8, wy0 = (warp >> 1) * 4;
    const int lx = wx0 + (lane & 7), ly = wy0 + (lane >> 3);
    const int px = tx * TILE + lx, py = ty * TILE + ly;
    const bool inside = px < W && py < H;
    const float fx = (float)lx, fy = (float)ly;
    const float bx0 = (float)wx0 - CULL_MARGIN, bx1 = (float)(wx0 + 7) + CULL_MARGIN;
    const float by0 = (float)wy0 - CULL_MARGIN, by1 = (float)(wy0 + 3) + CULL_MARGIN;
    const size_t N = (size_t)W * H;
    const size_t pid = (size_t)py * W + px;

    const uint32_t range_x = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - range_x);
    const float4* src = planes + range_x;

    const int last = inside ? (int)n_contrib[pid] : 0;   // entries [0, last) contribute
    const int wlast = __reduce_max_sync(FULLMASK, last);
    if (threadIdx.x == 0) {
        s_maxlast = 0;
#pragma unroll
        for (int i = 0; i < EWA_BWD_STAGES; i++) { mbar_init(&full_bar[i], 1); s_arrive[i] = 0; }
        mbar_fence_init();
    }
    __syncthreads();
    if (lane == 0 && wlast > 0) atomicMax(&s_maxlast, wlast);
    __syncthreads();
    const int maxlast = min(s_maxlast, n);
    if (maxlast <= 0) return;
    const int nb = (maxlast + RBATCH - 1) / RBATCH;

    auto batch_count = [&](int b) { return min(RBATCH, n - b * RBATCH); };
    if (threadIdx.x == 0)
        for (int i = 0; i < EWA_BWD_STAGES && nb - 1 - i >= 0; i++) {
            const int b = nb - 1 - i;
            ewa_issue_batch<NPL>(sbuf[i], src, pstride, b * RBATCH, batch_count(b), &full_bar[i]);
        }

    const float T_final = inside ? final_T[pid] : 0.f;
    float dpx0 = 0.f, dpx1 = 0.f, dpx2 = 0.f;
    float dam0 = 0.f, dam1 = 0.f, dam2 = 0.f, dam3 = 0.f, dam4 = 0.f;
    if (inside) {
        dpx0 = dL_dpix[pid]; dpx1 = dL_dpix[pid + N]; dpx2 = dL_dpix[pid + 2 * N];
        if (GEO) {   // L/backward.cu:469-481: fold dL/dplane_depth into dL/dout_all_map[0,1,2,4]
            dam0 = dL_dout_all_map[pid]; dam1 = dL_dout_all_map[pid + N]; dam2 = dL_dout_all_map[pid + 2 * N];
            dam3 = dL_dout_all_map[pid + 3 * N]; dam4 = dL_dout_all_map[pid + 4 * N];
            const float rx = (float)(((double)px - W * 0.5) / (double)focal_x);
            const float ry = (float)(((double)py - H * 0.5) / (double)focal_y);
            const float n0 = all_map_pixels[pid], n1 = all_map_pixels[pid + N], n2 = all_map_pixels[pid + 2 * N];
            const float distance = all_map_pixels[pid + 4 * N];
            const float tmp = (float)((double)(n0 * rx + n1 * ry + n2) + 1.0e-8);
            const float dpd = dL_dout_plane_depth[pid];
            dam4 += (-dpd / tmp);
            dam0 += dpd * (distance / (tmp * tmp) * rx);
            dam1 += dpd * (distance / (tmp * tmp) * ry);
            dam2 += dpd * (distance / (tmp * tmp));
        }
    }
    const float bg_dot_dpixel = __ldg(bg) * dpx0 + __ldg(bg + 1) * dpx1 + __ldg(bg + 2) * dpx2;
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;   // G/backward.cu:460-461

    float T = T_final;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ar0 = 0.f, ar1 = 0.f, ar2 = 0.f;
    float lm0 = 0.f, lm1 = 0.f, lm2 = 0.f, lm3 = 0.f, lm4 = 0.f, am0 = 0.f, am1 = 0.f, am2 = 0.f, am3 = 0.f, am4 = 0.f;

    int stage = 0;
    uint32_t parity = 0;
    for (int it = 0; it < nb; it++) {
        const int b = nb - 1 - it;
        const int cnt = batch_count(b);
        const int base = b * RBATCH;
        mbar_wait(&full_bar[stage], parity);
        if (base < wlast) {
            const float4(*sb)[RBATCH] = sbuf[stage];
            for (int c0 = ((cnt - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
                if (base + c0 >= wlast) continue;
                const int e = c0 + lane;
                bool hit = false;
                if (e < cnt && base + e < wlast) {
                    if (USED) hit = (__float_as_uint(sb[1][e].w) >> (REC_USED_SHIFT + warp)) & 1u;
                    else hit = ewa_entry_hits_block(sb[0][e], sb[1][e], bx0, bx1, by0, by1);
                }
                uint32_t m = __ballot_sync(FULLMASK, hit);
                while (m) {
                    const int bit = 31 - __clz(m);
                    m &= ~(1u << bit);
                    const int j = c0 + bit;
                    const int pos = base + j;
                    const float4 p0 = sb[0][j], p1 = sb[1][j];
                    const EwaEval ev = ewa_eval(p0, p1, fx, fy);
                    const bool valid = ev.valid && pos < last;
                    if (!__any_sync(FULLMASK, valid)) continue;

                    const float4 pc = sb[2][j];
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = 0.f;
                    if (valid) {
                        const float alpha = ev.alpha, G = ev.G;
                        const float ria = fast_rcp(1.0f - alpha);
                        T = T * ria;
                        const float w = alpha * T;
                        const float omla = 1.f - last_alpha;
                        ar0 = last_alpha * lc0 + omla * ar0; lc0 = pc.x;
                        ar1 = last_alpha * lc1 + omla * ar1; lc1 = pc.y;
                        ar2 = last_alpha * lc2 + omla * ar2; lc2 = pc.z;
                        float dL_dalpha = (pc.x - ar0) * dpx0 + (pc.y - ar1) * dpx1 + (pc.z - ar2) * dpx2;
                        v[6] = w * dpx0; v[7] = w * dpx1; v[8] = w * dpx2;
                        if (GEO) {
                            const float4 pm = sb[NPL - 1][j];
                            am0 = last_alpha * lm0 + omla * am0; lm0 = pm.x;
                            am1 = last_alpha * lm1 + omla * am1; lm1 = pm.y;
                            am2 = last_alpha * lm2 + omla * am2; lm2 = pm.z;
                            am3 = last_alpha * lm3 + omla * am3; lm3 = pm.w;
                            am4 = last_alpha * lm4 + omla * am4; lm4 = pc.w;
                            dL_dalpha += (pm.x - am0) * dam0 + (pm.y - am1) * dam1 + (pm.z - am2) * dam2 +
                                         (pm.w - am3) * dam3 + (pc.w - am4) * dam4;
                            v[11] = w * dam0; v[12] = w * dam1; v[13] = w * dam2; v[14] = w * dam3; v[15] = w * dam4;
                        }
                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final * ria) * bg_dot_dpixel;
                        const float dL_dG = p1.y * dL_dalpha;
                        const float gdx = G * ev.dx, gdy = G * ev.dy;
                        const float dG_ddelx = -gdx * p0.z - gdy * p0.w;
                        const float dG_ddely = -gdy * p1.x - gdx * p0.w;
                        v[0] = dL_dG * dG_ddelx * ddelx_dx;
                        v[1] = dL_dG * dG_ddely * ddely_dy;
                        v[2] = -0.5f * gdx * ev.dx * dL_dG;
                        v[3] = -0.5f * gdx * ev.dy * dL_dG;
                        v[4] = -0.5f * gdy * ev.dy * dL_dG;
                        v[5] = G * dL_dalpha;
                        if (MODE != 0) { v[9] = fabsf(v[0]); v[10] = fabsf(v[1]); }
                    }
                    const float red = ewa_reduce16(v, lane);
                    const uint32_t g = __float_as_uint(p1.w) & (USED ? REC_INDEX_MASK_USED : ~REC_FLAG_ALWAYS);
                    const int vi = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    if ((lane & 1) == 0 && vi < NV) atomicAdd(gacc + (size_t)g * EWA_GACC + vi, red);
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            const int old = atomicAdd(&s_arrive[stage], 1);
            if ((old & (EWA_BWD_WARPS - 1)) == EWA_BWD_WARPS - 1 && it + EWA_BWD_STAGES < nb) {
                __threadfence_block();
                fence_proxy_async();
                const int b2 = nb - 1 - (it + EWA_BWD_STAGES);
                ewa_issue_batch<NPL>(sbuf[stage], src, pstride, b2 * RBATCH, batch_count(b2), &full_bar[stage]);
            }
        }
        if (++stage == EWA_BWD_STAGES) { stage = 0; parity ^= 1u; }
    }
}
template __global__ void ewa_render_bwd<0, false>(const uint32_t*, const float4*, size_t, int, int, int, const float*, float, float,
    const float*, const uint32_t*, const float*, const float*, const float*, const float*, float*);
template __global__ void ewa_render_bwd<0, true>(const uint32_t*, const float4*, size_t, int, int, int, const float*, float, float,
    const float*, const uint32_t*, const float*, const float*, const float*, const float*, float*);
template __global__ void ewa_render_bwd<1, false>(const uint32_t*, const float4*, size_t, int, int, int, const float*, float, float,
    const float*, const uint32_t*, const float*, const float*, const float*, const float*, float*);
template __global__ void ewa_render_bwd<1, true>(const uint32_t*, const float4*, size_t, int, int, int, const float*, float, float,
    const float*, const uint32_t*, const float*, const float*, const float*, const float*, float*);
template __global__ void ewa_render_bwd<2, false>(const uint32_t*, const float4*, size_t, int, int, int, const float*, float, float,
    const float*, const uint32_t*, const float*, const float*, const float*, const float*, float*);
template __global__ void ewa_render_bwd<2, true>(const uint32_t*, const float4*, size_t, int, int, int, const float*, float, float,
    const float*, const uint32_t*, const float*, const float*, const float*, const float*, float*);

}  // namespace gsr
