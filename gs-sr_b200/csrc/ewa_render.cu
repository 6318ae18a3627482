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
    const uint64_t* sorted = sort_tile_bucket(keys + begin, n, skeys, bs, reinterpret_cast<uint64_t*>(planes + begin));
    const float ox = (float)((tile % gx) * TILE), oy = (float)((tile / gx) * TILE);
#pragma unroll 2
    for (int i = threadIdx.x; i < n; i += blockDim.x) {      // two entries' gathers in flight per thread (as in sort_build_records)
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
    // the operation sequence nvcc emits for the reference's  -0.5f * (a dx dx + c dy dy) - b dx dy  (G/forward.cu:339, read
    // off oracle/_ref/libref_gaussian.so): fma(dx, dx a, dy (dy c)) ; fma(that, -0.5, -(dy (dx b))).  Spelled out so that
    // `power` -- and with it the power > 0 test -- is bit-identical to the reference whatever code surrounds this function
    // (left to the compiler, a refactoring of the forward changed the contraction and moved an Octree-PGSR iteration's
    // agreement with the reference kernels from 4.8e-4 to 1.1e-3).  dx, dy are exact in both (tile-local vs global pixels).
    const float quad = __fmaf_rn(e.dx, __fmul_rn(e.dx, p0.z), __fmul_rn(e.dy, __fmul_rn(e.dy, p1.x)));
    const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(e.dy, __fmul_rn(e.dx, p0.w)));
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

    // accumulators paired for the packed FMAs (render_common.cuh): C01 = colour.rg, C2A4 = (colour.b, all_map[4]) --
    // the (x, y) / (z, w) halves of the colour plane -- and A01, A23 = the halves of the all_map plane
    float T = 1.0f;
    float2 C01 = make_float2(0.f, 0.f), C2A4 = C01, A01 = C01, A23 = C01;
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
                            if (GEO) {      // 8 FMAs as 4 packed ones (measured: plane forward 650 -> 607 us at cfg-4)
                                const float4 pm = sb[3][j];
                                C01 = ffma2s(make_float2(pc.x, pc.y), w, C01);
                                C2A4 = ffma2s(make_float2(pc.z, pc.w), w, C2A4);
                                A01 = ffma2s(make_float2(pm.x, pm.y), w, A01);
                                A23 = ffma2s(make_float2(pm.z, pm.w), w, A23);
                            } else {        // three scalar FMAs (packing them made the 3DGS forward 5 % slower)
                                C01.x = fmaf(pc.x, w, C01.x); C01.y = fmaf(pc.y, w, C01.y); C2A4.x = fmaf(pc.z, w, C2A4.x);
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
                    }
                }
                if (mark_plane != nullptr && ((used >> lane) & 1u))
                    atomicOr(reinterpret_cast<uint32_t*>(mark_plane + range_x + b * RBATCH + c0 + lane) + 3,
                             1u << (REC_USED_SHIFT + warp));
                // one exit vote per chunk instead of one per survivor (as in surfel_render_fwd): a warp that finishes inside a
                // chunk skips the chunk's remaining survivors at the any-vote above
                warp_done = __all_sync(FULLMASK, done);
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
        const float C0 = C01.x, C1 = C01.y, C2 = C2A4.x, A0 = A01.x, A1 = A01.y, A2 = A23.x, A3 = A23.y, A4 = C2A4.y;
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
// Same two-phase scheme as surfel_render_bwd.cu.  Phase 1 (lane = pixel) walks the contributing splats back to front
// and parks TWO numbers per (pixel, splat) pair in a per-warp shared-memory slot: v = G dL/dalpha and the blend weight
// w.  Everything the reference accumulates per pair is linear in them once the splat's constants are factored out:
//   dL/dopacity = sum v;  dL/dcolour = sum w dL/dpixel;  dL/dall_map = sum w dL/dout_all_map;
//   dL/dmean2D, dL/dconic = opacity x (first / second MOMENTS of v about the splat centre) x conic terms
//   (G/backward.cu:507-548: dG_ddelx = -G (a dx + b dy), dL_dconic = -0.5 G d d^T dL_dG).
// Phase 2 (lane = pending splat x half of the block's pixels) therefore only accumulates 6 moments of v and the
// colour / all_map sums with one FMA each, shifts the moments from block-local pixel coordinates to the splat centre
// once per splat, and flushes with 128-bit red.global.add.v4.  |dL/dmean2D| (plane) is the one non-linear sum: it takes
// the absolute value per pair inside phase 2.  All linear per-pixel channels share one "blended behind" recurrence
// (S_i = <channel values of splat i, upstream grads>, rec <- rec + alpha (S - rec)), see surfel_render_bwd.cu.
__device__ __forceinline__ void ewa_red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// MODE 0: 3DGS (9 sums)   1: plane without render_geo (+ |dL_dmean2D|, 11 sums)   2: plane with render_geo (16 sums)
#ifndef GSR_EWA_BWD_MINB
#define GSR_EWA_BWD_MINB 3
#endif
#ifndef GSR_EWA_BWD_WARPS
#define GSR_EWA_BWD_WARPS 8
#endif
constexpr int EWA_BWD_WARPS = GSR_EWA_BWD_WARPS, EWA_BWD_SPLIT = 8 / EWA_BWD_WARPS, EWA_BWD_STAGES = 4, EWA_BWD_BATCH = 32;
constexpr int EWA_SLOTS = 16, EWA_LPS = 32 / EWA_SLOTS, EWA_PPL = 32 / EWA_LPS;     // pending pairs per warp; lanes / pixels per pair in phase 2
constexpr int EWA_SLOT_STRIDE = 32 * 2 + 2;                                           // words; (stride / 2) odd
int ewa_bwd_ctas_per_tile() { return EWA_BWD_SPLIT; }
template <int MODE>
constexpr size_t ewa_bwd_smem() {
    return (size_t)EWA_BWD_STAGES * (MODE == 2 ? EWA_PLANES_GEO : EWA_PLANES) * EWA_BWD_BATCH * 16      // ring
           + (size_t)EWA_BWD_WARPS * EWA_SLOTS * EWA_SLOT_STRIDE * 4                                     // pending pairs
           + (size_t)EWA_BWD_WARPS * EWA_SLOTS * 32                                                      // slot headers
           + (size_t)EWA_BWD_WARPS * EWA_LPS * (EWA_PPL * (MODE == 2 ? 2 : 1) + 1) * 16 + 128;           // pixel constants, misc
}

// USED: walk the forward's "blended" marks in the record word instead of repeating the cull test (P < 2^23)
template <int MODE, bool USED>
__global__ void __launch_bounds__(EWA_BWD_WARPS * 32, GSR_EWA_BWD_MINB)
ewa_render_bwd(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ planes, size_t pstride, int W, int H,
               int gx, const float* __restrict__ bg, float focal_x, float focal_y, const float* __restrict__ final_T,
               const uint32_t* __restrict__ n_contrib, const float* __restrict__ all_map_pixels,
               const float* __restrict__ dL_dpix, const float* __restrict__ dL_dout_all_map,
               const float* __restrict__ dL_dout_plane_depth, float* __restrict__ gacc, int sync_ring) {
    constexpr bool GEO = MODE == 2;
    constexpr int NPL = GEO ? EWA_PLANES_GEO : EWA_PLANES;
    constexpr int PIXV = GEO ? 2 : 1;                 // float4 of upstream gradients per pixel
    extern __shared__ __align__(128) unsigned char ewa_smem[];
    float4(*sbuf)[NPL][EWA_BWD_BATCH] = reinterpret_cast<float4(*)[NPL][EWA_BWD_BATCH]>(ewa_smem);
    constexpr size_t O_PEND = (size_t)EWA_BWD_STAGES * NPL * EWA_BWD_BATCH * 16;
    constexpr size_t O_HDR = O_PEND + (size_t)EWA_BWD_WARPS * EWA_SLOTS * EWA_SLOT_STRIDE * 4;
    constexpr size_t O_PIX = O_HDR + (size_t)EWA_BWD_WARPS * EWA_SLOTS * 32;
    constexpr int PIX_GROUP = EWA_PPL * PIXV + 1;   // float4 per phase-2 lane group; odd spacing keeps the groups' loads in different banks
    constexpr size_t O_MISC = O_PIX + (size_t)EWA_BWD_WARPS * EWA_LPS * PIX_GROUP * 16;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ewa_smem + O_MISC);
    int* s_arrive = reinterpret_cast<int*>(full_bar + 8);
    int* s_maxlast = s_arrive + 8;

    const int tile = blockIdx.x / EWA_BWD_SPLIT;
    const int tx = tile % gx, ty = tile / gx;
    const int wic = threadIdx.x >> 5;
    const int warp = wic + (blockIdx.x % EWA_BWD_SPLIT) * EWA_BWD_WARPS, lane = threadIdx.x & 31;
    const int wx0 = (warp & 1) * 8, wy0 = (warp >> 1) * 4;
    const int lx = wx0 + (lane & 7), ly = wy0 + (lane >> 3);
    const int px = tx * TILE + lx, py = ty * TILE + ly;
    const bool inside = px < W && py < H;
    const float fx = (float)lx, fy = (float)ly;
    const float bx0 = (float)wx0 - CULL_MARGIN, bx1 = (float)(wx0 + 7) + CULL_MARGIN;
    const float by0 = (float)wy0 - CULL_MARGIN, by1 = (float)(wy0 + 3) + CULL_MARGIN;
    const size_t N = (size_t)W * H;
    const size_t pid = (size_t)py * W + px;
    float* pend = reinterpret_cast<float*>(ewa_smem + O_PEND) + (size_t)wic * EWA_SLOTS * EWA_SLOT_STRIDE;
    float4* hdr = reinterpret_cast<float4*>(ewa_smem + O_HDR) + wic * EWA_SLOTS * 2;
    float4* pixc = reinterpret_cast<float4*>(ewa_smem + O_PIX) + wic * EWA_LPS * PIX_GROUP;

    const uint32_t range_x = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - range_x);
    const float4* src = planes + range_x;

    const int last = inside ? (int)n_contrib[pid] : 0;   // entries [0, last) contribute
    const int wlast = __reduce_max_sync(FULLMASK, last);
    if (threadIdx.x == 0) {
        *s_maxlast = 0;
#pragma unroll
        for (int i = 0; i < EWA_BWD_STAGES; i++) { mbar_init(&full_bar[i], 1); s_arrive[i] = 0; }
        mbar_fence_init();
    }
    __syncthreads();
    if (lane == 0 && wlast > 0) atomicMax(s_maxlast, wlast);
    __syncthreads();
    const int maxlast = min(*s_maxlast, n);
    if (maxlast <= 0) return;
    const int nb = (maxlast + EWA_BWD_BATCH - 1) / EWA_BWD_BATCH;

    auto batch_count = [&](int b) { return min(EWA_BWD_BATCH, n - b * EWA_BWD_BATCH); };
    auto issue = [&](int stage, int b) {
        const uint32_t bytes = (uint32_t)batch_count(b) * 16u;
        mbar_expect_tx(&full_bar[stage], bytes * NPL);
#pragma unroll
        for (int pl = 0; pl < NPL; pl++) bulk_g2s(&sbuf[stage][pl][0], src + pl * pstride + b * EWA_BWD_BATCH, bytes, &full_bar[stage]);
    };
    if (threadIdx.x == 0)
        for (int i = 0; i < EWA_BWD_STAGES && nb - 1 - i >= 0; i++) issue(i, nb - 1 - i);

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
    const float k_tail = -T_final * (__ldg(bg) * dpx0 + __ldg(bg + 1) * dpx1 + __ldg(bg + 2) * dpx2);   // G/backward.cu:497-500
    {
        float4* mine = pixc + (lane / EWA_PPL) * PIX_GROUP + (lane % EWA_PPL) * PIXV;
        mine[0] = make_float4(dpx0, dpx1, dpx2, dam0);
        if (GEO) mine[1] = make_float4(dam1, dam2, dam3, dam4);
    }
    __syncwarp();

    // ---- phase 2 ----
    const int ps = lane % EWA_SLOTS, ph = lane / EWA_SLOTS;
    const float half_W = 0.5f * (float)W, half_H = 0.5f * (float)H;     // G/backward.cu:460-461
    auto flush = [&](int np) {
        __syncwarp();
        float s0 = 0.f, sx = 0.f, sy = 0.f, sxx = 0.f, sxy = 0.f, syy = 0.f, ax = 0.f, ay = 0.f;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
        float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
        if (ps < np) {
            h0 = hdr[ps * 2]; h1 = hdr[ps * 2 + 1];          // (index bits, centre x, centre y, opacity) (conic a, b, c, -)
            const float X0 = h0.y - (float)wx0, Y0 = h0.z - (float)(wy0 + ph * (EWA_PPL / 8));   // centre in this lane's pixel frame
            const float* pb = pend + ps * EWA_SLOT_STRIDE + ph * (EWA_PPL * 2);
            const float4* pk = pixc + ph * PIX_GROUP;
#pragma unroll
            for (int i = 0; i < EWA_PPL; i++) {
                const float2 u = *reinterpret_cast<const float2*>(pb + i * 2);        // v, w
                const float4 k0 = pk[i * PIXV];
                const float xi = (float)(i & 7), yi = (float)(i >> 3);
                s0 += u.x;
                if ((i & 7) != 0) { sx = fmaf(xi, u.x, sx); sxx = fmaf(xi * xi, u.x, sxx); }
                if ((i >> 3) != 0) { sy = fmaf(yi, u.x, sy); syy = fmaf(yi * yi, u.x, syy); }
                if ((i & 7) != 0 && (i >> 3) != 0) sxy = fmaf(xi * yi, u.x, sxy);
                if (MODE != 0) {                                                      // |dL_dmean2D| per pair (L/backward.cu:602-603)
                    const float dx = X0 - xi, dy = Y0 - yi, av = fabsf(u.x);
                    ax = fmaf(av, fabsf(fmaf(dx, h1.x, dy * h1.y)), ax);
                    ay = fmaf(av, fabsf(fmaf(dy, h1.z, dx * h1.y)), ay);
                }
                // (scalar on purpose: the packed form pushed the 80-register kernels into more spills, plane backward +6 %)
                c0 = fmaf(u.y, k0.x, c0); c1 = fmaf(u.y, k0.y, c1); c2 = fmaf(u.y, k0.z, c2);
                if (GEO) {
                    const float4 k1 = pk[i * PIXV + 1];
                    m0 = fmaf(u.y, k0.w, m0); m1 = fmaf(u.y, k1.x, m1); m2 = fmaf(u.y, k1.y, m2);
                    m3 = fmaf(u.y, k1.z, m3); m4 = fmaf(u.y, k1.w, m4);
                }
            }
            // moments about the splat centre: d = centre - pixel
            const float qx = X0 * s0 - sx, qy = Y0 * s0 - sy;
            const float qxx = fmaf(X0, X0 * s0 - 2.f * sx, sxx), qyy = fmaf(Y0, Y0 * s0 - 2.f * sy, syy);
            const float qxy = fmaf(X0, Y0 * s0 - sy, fmaf(-Y0, sx, sxy));
            sx = qx; sy = qy; sxx = qxx; syy = qyy; sxy = qxy;
        }
#pragma unroll
        for (int o = EWA_SLOTS; o < 32; o <<= 1) {
            s0 += __shfl_xor_sync(FULLMASK, s0, o); sx += __shfl_xor_sync(FULLMASK, sx, o); sy += __shfl_xor_sync(FULLMASK, sy, o);
            sxx += __shfl_xor_sync(FULLMASK, sxx, o); sxy += __shfl_xor_sync(FULLMASK, sxy, o); syy += __shfl_xor_sync(FULLMASK, syy, o);
            c0 += __shfl_xor_sync(FULLMASK, c0, o); c1 += __shfl_xor_sync(FULLMASK, c1, o); c2 += __shfl_xor_sync(FULLMASK, c2, o);
            if (MODE != 0) { ax += __shfl_xor_sync(FULLMASK, ax, o); ay += __shfl_xor_sync(FULLMASK, ay, o); }
            if (GEO) {
                m0 += __shfl_xor_sync(FULLMASK, m0, o); m1 += __shfl_xor_sync(FULLMASK, m1, o); m2 += __shfl_xor_sync(FULLMASK, m2, o);
                m3 += __shfl_xor_sync(FULLMASK, m3, o); m4 += __shfl_xor_sync(FULLMASK, m4, o);
            }
        }
        if (ph == 0 && ps < np) {
            const float op = h0.w, ca = h1.x, cb = h1.y, cc = h1.z;
            float* acc = gacc + (size_t)__float_as_uint(h0.x) * EWA_GACC;
            // dL_dG dG/ddelx = -opacity v (a dx + b dy);  dL_dconic = -0.5 opacity v d d^T
            const float g0 = -op * (ca * sx + cb * sy) * half_W, g1 = -op * (cc * sy + cb * sx) * half_H;
            ewa_red_add_v4(acc + 0, g0, g1, -0.5f * op * sxx, -0.5f * op * sxy);
            ewa_red_add_v4(acc + 4, -0.5f * op * syy, s0, c0, c1);
            if (MODE == 0) atomicAdd(acc + 8, c2);
            else ewa_red_add_v4(acc + 8, c2, op * ax * half_W, op * ay * half_H, m0);
            if (GEO) ewa_red_add_v4(acc + 12, m1, m2, m3, m4);
        }
        __syncwarp();
    };

    float T = T_final, rec = 0.f;
    int npend = 0;
    int stage = 0;
    uint32_t parity = 0;
    for (int it = 0; it < nb; it++) {
        const int b = nb - 1 - it;
        const int cnt = batch_count(b);
        const int base = b * EWA_BWD_BATCH;
        mbar_wait(&full_bar[stage], parity);
        if (base < wlast) {
            const float4(*sb)[EWA_BWD_BATCH] = sbuf[stage];
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
                    float2 o = make_float2(0.f, 0.f);
                    if (valid) {
                        const float alpha = ev.alpha;
                        const float ria = fast_rcp(1.0f - alpha);
                        T = T * ria;
                        float S = pc.x * dpx0;
                        S = fmaf(pc.y, dpx1, S); S = fmaf(pc.z, dpx2, S);
                        if (GEO) {
                            const float4 pm = sb[NPL - 1][j];
                            S = fmaf(pm.x, dam0, S); S = fmaf(pm.y, dam1, S); S = fmaf(pm.z, dam2, S);
                            S = fmaf(pm.w, dam3, S); S = fmaf(pc.w, dam4, S);
                        }
                        const float D = S - rec;
                        rec = fmaf(alpha, D, rec);
                        const float dL_dalpha = fmaf(D, T, k_tail * ria);
                        o = make_float2(ev.G * dL_dalpha, alpha * T);
                    }
                    *reinterpret_cast<float2*>(pend + npend * EWA_SLOT_STRIDE + lane * 2) = o;
                    if (lane == 0) {
                        const uint32_t g = __float_as_uint(p1.w) & (USED ? REC_INDEX_MASK_USED : ~REC_FLAG_ALWAYS);
                        hdr[npend * 2] = make_float4(__uint_as_float(g), p0.x, p0.y, p1.y);
                        hdr[npend * 2 + 1] = make_float4(p0.z, p0.w, p1.x, 0.f);
                    }
                    if (++npend == EWA_SLOTS) { flush(EWA_SLOTS); npend = 0; }
                }
            }
        }
        if (sync_ring) __syncthreads();          // validation only (gsr_set_option("dbg", 2)), see surfel_render_bwd.cu
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            const int old = atomicAdd(&s_arrive[stage], 1);
            if ((old & (EWA_BWD_WARPS - 1)) == EWA_BWD_WARPS - 1 && it + EWA_BWD_STAGES < nb) {
                __threadfence_block();
                fence_proxy_async();
                issue(stage, nb - 1 - (it + EWA_BWD_STAGES));
            }
        }
        if (++stage == EWA_BWD_STAGES) { stage = 0; parity ^= 1u; }
    }
    if (npend) flush(npend);
}

template <int MODE, bool USED>
static cudaError_t ewa_bwd_launch(dim3 grid, cudaStream_t s, const uint32_t* tile_offset, const float4* planes, size_t pstride, int W,
                                  int H, int gx, const float* bg, float focal_x, float focal_y, const float* final_T,
                                  const uint32_t* n_contrib, const float* amp, const float* dL_dpix, const float* dam, const float* dpd,
                                  float* gacc, int sync_ring) {
    static bool ready[64] = {};          // the opt-in shared-memory size is a per-device function attribute
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !ready[dev]) {
        e = cudaFuncSetAttribute(ewa_render_bwd<MODE, USED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ewa_bwd_smem<MODE>());
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) ready[dev] = true;
    }
    ewa_render_bwd<MODE, USED><<<grid, EWA_BWD_WARPS * 32, ewa_bwd_smem<MODE>(), s>>>(tile_offset, planes, pstride, W, H, gx, bg, focal_x,
                                                                                     focal_y, final_T, n_contrib, amp, dL_dpix, dam, dpd, gacc, sync_ring);
    return cudaGetLastError();
}

cudaError_t launch_ewa_render_bwd(int mode, bool used, int ntiles, const uint32_t* tile_offset, const float4* planes, size_t pstride,
                                  int W, int H, int gx, const float* bg, float focal_x, float focal_y, const float* final_T,
                                  const uint32_t* n_contrib, const float* amp, const float* dL_dpix, const float* dam, const float* dpd,
                                  float* gacc, int sync_ring, cudaStream_t s) {
    const dim3 grid(ntiles * EWA_BWD_SPLIT);
#define GSR_EWA_BWD(M, U) ewa_bwd_launch<M, U>(grid, s, tile_offset, planes, pstride, W, H, gx, bg, focal_x, focal_y, final_T, n_contrib, amp, dL_dpix, dam, dpd, gacc, sync_ring)
    if (mode == 2) return used ? GSR_EWA_BWD(2, true) : GSR_EWA_BWD(2, false);
    if (mode == 1) return used ? GSR_EWA_BWD(1, true) : GSR_EWA_BWD(1, false);
    return used ? GSR_EWA_BWD(0, true) : GSR_EWA_BWD(0, false);
#undef GSR_EWA_BWD
}

}  // namespace gsr
