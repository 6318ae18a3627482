// surfel_render_fwd.cu -- per-tile front-to-back alpha blend of 2D Gaussian surfels
// (colour + 11 auxiliary channels) for sm_100a.
//
// Result contract = reference renderCUDA, S/cuda_rasterizer/forward.cu:256-448
// (ray-splat intersection :351-368, low-pass :362-366, alpha :381-389, distortion
// moments :395-400, median bookkeeping :402-408, outputs :427-447).
//
// B200 design (not the reference's):
//   * one CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel block;
//   * the tile's sorted record stream is contiguous in each of six float4 planes in HBM:
//     one elected thread pulls 128-entry batches (6 x 2 KB cp.async.bulk, SASS UBLKCP) into
//     a double-buffered shared ring guarded by mbarriers, overlapping the next batch's
//     HBM/L2 latency with the current batch's blend;
//   * the ray-splat intersection uses the adjugate rows stored in the record:
//     p = a x + b y + c (6 FMA), s = p.xy / p.z, depth = det(T) / p.z -- algebraically the
//     reference's k x l / (s.Tw) form, in tile-local coordinates (small integers);
//   * each warp culls the batch 32 entries at a time: lane j tests entry j's conic
//     {rho3d <= tau} and low-pass disc against the warp's 8x4 block (exact ellipse-vs-
//     rectangle test, cull.cuh), one ballot, and only surviving entries are evaluated --
//     the skipped pairs provably have alpha < 1/255 (the reference `continue`s on them);
//   * MUFU rcp/ex2 approximations (1 ulp / 2^-22) replace IEEE division and expf;
//   * a warp stops as soon as its 32 pixels are done (the reference waits for all 256).
#include "common.cuh"
#include "async_copy.cuh"
#include "cull.cuh"
#include "render_common.cuh"

namespace gsr {

#ifndef GSR_FWD_WARPS
#define GSR_FWD_WARPS 8
#endif
constexpr int FWD_WARPS = GSR_FWD_WARPS;      // warps per CTA: 8 = whole 16x16 tile, 4 = half tile (16x8)
constexpr int FWD_SPLIT = 8 / FWD_WARPS;

// MARK: P < 2^23, the record word has room for the per-warp-block "blended" marks handed to the backward
template <bool MARK>
#ifndef GSR_FWD_PAIR2
#define GSR_FWD_PAIR2 1        // 1.079 -> 1.059 ms at cfg-B
#endif
#ifndef GSR_FWD_MINB
#define GSR_FWD_MINB (32 / FWD_WARPS)
#endif
__global__ void __launch_bounds__(FWD_WARPS * 32, GSR_FWD_MINB)
surfel_render_fwd(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ planes, size_t pstride, int W,
                  int H, int gx, const float* __restrict__ bg, float* __restrict__ final_T,
                  uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                  float* __restrict__ out_others, float4* __restrict__ mark_plane) {
    __shared__ __align__(128) float4 sbuf[2][REC_PLANES][RBATCH];
    __shared__ __align__(8) uint64_t full_bar[2];

    const int tile = blockIdx.x / FWD_SPLIT;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = (threadIdx.x >> 5) + (blockIdx.x % FWD_SPLIT) * FWD_WARPS, lane = threadIdx.x & 31;
    const int wx0 = (warp & 1) * 8, wy0 = (warp >> 1) * 4;
    const int lx = wx0 + (lane & 7), ly = wy0 + (lane >> 3);
    const int px = tx * TILE + lx, py = ty * TILE + ly;
    const bool inside = px < W && py < H;
    const float fx = (float)lx, fy = (float)ly;
    // warp block rectangle (continuous, widened) for the cull test
    const float bx0 = (float)wx0 - CULL_MARGIN, bx1 = (float)(wx0 + 7) + CULL_MARGIN;
    const float by0 = (float)wy0 - CULL_MARGIN, by1 = (float)(wy0 + 3) + CULL_MARGIN;

    const uint32_t range_x = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - range_x);
    const int nb = (n + RBATCH - 1) / RBATCH;
    const float4* src = planes + range_x;
    constexpr uint32_t idx_mask = MARK ? REC_INDEX_MASK_USED : ~REC_FLAG_ALWAYS;

    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int b = 0; b < 2 && b < nb; b++) issue_batch(sbuf[b], src, pstride, b * RBATCH, min(RBATCH, n - b * RBATCH), &full_bar[b]);

    float T = 1.0f;
    // accumulators paired for the packed FMAs (the record's (x, y) / (z, w) halves are the multiplicands):
    // NA = (normal.x, normal.y), NC = (normal.z, colour.r), CC = (colour.g, colour.b), DM = (depth, M1)
    float2 NA = make_float2(0.f, 0.f), NC = NA, CC = NA, DM = NA;
    float M2 = 0.f, dist = 0.f;
    uint32_t last_contrib = 0, med_contrib = 0;   // median depth / index / normal are re-derived from med_contrib after the walk
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
                if (e < cnt) hit = entry_hits_block(sb[0][e], sb[1][e], sb[2][e], sb[3][e], bx0, bx1, by0, by1);
                uint32_t m = __ballot_sync(FULLMASK, hit);
                uint32_t used = 0;                 // entries of this chunk blended on >= 1 pixel of this warp's block
                // blend of one evaluated survivor (entry j = c0 + bitpos); the state chain T / done runs through it in list order
                auto blend = [&](const PairEval& ev, int bitpos) {
                    const int j = c0 + bitpos;
                    bool valid = ev.valid && !done;
                    if (__any_sync(FULLMASK, valid)) {
                        const float test_T = T * (1.0f - ev.alpha);
                        if (valid && test_T < T_EPS) { done = true; valid = false; }
                        used |= 1u << bitpos;      // superset: also set when every contributing lane just terminated
                        if (valid) {
                            const float4 pn = sb[4][j], pc = sb[5][j];
                            const float w = ev.alpha * T;
                            const float A = 1.0f - T;
                            const float mm = MSCALE * (1.0f - NEAR_N * fast_rcp(ev.depth));
                            dist += (mm * mm * A + M2 - 2.0f * mm * DM.y) * w;
                            DM = ffma2s(make_float2(ev.depth, mm), w, DM);
                            M2 += mm * mm * w;
                            const uint32_t pos = (uint32_t)(b * RBATCH + j + 1);
                            if (T > 0.5f) med_contrib = pos;          // S/forward.cu:402-408: the last splat blended while T > 0.5
                            NA = ffma2s(make_float2(pn.x, pn.y), w, NA);
                            NC = ffma2s(make_float2(pn.z, pn.w), w, NC);
                            CC = ffma2s(make_float2(pc.x, pc.y), w, CC);
                            T = test_T;
                            last_contrib = pos;
                        }
                    }
                };
#if GSR_FWD_PAIR2
                // survivors two at a time: the two evaluations are independent instruction chains (each ends in a MUFU.EX2
                // behind a MUFU.RCP), interleaved they hide each other's latencies; the blends follow in list order
                while (m) {
                    const int b1 = __ffs(m) - 1;
                    m &= m - 1;
                    const bool two = m != 0u;
                    const int b2 = two ? __ffs(m) - 1 : b1;
                    m &= m - 1;          // (0 & anything stays 0)
                    const int j1 = c0 + b1, j2 = c0 + b2;
                    const PairEval e1 = eval_pair(sb[0][j1], sb[1][j1], sb[2][j1], sb[3][j1], fx, fy);
                    const PairEval e2 = eval_pair(sb[0][j2], sb[1][j2], sb[2][j2], sb[3][j2], fx, fy);
                    blend(e1, b1);
                    if (two) blend(e2, b2);
                }
#else
                while (m) {
                    const int bitpos = __ffs(m) - 1;
                    const int j = c0 + bitpos;
                    m &= m - 1;
                    const PairEval ev = eval_pair(sb[0][j], sb[1][j], sb[2][j], sb[3][j], fx, fy);
                    blend(ev, bitpos);
                }
#endif
                // checked once per 32-entry chunk: after the last pixel finishes, the rest of the chunk only evaluates
                if (__all_sync(FULLMASK, done)) warp_done = true;
                // hand the contributing entries to the backward: one predicated red.or per chunk into the record word
                if (MARK && ((used >> lane) & 1u))
                    atomicOr(reinterpret_cast<uint32_t*>(mark_plane + range_x + b * RBATCH + c0 + lane) + 3,
                             1u << (REC_USED_SHIFT + warp));
            }
        }
        // stage is free once every warp is past it; refill it with batch b+2
        const int all_done = __syncthreads_and(warp_done);
        if (all_done) {
            // drain the copy already in flight for batch b+1 before the CTA may exit
            if (threadIdx.x == 0 && b + 1 < nb) mbar_wait(&full_bar[(b + 1) & 1], (uint32_t)(((b + 1) >> 1) & 1));
            break;
        }
        if (threadIdx.x == 0 && b + 2 < nb) {
            fence_proxy_async();
            issue_batch(sbuf[stage], src, pstride, (b + 2) * RBATCH, min(RBATCH, n - (b + 2) * RBATCH), &full_bar[stage]);
        }
    }

    if (inside) {
        const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
        const size_t N = (size_t)W * H;
        const size_t pid = (size_t)py * W + px;
        // median splat: one record read and one re-evaluation per pixel (identical arithmetic, hence identical depth)
        // instead of six predicated moves per blended pair
        float med_depth = 0.f, mn0 = 0.f, mn1 = 0.f, mn2 = 0.f;
        int surf_idx = -1;
        if (med_contrib > 0) {
            const float4* rec = src + (med_contrib - 1);
            const float4 qa = __ldg(rec), qb = __ldg(rec + pstride), qc = __ldg(rec + 2 * pstride), qd = __ldg(rec + 3 * pstride),
                         pn = __ldg(rec + 4 * pstride);
            med_depth = eval_pair(qa, qb, qc, qd, fx, fy).depth;
            surf_idx = (int)(__float_as_uint(qd.w) & idx_mask);
            mn0 = pn.x; mn1 = pn.y; mn2 = pn.z;
        }
        final_T[pid] = T;
        const float C0 = NC.y, C1 = CC.x, C2 = CC.y, N0 = NA.x, N1 = NA.y, N2 = NC.x, Dacc = DM.x, M1 = DM.y;
        final_T[pid + N] = M1;
        final_T[pid + 2 * N] = M2;
        n_contrib[pid] = last_contrib;
        n_contrib[pid + N] = med_contrib;
        out_color[pid] = C0 + T * bg0;
        out_color[pid + N] = C1 + T * bg1;
        out_color[pid + 2 * N] = C2 + T * bg2;
        out_others[pid + 0 * N] = Dacc;
        out_others[pid + 1 * N] = 1.0f - T;
        out_others[pid + 2 * N] = N0;
        out_others[pid + 3 * N] = N1;
        out_others[pid + 4 * N] = N2;
        out_others[pid + 5 * N] = med_depth;
        out_others[pid + 6 * N] = dist;
        out_others[pid + 7 * N] = (float)surf_idx;
        out_others[pid + 8 * N] = mn0;
        out_others[pid + 9 * N] = mn1;
        out_others[pid + 10 * N] = mn2;
    }
}

template __global__ void surfel_render_fwd<false>(const uint32_t*, const float4*, size_t, int, int, int, const float*, float*, uint32_t*,
                                                 float*, float*, float4*);
template __global__ void surfel_render_fwd<true>(const uint32_t*, const float4*, size_t, int, int, int, const float*, float*, uint32_t*,
                                                float*, float*, float4*);
int fwd_ctas_per_tile() { return FWD_SPLIT; }

}  // namespace gsr
