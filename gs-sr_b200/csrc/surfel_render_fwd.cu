// surfel_render_fwd.cu -- per-tile front-to-back alpha blend of 2D Gaussian
// surfels (colour + 11 auxiliary channels) for sm_100a.
//
// Result contract = reference renderCUDA, S/cuda_rasterizer/forward.cu:256-448
// (ray-splat intersection :351-368, low-pass :362-366, alpha :381-389,
// distortion moments :395-400, median bookkeeping :402-408, outputs :427-447).
//
// B200 design (not the reference's):
//   * one CTA per 16x16 tile, 8 warps, each warp owns an 8x4 pixel block;
//   * the tile's sorted SplatRec stream is contiguous in HBM: one elected
//     thread pulls 128-record batches (10 KB) into a double-buffered shared
//     ring with cp.async.bulk + mbarrier (SASS UBLKCP), overlapping the next
//     batch's HBM/L2 latency with the current batch's blend;
//   * each warp culls the batch 32 records at a time: lane j tests record j's
//     conservative pixel bounds against the warp's 8x4 block, one ballot, and
//     only the surviving records are evaluated (the culled pairs provably
//     have alpha < 1/255, i.e. the reference would `continue` on them);
//   * records carry tile-local homography rows, so pixel coordinates are small
//     integers (better conditioned than the reference's global-pixel form);
//   * a warp stops as soon as its 32 pixels are done (the reference waits for
//     all 256), the CTA stops when all 8 warps have.
#include "common.cuh"
#include "async_copy.cuh"

namespace gsr {

constexpr int FWD_BATCH = 128;
constexpr uint32_t FULLMASK = 0xffffffffu;

__global__ void __launch_bounds__(TILE_PIX)
surfel_render_fwd(const uint2* __restrict__ ranges, const SplatRec* __restrict__ recs, int W, int H,
                  int gx, const float* __restrict__ bg, float* __restrict__ final_T,
                  uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                  float* __restrict__ out_others) {
    __shared__ __align__(128) SplatRec sbuf[2][FWD_BATCH];
    __shared__ __align__(8) uint64_t full_bar[2];

    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wx0 = (warp & 1) * 8, wy0 = (warp >> 1) * 4;
    const int lx = wx0 + (lane & 7), ly = wy0 + (lane >> 3);
    const int px = tx * TILE + lx, py = ty * TILE + ly;
    const bool inside = px < W && py < H;
    const float fx = (float)lx, fy = (float)ly;

    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    const int nb = (n + FWD_BATCH - 1) / FWD_BATCH;
    const SplatRec* src = recs + range.x;

    if (threadIdx.x == 0) {
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2 && b < nb; b++) {
            uint32_t bytes = (uint32_t)(min(FWD_BATCH, n - b * FWD_BATCH) * sizeof(SplatRec));
            mbar_expect_tx(&full_bar[b], bytes);
            bulk_g2s(&sbuf[b][0], src + b * FWD_BATCH, bytes, &full_bar[b]);
        }
    }

    float T = 1.0f;
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, N0 = 0.f, N1 = 0.f, N2 = 0.f;
    float Dacc = 0.f, M1 = 0.f, M2 = 0.f, dist = 0.f;
    float med_depth = 0.f, mn0 = 0.f, mn1 = 0.f, mn2 = 0.f;
    int surf_idx = -1;
    uint32_t last_contrib = 0, med_contrib = 0;
    bool done = !inside;
    bool warp_done = __all_sync(FULLMASK, done);
    const float MSCALE = FAR_N / (FAR_N - NEAR_N);

    for (int b = 0; b < nb; b++) {
        const int stage = b & 1;
        const uint32_t parity = (uint32_t)((b >> 1) & 1);
        const int cnt = min(FWD_BATCH, n - b * FWD_BATCH);
        if (!warp_done) {
            mbar_wait(&full_bar[stage], parity);
            const SplatRec* sb = sbuf[stage];
            for (int c0 = 0; c0 < cnt && !warp_done; c0 += 32) {
                const int e = c0 + lane;
                uint32_t bits = BOUNDS_EMPTY;
                if (e < cnt) bits = __float_as_uint(sb[e].cb.w);
                const int bx0 = bits & 15, bx1 = (bits >> 4) & 15, by0 = (bits >> 8) & 15, by1 = (bits >> 12) & 15;
                const bool hit = !(bits & BOUNDS_EMPTY) && bx0 <= wx0 + 7 && bx1 >= wx0 && by0 <= wy0 + 3 && by1 >= wy0;
                uint32_t m = __ballot_sync(FULLMASK, hit);
                while (m) {
                    const int j = c0 + __ffs(m) - 1;
                    m &= m - 1;
                    const float4* r = reinterpret_cast<const float4*>(sb + j);
                    const float4 tu = r[0], tv = r[1], tw = r[2];
                    bool valid = !done;
                    // ray-splat intersection in tile-local pixel coordinates
                    const float kx = fmaf(fx, tw.x, -tu.x), ky = fmaf(fx, tw.y, -tu.y), kz = fmaf(fx, tw.z, -tu.z);
                    const float l0 = fmaf(fy, tw.x, -tv.x), l1 = fmaf(fy, tw.y, -tv.y), l2 = fmaf(fy, tw.z, -tv.z);
                    const float p0 = ky * l2 - kz * l1, p1 = kz * l0 - kx * l2, p2 = kx * l1 - ky * l0;
                    valid = valid && (p2 != 0.0f);
                    const float ip = __frcp_rn(p2);
                    const float s0 = p0 * ip, s1 = p1 * ip;
                    const float rho3d = s0 * s0 + s1 * s1;
                    const float d0 = tu.w - fx, d1 = tv.w - fy;
                    const float rho2d = FILTER_INV_SQUARE * (d0 * d0 + d1 * d1);
                    const float rho = fminf(rho3d, rho2d);
                    const float depth = (rho3d <= rho2d) ? (s0 * tw.x + s1 * tw.y) + tw.z : tw.z;
                    valid = valid && !(depth < NEAR_N);
                    const float power = -0.5f * rho;
                    valid = valid && !(power > 0.0f);
                    const float alpha = fminf(ALPHA_MAX, tw.w * __expf(power));
                    valid = valid && !(alpha < ALPHA_MIN);
                    if (__any_sync(FULLMASK, valid)) {
                        const float test_T = T * (1.0f - alpha);
                        if (valid && test_T < T_EPS) { done = true; valid = false; }
                        if (valid) {
                            const float4 ng = r[3], cb = r[4];
                            const float w = alpha * T;
                            const float A = 1.0f - T;
                            const float mm = MSCALE * (1.0f - NEAR_N * __frcp_rn(depth));
                            dist += (mm * mm * A + M2 - 2.0f * mm * M1) * w;
                            Dacc += depth * w;
                            M1 += mm * w;
                            M2 += mm * mm * w;
                            const uint32_t pos = (uint32_t)(b * FWD_BATCH + j + 1);
                            if (T > 0.5f) {
                                med_depth = depth;
                                surf_idx = (int)__float_as_uint(ng.w);
                                mn0 = ng.x; mn1 = ng.y; mn2 = ng.z;
                                med_contrib = pos;
                            }
                            N0 += ng.x * w; N1 += ng.y * w; N2 += ng.z * w;
                            C0 += cb.x * w; C1 += cb.y * w; C2 += cb.z * w;
                            T = test_T;
                            last_contrib = pos;
                        }
                        if (__all_sync(FULLMASK, done)) { warp_done = true; break; }
                    }
                }
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
            uint32_t bytes = (uint32_t)(min(FWD_BATCH, n - (b + 2) * FWD_BATCH) * sizeof(SplatRec));
            fence_proxy_async();
            mbar_expect_tx(&full_bar[stage], bytes);
            bulk_g2s(&sbuf[stage][0], src + (b + 2) * FWD_BATCH, bytes, &full_bar[stage]);
        }
    }

    if (inside) {
        const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
        const size_t N = (size_t)W * H;
        const size_t pid = (size_t)py * W + px;
        final_T[pid] = T;
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

}  // namespace gsr
