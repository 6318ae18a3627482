// surfel_render_bwd.cu -- per-tile back-to-front gradient walk of the 2DGS surfel blend
// for sm_100a.
//
// Result contract = reference backward renderCUDA, S/cuda_rasterizer/backward.cu:143-447
// (per-pair maths :287-444 incl. the distortion terms :347-364, the median-depth term
// :349-352, the median-normal quirk :381, the background term :391-394, the ray-splat
// branch :403-433 and the low-pass branch :434-441).
//
// B200 design:
//   * same warp/pixel mapping (8x4 pixels per warp), cp.async.bulk staging and exact warp-block
//     culling as the forward, but one CTA = 4 warps = HALF a tile (16x8 pixels; two CTAs per
//     tile walk the same list) and the warps of a CTA are decoupled: a 3-stage ring whose
//     stages are refilled by whichever warp arrives last, no __syncthreads() in the walk;
//     the walk starts at the half tile's last needed batch (max last_contributor);
//   * geometry gradients are accumulated as MOMENTS of dL/dp (p = a x + b y + c, the
//     adjugate-form intersection): M0 = sum dp, MX = sum x~ dp, MY = sum y~ dp with (x~, y~)
//     measured from the splat's rounded screen centre, plus dL/d det(T).  The cross products
//     that turn them into dL/dT (S/backward.cu:413-421) are linear, so they are applied ONCE
//     per Gaussian in the backward preprocess instead of once per (pixel, splat) pair;
//   * per-splat sums over the warp's 32 pixels are formed in registers with a transposed
//     shuffle reduction (16 values -> 16 shuffles) and flushed with ONE 16-lane
//     red.global.add.f32 burst into the Gaussian's 80-byte accumulator; the reference issues
//     19 global float atomics per (pixel, splat) pair.
#include "common.cuh"
#include "async_copy.cuh"
#include "cull.cuh"
#include "render_common.cuh"

namespace gsr {

// Transposed warp reduction of 16 per-lane values: afterwards lane L holds the 32-lane sum of
// v[idx], idx = ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1) (lanes 2k, 2k+1 both).
__device__ __forceinline__ float warp_reduce16_transposed(float (&v)[16], int lane) {
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

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULLMASK, x, o);
    return x;
}

#ifndef GSR_BWD_BATCH
#define GSR_BWD_BATCH 128
#define GSR_BWD_STAGES 3
#endif
constexpr int BWD_BATCH = GSR_BWD_BATCH;      // record entries per ring stage
constexpr int BWD_STAGES = GSR_BWD_STAGES;
#ifndef GSR_BWD_WARPS
#define GSR_BWD_WARPS 4   // measured on cfg-B (B200): 8 warps 3.18 ms, 4 warps 3.08 ms, 2 warps 3.23 ms
#endif
constexpr int BWD_WARPS = GSR_BWD_WARPS;      // warps per CTA: 8 = whole 16x16 tile, 4 = half tile (16x8), 2 = 16x4 strip
constexpr int BWD_SPLIT = 8 / BWD_WARPS;      // CTAs per tile
constexpr int BWD_MINB = BWD_WARPS == 8 ? 3 : (BWD_WARPS == 4 ? 6 : 12);

// USED: the record word carries the forward's per-warp-block "blended" bits (P < 2^23), no cull test here
template <bool USED>
__global__ void __launch_bounds__(BWD_WARPS * 32, BWD_MINB)
surfel_render_bwd(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ planes, size_t pstride, int W,
                  int H, int gx, const float* __restrict__ bg, const float* __restrict__ final_T,
                  const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                  const float* __restrict__ dL_dothers, float* __restrict__ gacc) {
    // BWD_STAGES-deep ring of record batches.  Warps are NOT synchronised per batch: every warp waits on the
    // stage's full barrier, consumes (or skips) the batch and then arrives on the stage's counter; the warp whose
    // arrival is the 8th re-arms the barrier and issues the bulk copy of the batch BWD_STAGES ahead into the freed
    // stage.  A fast warp can therefore run up to BWD_STAGES batches ahead of the slowest one instead of idling at
    // a CTA barrier after every 128 entries (barrier stalls were 25 % of all warp time with __syncthreads()).
    __shared__ __align__(128) float4 sbuf[BWD_STAGES][REC_PLANES][BWD_BATCH];
    __shared__ __align__(8) uint64_t full_bar[BWD_STAGES];
    __shared__ int s_arrive[BWD_STAGES];
    __shared__ int s_maxlast;

    const int tile = blockIdx.x / BWD_SPLIT;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = (threadIdx.x >> 5) + (blockIdx.x % BWD_SPLIT) * BWD_WARPS, lane = threadIdx.x & 31;
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

    const int last = inside ? (int)n_contrib[pid] : 0;  // entries [0, last) contribute
    const int medpos = inside ? (int)n_contrib[pid + N] - 1 : -1;
    const int wlast = __reduce_max_sync(FULLMASK, last);
    if (threadIdx.x == 0) {
        s_maxlast = 0;
#pragma unroll
        for (int i = 0; i < BWD_STAGES; i++) { mbar_init(&full_bar[i], 1); s_arrive[i] = 0; }
        mbar_fence_init();
    }
    __syncthreads();
    if (lane == 0 && wlast > 0) atomicMax(&s_maxlast, wlast);
    __syncthreads();
    const int maxlast = min(s_maxlast, n);
    if (maxlast <= 0) return;
    const int nb = (maxlast + BWD_BATCH - 1) / BWD_BATCH;  // batches [0, nb), walked from nb-1 down

    auto batch_count = [&](int b) { return min(BWD_BATCH, n - b * BWD_BATCH); };
    if (threadIdx.x == 0)
        for (int i = 0; i < BWD_STAGES && nb - 1 - i >= 0; i++) {
            const int b = nb - 1 - i;
            issue_batch(sbuf[i], src, pstride, b * BWD_BATCH, batch_count(b), &full_bar[i]);
        }

    // per-pixel constants
    const float T_final = inside ? final_T[pid] : 0.f;
    const float final_D = inside ? final_T[pid + N] : 0.f;
    const float final_D2 = inside ? final_T[pid + 2 * N] : 0.f;
    const float final_A = 1.0f - T_final;
    float dpx0 = 0.f, dpx1 = 0.f, dpx2 = 0.f, dL_ddepth = 0.f, dL_daccum = 0.f, dL_dreg = 0.f;
    float dn0 = 0.f, dn1 = 0.f, dn2 = 0.f, dL_dmedian_depth = 0.f, dmn0 = 0.f, dmn1 = 0.f, dmn2 = 0.f;
    if (inside) {
        dpx0 = dL_dpix[pid]; dpx1 = dL_dpix[pid + N]; dpx2 = dL_dpix[pid + 2 * N];
        dL_ddepth = dL_dothers[pid + 0 * N];
        dL_daccum = dL_dothers[pid + 1 * N];
        dn0 = dL_dothers[pid + 2 * N]; dn1 = dL_dothers[pid + 3 * N]; dn2 = dL_dothers[pid + 4 * N];
        dL_dmedian_depth = dL_dothers[pid + 5 * N];
        dL_dreg = dL_dothers[pid + 6 * N];
        dmn0 = dL_dothers[pid + 8 * N]; dmn1 = dL_dothers[pid + 9 * N]; dmn2 = dL_dothers[pid + 10 * N];
    }
    const float bg_dot_dpixel = __ldg(bg) * dpx0 + __ldg(bg + 1) * dpx1 + __ldg(bg + 2) * dpx2;
    const float ox = (float)(tx * TILE), oy = (float)(ty * TILE);
    (void)ox; (void)oy;

    // running state of the reverse walk
    float T = T_final;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ar0 = 0.f, ar1 = 0.f, ar2 = 0.f;
    float last_depth = 0.f, accum_depth_rec = 0.f, accum_alpha_rec = 0.f;
    float ln0 = 0.f, ln1 = 0.f, ln2 = 0.f, an0 = 0.f, an1 = 0.f, an2 = 0.f, last_dL_dT = 0.f;

    int stage = 0;
    uint32_t parity = 0;
    for (int it = 0; it < nb; it++) {
        const int b = nb - 1 - it;
        const int cnt = batch_count(b);
        const int base = b * BWD_BATCH;
        mbar_wait(&full_bar[stage], parity);
        if (base < wlast) {
            const float4(*sb)[BWD_BATCH] = sbuf[stage];
            for (int c0 = ((cnt - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
                if (base + c0 >= wlast) continue;
                const int e = c0 + lane;
                bool hit = false;
                if (e < cnt && base + e < wlast) {
                    if (USED) hit = (__float_as_uint(sb[3][e].w) >> (REC_USED_SHIFT + warp)) & 1u;
                    else hit = entry_hits_block(sb[0][e], sb[1][e], sb[2][e], sb[3][e], bx0, bx1, by0, by1);
                }
                uint32_t m = __ballot_sync(FULLMASK, hit);
                while (m) {
                    const int bit = 31 - __clz(m);
                    m &= ~(1u << bit);
                    const int j = c0 + bit;
                    const int pos = base + j;  // 0-based list position == reference `contributor`
                    const float4 qa = sb[0][j], qb = sb[1][j], qc = sb[2][j], qd = sb[3][j];
                    const PairEval ev = eval_pair(qa, qb, qc, qd, fx, fy);
                    const bool valid = ev.valid && pos < last;
                    if (!__any_sync(FULLMASK, valid)) continue;

                    const float4 pn = sb[4][j], pc = sb[5][j];
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = 0.f;
                    float vo = 0.f, gm0 = 0.f, gm1 = 0.f, gz = 0.f;
                    bool lowpass = false;
                    if (valid) {
                        const float alpha = ev.alpha, G = ev.G, c_d = ev.depth;
                        const float ria = fast_rcp(1.0f - alpha);
                        T = T * ria;
                        const float w = alpha * T;
                        const float omla = 1.f - last_alpha;
                        // colour
                        ar0 = last_alpha * lc0 + omla * ar0; lc0 = pn.w;
                        ar1 = last_alpha * lc1 + omla * ar1; lc1 = pc.x;
                        ar2 = last_alpha * lc2 + omla * ar2; lc2 = pc.y;
                        float dL_dalpha = (pn.w - ar0) * dpx0 + (pc.x - ar1) * dpx1 + (pc.y - ar2) * dpx2;
                        v[10] = w * dpx0; v[11] = w * dpx1; v[12] = w * dpx2;
                        // depth distortion + median depth
                        float dL_dz = (pos == medpos) ? dL_dmedian_depth : 0.f;
                        const float icd = fast_rcp(c_d);
                        const float m_d = MSCALE * (1.0f - NEAR_N * icd);
                        const float dmd_dd = DMD * icd * icd;
                        const float dL_dweight = (final_D2 + m_d * m_d * final_A - 2.0f * m_d * final_D) * dL_dreg;
                        dL_dalpha += dL_dweight - last_dL_dT;
                        last_dL_dT = dL_dweight * alpha + (1.0f - alpha) * last_dL_dT;
                        dL_dz += 2.0f * w * (m_d * final_A - final_D) * dL_dreg * dmd_dd;
                        // expected depth, alpha
                        accum_depth_rec = last_alpha * last_depth + omla * accum_depth_rec;
                        last_depth = c_d;
                        dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
                        accum_alpha_rec = last_alpha + omla * accum_alpha_rec;
                        dL_dalpha += (1.0f - accum_alpha_rec) * dL_daccum;
                        // normals (incl. the every-splat median-normal term)
                        an0 = last_alpha * ln0 + omla * an0; ln0 = pn.x;
                        an1 = last_alpha * ln1 + omla * an1; ln1 = pn.y;
                        an2 = last_alpha * ln2 + omla * an2; ln2 = pn.z;
                        dL_dalpha += (pn.x - an0) * dn0 + (pn.y - an1) * dn1 + (pn.z - an2) * dn2;
                        v[13] = w * dn0 + dmn0; v[14] = w * dn1 + dmn1; v[15] = w * dn2 + dmn2;

                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final * ria) * bg_dot_dpixel;
                        const float dL_dG = qc.w * dL_dalpha;
                        dL_dz += w * dL_ddepth;
                        vo = G * dL_dalpha;
                        if (ev.ray) {
                            // dL/ds = dL_dG * (-G) * s ;  s = p.xy / p.z ; depth = det(T) / p.z
                            const float gs = dL_dG * -G * ev.ip;
                            const float a0 = gs * ev.s0, a1 = gs * ev.s1;
                            const float a2 = -(a0 * ev.s0 + a1 * ev.s1) - dL_dz * c_d * ev.ip;
                            const float xs = fx - pc.z, ys = fy - pc.w;   // measured from the splat's rounded centre
                            v[0] = a0; v[1] = a1; v[2] = a2;
                            v[3] = xs * a0; v[4] = xs * a1; v[5] = xs * a2;
                            v[6] = ys * a0; v[7] = ys * a1; v[8] = ys * a2;
                            v[9] = dL_dz * ev.ip;
                        } else {
                            gm0 = dL_dG * (-G * FILTER_INV_SQUARE * ev.d0);
                            gm1 = dL_dG * (-G * FILTER_INV_SQUARE * ev.d1);
                            gz = dL_dz;
                            lowpass = true;
                        }
                    }
                    const bool any_lowpass = __any_sync(FULLMASK, lowpass);
                    const float red = warp_reduce16_transposed(v, lane);
                    const float so = warp_sum(vo);
                    const uint32_t g = __float_as_uint(qd.w) & (USED ? REC_INDEX_MASK_USED : ~REC_FLAG_ALWAYS);
                    float* acc = gacc + (size_t)g * GACC_STRIDE;
                    const int vi = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    if ((lane & 1) == 0) atomicAdd(acc + vi, red);
                    if (lane == 1) atomicAdd(acc + 16, so);
                    if (any_lowpass) {
                        const float r0 = warp_sum(gm0), r1 = warp_sum(gm1), r2 = warp_sum(gz);
                        if (lane == 0) { atomicAdd(acc + 17, r2); atomicAdd(acc + 18, r0); atomicAdd(acc + 19, r1); }
                    }
                }
            }
        }
        // release the stage; the last of the 8 warps to arrive refills it with the batch BWD_STAGES ahead
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            const int old = atomicAdd(&s_arrive[stage], 1);
            if ((old & (BWD_WARPS - 1)) == BWD_WARPS - 1 && it + BWD_STAGES < nb) {
                __threadfence_block();
                fence_proxy_async();
                const int b2 = nb - 1 - (it + BWD_STAGES);
                issue_batch(sbuf[stage], src, pstride, b2 * BWD_BATCH, batch_count(b2), &full_bar[stage]);
            }
        }
        if (++stage == BWD_STAGES) { stage = 0; parity ^= 1u; }
    }
}

template __global__ void surfel_render_bwd<false>(const uint32_t*, const float4*, size_t, int, int, int, const float*, const float*,
                                                 const uint32_t*, const float*, const float*, float*);
template __global__ void surfel_render_bwd<true>(const uint32_t*, const float4*, size_t, int, int, int, const float*, const float*,
                                                const uint32_t*, const float*, const float*, float*);
int bwd_ctas_per_tile() { return BWD_SPLIT; }

}  // namespace gsr
