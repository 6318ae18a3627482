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
//     tile walk the same list) and the warps of a CTA are decoupled: a ring whose stages are
//     refilled by whichever warp arrives last, no __syncthreads() in the walk; the walk starts
//     at the half tile's last needed batch (max last_contributor);
//   * TWO PHASES per warp.  Phase 1 is pixel-parallel (lane = pixel) and sequential over the
//     contributing splats, because the blend state (T and the "blended behind" recurrence) is a
//     chain along the list: it turns one (pixel, splat) pair into six numbers -- dL/dp (3),
//     dL/d det(T), the blend weight w and G dL/dalpha -- and parks them in a per-warp
//     shared-memory slot.  Phase 2 runs whenever the slots are full and is SPLAT-parallel
//     (lane = pending splat x a run of the block's pixels): every lane streams its slot's
//     numbers and accumulates the 17 per-splat sums in registers, so the 32-pixel reduction costs
//     one FADD/FFMA per value instead of a shuffle network, and the lanes that idle in phase 1
//     (a splat covers ~12 of a block's 32 pixels) cost nothing here;
//   * all linear per-pixel channels (colour, depth, normal, distortion weight) share ONE
//     recurrence: with S_i = <channel values of splat i, upstream grads of the pixel> the
//     reference's "accum_rec" chains (S/backward.cu:317-388) collapse to rec <- rec + alpha (S - rec),
//     and the alpha-channel term (1 - accum_alpha_rec) T_i equals T_final / (1 - alpha_i), so it folds
//     into the background term;
//   * geometry gradients are accumulated as MOMENTS of dL/dp (p = a x + b y + c, the
//     adjugate-form intersection): M0 = sum dp, MX = sum x~ dp, MY = sum y~ dp with (x~, y~)
//     measured from the splat's rounded screen centre, plus dL/d det(T).  The cross products
//     that turn them into dL/dT (S/backward.cu:413-421) are linear, so they are applied ONCE
//     per Gaussian in the backward preprocess instead of once per (pixel, splat) pair;
//   * a pending splat is flushed with four 128-bit red.global.add.v4.f32 + one scalar reduction
//     into the Gaussian's 80-byte accumulator; the reference issues 19 global float atomics per
//     (pixel, splat) pair.
#include "common.cuh"
#include "async_copy.cuh"
#include "cull.cuh"
#include "render_common.cuh"

namespace gsr {

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULLMASK, x, o);
    return x;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

#ifndef GSR_BWD_BATCH
#define GSR_BWD_BATCH 32
#endif
#ifndef GSR_BWD_STAGES
#define GSR_BWD_STAGES 5
#endif
#ifndef GSR_BWD_WARPS
#define GSR_BWD_WARPS 8
#endif
#ifndef GSR_BWD_SLOTS
#define GSR_BWD_SLOTS 8
#endif
constexpr int BWD_BATCH = GSR_BWD_BATCH;      // record entries per ring stage
constexpr int BWD_STAGES = GSR_BWD_STAGES;
constexpr int BWD_WARPS = GSR_BWD_WARPS;      // warps per CTA: 8 = whole 16x16 tile, 4 = half tile (16x8), 2 = 16x4 strip
constexpr int BWD_SPLIT = 8 / BWD_WARPS;      // CTAs per tile
constexpr int BWD_SLOTS = GSR_BWD_SLOTS;      // pending (warp block, splat) pairs per warp between two phase-2 runs
constexpr int BWD_LPS = 32 / BWD_SLOTS;       // phase 2: lanes per pending splat
constexpr int BWD_PPL = 32 / BWD_LPS;         // phase 2: pixels per lane (a run of whole block rows)
constexpr int PAIR_VALS = 6;                  // a0 a1 | a2 q | w v
constexpr int SLOT_STRIDE = 32 * PAIR_VALS + 2;   // words; (stride / 2) odd keeps the 64-bit slot reads conflict-free
static_assert(BWD_SLOTS == 8 || BWD_SLOTS == 16 || BWD_SLOTS == 32, "slots per warp");
static_assert(BWD_WARPS == 2 || BWD_WARPS == 4 || BWD_WARPS == 8, "warps per CTA");

// dynamic shared memory layout (bytes)
constexpr size_t SM_RING = (size_t)BWD_STAGES * REC_PLANES * BWD_BATCH * 16;
constexpr size_t SM_PEND = (size_t)BWD_WARPS * BWD_SLOTS * SLOT_STRIDE * 4;
constexpr size_t SM_HDR = (size_t)BWD_WARPS * BWD_SLOTS * 16;
// per pixel: (dpx0 dpx1 dpx2 dn0) (dn1 dn2 - -); the median-normal gradients (quirk Q1, never set by GS-SR) are re-read from
// global memory in the rare path that needs them.  The pixels of one phase-2 lane group are contiguous and
// the groups are spaced an odd number of float4 apart: the BWD_LPS addresses one phase-2 load touches then fall into
// different banks (with a multiple of 128 B between them every such load was a BWD_LPS-way conflict: 8 wavefronts / pair).
constexpr int PIXC_GROUP = BWD_PPL * 2 + 1;
constexpr size_t SM_PIXC = (size_t)BWD_WARPS * BWD_LPS * PIXC_GROUP * 16;
constexpr size_t SM_MISC = 128;
constexpr size_t BWD_SMEM = SM_RING + SM_PEND + SM_HDR + SM_PIXC + SM_MISC;
constexpr int BWD_MAXB = (int)((227 * 1024) / (BWD_SMEM + 1024));
#ifndef GSR_BWD_MINB
#define GSR_BWD_MINB (BWD_MAXB < 1 ? 1 : (BWD_MAXB > 16 ? 16 : BWD_MAXB))
#endif
constexpr int BWD_MINB = GSR_BWD_MINB;

// USED: the record word carries the forward's per-warp-block "blended" bits (P < 2^23), no cull test here
template <bool USED>
__global__ void __launch_bounds__(BWD_WARPS * 32, BWD_MINB)
surfel_render_bwd(const uint32_t* __restrict__ tile_offset, const float4* __restrict__ planes, size_t pstride, int W,
                  int H, int gx, const float* __restrict__ bg, const float* __restrict__ final_T,
                  const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                  const float* __restrict__ dL_dothers, float* __restrict__ gacc, int sync_ring) {
    // BWD_STAGES-deep ring of record batches.  Warps are NOT synchronised per batch: every warp waits on the
    // stage's full barrier, consumes (or skips) the batch and then arrives on the stage's counter; the warp whose
    // arrival is the last re-arms the barrier and issues the bulk copy of the batch BWD_STAGES ahead into the freed
    // stage.  A fast warp can therefore run up to BWD_STAGES batches ahead of the slowest one instead of idling at
    // a CTA barrier after every batch (barrier stalls were 25 % of all warp time with __syncthreads()).
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4(*sbuf)[REC_PLANES][BWD_BATCH] = reinterpret_cast<float4(*)[REC_PLANES][BWD_BATCH]>(smem_raw);
    float* pend_all = reinterpret_cast<float*>(smem_raw + SM_RING);
    float4* hdr_all = reinterpret_cast<float4*>(smem_raw + SM_RING + SM_PEND);
    float4* pixc_all = reinterpret_cast<float4*>(smem_raw + SM_RING + SM_PEND + SM_HDR);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + SM_RING + SM_PEND + SM_HDR + SM_PIXC);
    int* s_arrive = reinterpret_cast<int*>(full_bar + 8);
    int* s_maxlast = s_arrive + 8;

    const int tile = blockIdx.x / BWD_SPLIT;
    const int tx = tile % gx, ty = tile / gx;
    const int wic = threadIdx.x >> 5;                                   // warp in CTA
    const int warp = wic + (blockIdx.x % BWD_SPLIT) * BWD_WARPS, lane = threadIdx.x & 31;   // warp block in tile
    const int wx0 = (warp & 1) * 8, wy0 = (warp >> 1) * 4;
    const int lx = wx0 + (lane & 7), ly = wy0 + (lane >> 3);
    const int px = tx * TILE + lx, py = ty * TILE + ly;
    const bool inside = px < W && py < H;
    const float fx = (float)lx, fy = (float)ly;
    const float bx0 = (float)wx0 - CULL_MARGIN, bx1 = (float)(wx0 + 7) + CULL_MARGIN;
    const float by0 = (float)wy0 - CULL_MARGIN, by1 = (float)(wy0 + 3) + CULL_MARGIN;
    const size_t N = (size_t)W * H;
    const size_t pid = (size_t)py * W + px;
    float* pend = pend_all + (size_t)wic * BWD_SLOTS * SLOT_STRIDE;
    float4* hdr = hdr_all + wic * BWD_SLOTS;
    float4* pixc = pixc_all + wic * BWD_LPS * PIXC_GROUP;

    const uint32_t range_x = tile_offset[tile];
    const int n = (int)(tile_offset[tile + 1] - range_x);
    const float4* src = planes + range_x;

    const int last = inside ? (int)n_contrib[pid] : 0;  // entries [0, last) contribute
    const int medpos = inside ? (int)n_contrib[pid + N] - 1 : -1;
    const int wlast = __reduce_max_sync(FULLMASK, last);
    if (threadIdx.x == 0) {
        *s_maxlast = 0;
#pragma unroll
        for (int i = 0; i < BWD_STAGES; i++) { mbar_init(&full_bar[i], 1); s_arrive[i] = 0; }
        mbar_fence_init();
    }
    __syncthreads();
    if (lane == 0 && wlast > 0) atomicMax(s_maxlast, wlast);
    __syncthreads();
    const int maxlast = min(*s_maxlast, n);
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
    // alpha channel + background (S/backward.cu:372-376,391-394): (1 - accum_alpha_rec) T_i == T_final / (1 - alpha_i)
    const float k_tail = T_final * (dL_daccum - (__ldg(bg) * dpx0 + __ldg(bg + 1) * dpx1 + __ldg(bg + 2) * dpx2));
    const float m2fD = -2.0f * final_D;
    const float reg2 = 2.0f * dL_dreg;
    // phase 2 reads the upstream gradients of all 32 pixels of the block
    {
        float4* mine = pixc + (lane / BWD_PPL) * PIXC_GROUP + (lane % BWD_PPL) * 2;
        mine[0] = make_float4(dpx0, dpx1, dpx2, dn0);
        mine[1] = make_float4(dn1, dn2, 0.f, 0.f);
    }
    const bool has_dmn = __any_sync(FULLMASK, dmn0 != 0.f || dmn1 != 0.f || dmn2 != 0.f);   // quirk Q1; never set by GS-SR
    __syncwarp();

    // ---- phase 2: splat-parallel accumulation of the pending pairs and flush into gacc ----
    const int ps = lane % BWD_SLOTS, ph = lane / BWD_SLOTS;
    auto flush = [&](int np) {
        __syncwarp();
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, x0 = 0.f, x1 = 0.f, x2 = 0.f, y0 = 0.f, y1 = 0.f, y2 = 0.f, qs = 0.f;
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f, so = 0.f;
        if (ps < np) {
            const float* pb = pend + ps * SLOT_STRIDE + ph * (BWD_PPL * PAIR_VALS);
            const float4* pk = pixc + ph * PIXC_GROUP;
#pragma unroll
            for (int i = 0; i < BWD_PPL; i++) {
                const float2 u0 = *reinterpret_cast<const float2*>(pb + i * PAIR_VALS);
                const float2 u1 = *reinterpret_cast<const float2*>(pb + i * PAIR_VALS + 2);
                const float2 u2 = *reinterpret_cast<const float2*>(pb + i * PAIR_VALS + 4);
                const float4 k0 = pk[i * 2], k1 = pk[i * 2 + 1];
                m0 += u0.x; m1 += u0.y; m2 += u1.x;
                if ((i & 7) != 0) {
                    const float xi = (float)(i & 7);
                    x0 = fmaf(xi, u0.x, x0); x1 = fmaf(xi, u0.y, x1); x2 = fmaf(xi, u1.x, x2);
                }
                if ((i >> 3) != 0) {
                    const float yi = (float)(i >> 3);
                    y0 = fmaf(yi, u0.x, y0); y1 = fmaf(yi, u0.y, y1); y2 = fmaf(yi, u1.x, y2);
                }
                qs += u1.y;
                c0 = fmaf(u2.x, k0.x, c0); c1 = fmaf(u2.x, k0.y, c1); c2 = fmaf(u2.x, k0.z, c2);
                n0 = fmaf(u2.x, k0.w, n0); n1 = fmaf(u2.x, k1.x, n1); n2 = fmaf(u2.x, k1.y, n2);
                so += u2.y;
            }
            if (has_dmn) {
                // S/backward.cu:381 adds dL/dmedian_normal to EVERY contributing splat; pixel i of this lane's run:
                const int p0 = ph * BWD_PPL;
#pragma unroll 4
                for (int i = 0; i < BWD_PPL; i++)
                    if (pb[i * PAIR_VALS + 4] > 0.f) {       // w > 0 <=> the pair contributed
                        const int qx = tx * TILE + wx0 + ((p0 + i) & 7), qy = ty * TILE + wy0 + ((p0 + i) >> 3);
                        const size_t q = (size_t)qy * W + qx;
                        n0 += __ldg(dL_dothers + q + 8 * N); n1 += __ldg(dL_dothers + q + 9 * N); n2 += __ldg(dL_dothers + q + 10 * N);
                    }
            }
            // block-local pixel offsets -> offsets from the splat's moment origin (hdr.y, hdr.z; tile-local)
            const float4 h4 = hdr[ps];
            const float ox = (float)wx0 - h4.y, oy = (float)(wy0 + ph * (BWD_PPL / 8)) - h4.z;
            x0 = fmaf(ox, m0, x0); x1 = fmaf(ox, m1, x1); x2 = fmaf(ox, m2, x2);
            y0 = fmaf(oy, m0, y0); y1 = fmaf(oy, m1, y1); y2 = fmaf(oy, m2, y2);
        }
#pragma unroll
        for (int o = BWD_SLOTS; o < 32; o <<= 1) {
            m0 += __shfl_xor_sync(FULLMASK, m0, o); m1 += __shfl_xor_sync(FULLMASK, m1, o); m2 += __shfl_xor_sync(FULLMASK, m2, o);
            x0 += __shfl_xor_sync(FULLMASK, x0, o); x1 += __shfl_xor_sync(FULLMASK, x1, o); x2 += __shfl_xor_sync(FULLMASK, x2, o);
            y0 += __shfl_xor_sync(FULLMASK, y0, o); y1 += __shfl_xor_sync(FULLMASK, y1, o); y2 += __shfl_xor_sync(FULLMASK, y2, o);
            qs += __shfl_xor_sync(FULLMASK, qs, o);
            c0 += __shfl_xor_sync(FULLMASK, c0, o); c1 += __shfl_xor_sync(FULLMASK, c1, o); c2 += __shfl_xor_sync(FULLMASK, c2, o);
            n0 += __shfl_xor_sync(FULLMASK, n0, o); n1 += __shfl_xor_sync(FULLMASK, n1, o); n2 += __shfl_xor_sync(FULLMASK, n2, o);
            so += __shfl_xor_sync(FULLMASK, so, o);
        }
        if (ph == 0 && ps < np) {
            float* acc = gacc + (size_t)__float_as_uint(hdr[ps].x) * GACC_STRIDE;
            red_add_v4(acc + 0, m0, m1, m2, x0);
            red_add_v4(acc + 4, x1, x2, y0, y1);
            red_add_v4(acc + 8, y2, qs, c0, c1);
            red_add_v4(acc + 12, c2, n0, n1, n2);
            atomicAdd(acc + 16, so);
        }
        __syncwarp();
    };

    // running state of the reverse walk (phase 1)
    float T = T_final, rec = 0.f;
    int npend = 0;

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
                    float2 o0 = make_float2(0.f, 0.f), o1 = o0, o2 = o0;
                    float gm0 = 0.f, gm1 = 0.f, gz = 0.f;
                    bool lowpass = false;
                    if (valid) {
                        const float alpha = ev.alpha, G = ev.G, c_d = ev.depth;
                        const float ria = fast_rcp(1.0f - alpha);
                        T = T * ria;                                   // transmittance in front of this splat
                        const float w = alpha * T;
                        // distortion (S/backward.cu:347-364)
                        const float icd = fast_rcp(c_d);
                        const float m_d = fmaf(-MSCALE * NEAR_N, icd, MSCALE);
                        const float dmd_dd = DMD * icd * icd;
                        const float dL_dweight = fmaf(m_d, fmaf(m_d, final_A, m2fD), final_D2) * dL_dreg;
                        // S = <this splat's channel values, the pixel's upstream gradients>; rec = the same blended over the
                        // splats behind it (one chain instead of the reference's accum_rec per channel)
                        float S = fmaf(pn.w, dpx0, dL_dweight);
                        S = fmaf(pc.x, dpx1, S); S = fmaf(pc.y, dpx2, S);
                        S = fmaf(c_d, dL_ddepth, S);
                        S = fmaf(pn.x, dn0, S); S = fmaf(pn.y, dn1, S); S = fmaf(pn.z, dn2, S);
                        const float D = S - rec;
                        rec = fmaf(alpha, D, rec);
                        const float dL_dalpha = fmaf(D, T, k_tail * ria);
                        float dL_dz = fmaf(fmaf(m_d, final_A, -final_D) * reg2, dmd_dd, dL_ddepth) * w;
                        if (pos == medpos) dL_dz += dL_dmedian_depth;
                        const float v = G * dL_dalpha;                 // dL/dopacity share
                        const float dL_dG = qc.w * dL_dalpha;
                        o2 = make_float2(w, v);
                        if (ev.ray) {
                            // dL/ds = dL_dG * (-G) * s ;  s = p.xy / p.z ; depth = det(T) / p.z
                            const float gs = dL_dG * -G * ev.ip;
                            const float a0 = gs * ev.s0, a1 = gs * ev.s1;
                            const float q = dL_dz * ev.ip;
                            const float a2 = -fmaf(a0, ev.s0, fmaf(a1, ev.s1, q * c_d));
                            o0 = make_float2(a0, a1);
                            o1 = make_float2(a2, q);
                        } else {
                            gm0 = dL_dG * (-G * FILTER_INV_SQUARE * ev.d0);
                            gm1 = dL_dG * (-G * FILTER_INV_SQUARE * ev.d1);
                            gz = dL_dz;
                            lowpass = true;
                        }
                    }
                    const uint32_t g = __float_as_uint(qd.w) & (USED ? REC_INDEX_MASK_USED : ~REC_FLAG_ALWAYS);
                    float* slot = pend + npend * SLOT_STRIDE + lane * PAIR_VALS;
                    *reinterpret_cast<float2*>(slot) = o0;
                    *reinterpret_cast<float2*>(slot + 2) = o1;
                    *reinterpret_cast<float2*>(slot + 4) = o2;
                    if (lane == 0) hdr[npend] = make_float4(__uint_as_float(g), pc.z, pc.w, 0.f);
                    if (__any_sync(FULLMASK, lowpass)) {          // low-pass branch (S/backward.cu:434-441): rare, reduced directly
                        const float r0 = warp_sum(gm0), r1 = warp_sum(gm1), r2 = warp_sum(gz);
                        float* acc = gacc + (size_t)g * GACC_STRIDE;
                        if (lane == 0) { atomicAdd(acc + 17, r2); atomicAdd(acc + 18, r0); atomicAdd(acc + 19, r1); }
                    }
                    if (++npend == BWD_SLOTS) { flush(BWD_SLOTS); npend = 0; }
                }
            }
        }
        // release the stage; the last of the CTA's warps to arrive refills it with the batch BWD_STAGES ahead.
        // sync_ring (validation only, gsr_set_option("dbg", 2)): a CTA barrier per batch makes the refill trivially ordered
        // after every warp's reads -- the reference point the decoupled handshake is stress-tested against.
        if (sync_ring) __syncthreads();
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
    if (npend) flush(npend);
}

cudaError_t launch_surfel_render_bwd(bool used, int ntiles, const uint32_t* tile_offset, const float4* planes, size_t pstride,
                                     int W, int H, int gx, const float* bg, const float* final_T, const uint32_t* n_contrib,
                                     const float* dL_dpix, const float* dL_dothers, float* gacc, int sync_ring, cudaStream_t s) {
    // the opt-in shared-memory size is a per-device function attribute
    static bool ready[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !ready[dev]) {
        e = cudaFuncSetAttribute(surfel_render_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(surfel_render_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BWD_SMEM);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) ready[dev] = true;
    }
    const dim3 grid(ntiles * BWD_SPLIT), block(BWD_WARPS * 32);
    if (used)
        surfel_render_bwd<true><<<grid, block, BWD_SMEM, s>>>(tile_offset, planes, pstride, W, H, gx, bg, final_T, n_contrib, dL_dpix,
                                                               dL_dothers, gacc, sync_ring);
    else
        surfel_render_bwd<false><<<grid, block, BWD_SMEM, s>>>(tile_offset, planes, pstride, W, H, gx, bg, final_T, n_contrib, dL_dpix,
                                                                dL_dothers, gacc, sync_ring);
    return cudaGetLastError();
}

}  // namespace gsr
