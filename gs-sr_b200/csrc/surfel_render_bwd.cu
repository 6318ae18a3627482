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
//     culling as the forward (one CTA = 8 warps = one tile; GSR_BWD_WARPS = 4 / 2 builds half-tile /
//     strip CTAs that walk the same list), but the warps of a CTA are decoupled: a ring whose stages are
//     refilled by whichever warp arrives last, no __syncthreads() in the walk; the walk starts
//     at the CTA's last needed batch (max last_contributor);
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

#ifndef GSR_BWD_PAIR2
#define GSR_BWD_PAIR2 0         // evaluating two walked entries before their gradients: 1.76 -> 1.93 ms (the kernel is bound by the
#endif                          // shared-memory pipe and sits at its register cap; the forward gains 2 % from the same change)
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
    // distortion terms (S/backward.cu:347-364) with the upstream weight folded into the three per-pixel constants:
    //   dL/dweight = (m^2 A - 2 m D + D2) dL_dreg = m (m rA + rB) + rC,   2 (m A - D) dL_dreg = (m + m) rA + rB
    const float rA = final_A * dL_dreg, rB = -2.0f * final_D * dL_dreg, rC = final_D2 * dL_dreg;
    // upstream gradients paired against the record halves (normal.xy | normal.z, colour.r | colour.gb)
    const float2 k_dn01 = make_float2(dn0, dn1), k_dn2px0 = make_float2(dn2, dpx0), k_px12 = make_float2(dpx1, dpx2);
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
        // accumulators paired for the packed adds / FMAs: the slot's three 64-bit words (a0 a1 | a2 q | w v) and the
        // halves of the pixels' upstream-gradient words are the operands as they come out of shared memory
        const float2 z2 = make_float2(0.f, 0.f);
        float2 m01 = z2, m2q = z2, x01 = z2, y01 = z2, c01 = z2, c2n0 = z2, n12 = z2;
        float x2 = 0.f, y2 = 0.f, so = 0.f;
        if (ps < np) {
            const float* pb = pend + ps * SLOT_STRIDE + ph * (BWD_PPL * PAIR_VALS);
            const float4* pk = pixc + ph * PIXC_GROUP;
#pragma unroll
            for (int i = 0; i < BWD_PPL; i++) {
                const float2 u0 = *reinterpret_cast<const float2*>(pb + i * PAIR_VALS);
                const float2 u1 = *reinterpret_cast<const float2*>(pb + i * PAIR_VALS + 2);
                const float2 u2 = *reinterpret_cast<const float2*>(pb + i * PAIR_VALS + 4);
                const float4 k0 = pk[i * 2], k1 = pk[i * 2 + 1];
                m01 = fadd2(m01, u0);
                m2q = fadd2(m2q, u1);
                if ((i & 7) != 0) {
                    const float xi = (float)(i & 7);
                    x01 = ffma2s(u0, xi, x01); x2 = fmaf(xi, u1.x, x2);
                }
                if ((i >> 3) != 0) {
                    const float yi = (float)(i >> 3);
                    y01 = ffma2s(u0, yi, y01); y2 = fmaf(yi, u1.x, y2);
                }
                c01 = ffma2s(make_float2(k0.x, k0.y), u2.x, c01);
                c2n0 = ffma2s(make_float2(k0.z, k0.w), u2.x, c2n0);
                n12 = ffma2s(make_float2(k1.x, k1.y), u2.x, n12);
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
                        c2n0.y += __ldg(dL_dothers + q + 8 * N);
                        n12.x += __ldg(dL_dothers + q + 9 * N); n12.y += __ldg(dL_dothers + q + 10 * N);
                    }
            }
            // block-local pixel offsets -> offsets from the splat's moment origin (hdr.y, hdr.z; tile-local)
            const float4 h4 = hdr[ps];
            const float ox = (float)wx0 - h4.y, oy = (float)(wy0 + ph * (BWD_PPL / 8)) - h4.z;
            x01 = ffma2s(m01, ox, x01); x2 = fmaf(ox, m2q.x, x2);
            y01 = ffma2s(m01, oy, y01); y2 = fmaf(oy, m2q.x, y2);
        }
        float2 xys = make_float2(x2, y2);
        auto shfl2 = [](float2 v, int o) {
            return make_float2(__shfl_xor_sync(FULLMASK, v.x, o), __shfl_xor_sync(FULLMASK, v.y, o));
        };
#pragma unroll
        for (int o = BWD_SLOTS; o < 32; o <<= 1) {
            m01 = fadd2(m01, shfl2(m01, o)); m2q = fadd2(m2q, shfl2(m2q, o));
            x01 = fadd2(x01, shfl2(x01, o)); y01 = fadd2(y01, shfl2(y01, o));
            xys = fadd2(xys, shfl2(xys, o));
            c01 = fadd2(c01, shfl2(c01, o)); c2n0 = fadd2(c2n0, shfl2(c2n0, o)); n12 = fadd2(n12, shfl2(n12, o));
            so += __shfl_xor_sync(FULLMASK, so, o);
        }
        const float m0 = m01.x, m1 = m01.y, m2 = m2q.x, qs = m2q.y, x0 = x01.x, x1 = x01.y, y0 = y01.x, y1 = y01.y;
        const float c0 = c01.x, c1 = c01.y, c2 = c2n0.x, n0 = c2n0.y, n1 = n12.x, n2 = n12.y;
        x2 = xys.x; y2 = xys.y;
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
            uint32_t sb_addr;     // opaque to the compiler: otherwise it is rematerialised (S2UR + 5 uniform ops) per pair
            asm volatile("mov.u32 %0, %1;" : "=r"(sb_addr) : "r"(smem_u32(&sb[0][0])));
            constexpr uint32_t PLANE_B = BWD_BATCH * 16u;
            for (int c0 = ((cnt - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
                if (base + c0 >= wlast) continue;
                const int e = c0 + lane;
                bool hit = false;
                if (e < cnt && base + e < wlast) {
                    if (USED) hit = (__float_as_uint(sb[3][e].w) >> (REC_USED_SHIFT + warp)) & 1u;
                    else hit = entry_hits_block(sb[0][e], sb[1][e], sb[2][e], sb[3][e], bx0, bx1, by0, by1);
                }
                uint32_t m = __ballot_sync(FULLMASK, hit);
                // gradient of one evaluated pair (list position pos, record at shared address ra); the state chain T / rec runs
                // through it back to front
                auto pair_grad = [&](const PairEval& ev, const float opac, const float qdw, const uint32_t ra, const int pos) {
                    const bool valid = ev.valid && pos < last;
                    if (!__any_sync(FULLMASK, valid)) return;

                    const float4 pn = lds128(ra + 4 * PLANE_B), pc = lds128(ra + 5 * PLANE_B);
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
                        const float dL_dweight = fmaf(m_d, fmaf(m_d, rA, rB), rC);
                        // S = <this splat's channel values, the pixel's upstream gradients>; rec = the same blended over the
                        // splats behind it (one chain instead of the reference's accum_rec per channel)
                        float2 sp = fmul2(make_float2(pn.x, pn.y), k_dn01);
                        sp = ffma2(make_float2(pn.z, pn.w), k_dn2px0, sp);
                        sp = ffma2(make_float2(pc.x, pc.y), k_px12, sp);
                        const float S = (sp.x + sp.y) + fmaf(c_d, dL_ddepth, dL_dweight);
                        const float D = S - rec;
                        rec = fmaf(alpha, D, rec);
                        const float dL_dalpha = fmaf(D, T, k_tail * ria);
                        float dL_dz = fmaf(fmaf(m_d + m_d, rA, rB), dmd_dd, dL_ddepth) * w;
                        if (pos == medpos) dL_dz += dL_dmedian_depth;
                        const float v = G * dL_dalpha;                 // dL/dopacity share
                        const float dL_dG = opac * dL_dalpha;
                        o2 = make_float2(w, v);
                        if (ev.ray) {
                            // dL/ds = dL_dG * (-G) * s ;  s = p.xy / p.z ; depth = det(T) / p.z
                            const float gs = dL_dG * -G * ev.ip;
                            o0 = fmul2s(make_float2(ev.s0, ev.s1), gs);
                            const float a0 = o0.x, a1 = o0.y;
                            const float q = dL_dz * ev.ip;
                            const float a2 = -fmaf(a0, ev.s0, fmaf(a1, ev.s1, q * c_d));
                            o1 = make_float2(a2, q);
                        } else {
                            gm0 = dL_dG * (-G * FILTER_INV_SQUARE * ev.d0);
                            gm1 = dL_dG * (-G * FILTER_INV_SQUARE * ev.d1);
                            gz = dL_dz;
                            lowpass = true;
                        }
                    }
                    const uint32_t g = __float_as_uint(qdw) & (USED ? REC_INDEX_MASK_USED : ~REC_FLAG_ALWAYS);
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
                };
#if GSR_BWD_PAIR2
                // two walked entries at a time: their evaluations are independent chains and hide each other's latencies
                while (m) {
                    const int b1 = 31 - __clz(m);
                    m &= ~(1u << b1);
                    const bool two = m != 0u;
                    const int b2 = two ? 31 - __clz(m) : b1;
                    if (two) m &= ~(1u << b2);
                    const uint32_t r1 = sb_addr + (uint32_t)(c0 + b1) * 16u, r2 = sb_addr + (uint32_t)(c0 + b2) * 16u;
                    const float4 c1 = lds128(r1 + 2 * PLANE_B), d1 = lds128(r1 + 3 * PLANE_B);
                    const float4 c2 = lds128(r2 + 2 * PLANE_B), d2 = lds128(r2 + 3 * PLANE_B);
                    const PairEval e1 = eval_pair(lds128(r1), lds128(r1 + PLANE_B), c1, d1, fx, fy);
                    const PairEval e2 = eval_pair(lds128(r2), lds128(r2 + PLANE_B), c2, d2, fx, fy);
                    pair_grad(e1, c1.w, d1.w, r1, base + c0 + b1);
                    if (two) pair_grad(e2, c2.w, d2.w, r2, base + c0 + b2);
                }
#else
                while (m) {
                    const int bit = 31 - __clz(m);
                    m &= ~(1u << bit);
                    const int j = c0 + bit;
                    // record reads through an explicit 32-bit shared address (one uniform shift-add per pair; the generic
                    // form made ptxas rebuild the stage's shared-window base for every pair)
                    const uint32_t ra = sb_addr + (uint32_t)j * 16u;
                    const float4 qa = lds128(ra), qb = lds128(ra + PLANE_B), qc = lds128(ra + 2 * PLANE_B), qd = lds128(ra + 3 * PLANE_B);
                    const PairEval ev = eval_pair(qa, qb, qc, qd, fx, fy);
                    pair_grad(ev, qc.w, qd.w, ra, base + j);      // base + j: 0-based list position == reference `contributor`
                }
#endif
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
