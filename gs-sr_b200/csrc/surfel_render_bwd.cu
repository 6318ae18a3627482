// surfel_render_bwd.cu -- per-tile back-to-front gradient walk of the 2DGS
// surfel blend for sm_100a.
//
// Result contract = reference backward renderCUDA, S/cuda_rasterizer/backward.cu:143-447
// (per-pair maths :287-444 incl. the distortion terms :347-364, the median-depth
// term :349-352, the median-normal quirk :381, the background term :391-394,
// the ray-splat branch :403-433 and the low-pass branch :434-441).
//
// B200 design:
//   * same CTA/warp/pixel mapping and cp.async.bulk record staging as the
//     forward, walking the tile's record stream from its last needed batch
//     (max last_contributor over the tile) down to the first;
//   * warp-level culling against the records' conservative pixel bounds and
//     against the warp's own max last_contributor;
//   * per-splat gradient sums are formed IN REGISTERS across the warp's 32
//     pixels with a transposed shuffle reduction (16 values -> 16 shuffles
//     instead of 80), then ONE 16-lane red.global.add.f32 burst into the
//     Gaussian's 80-byte accumulator -- the reference issues 19 global float
//     atomics per (pixel, splat) pair.
#include "common.cuh"
#include "async_copy.cuh"

namespace gsr {

constexpr int BWD_BATCH = 128;
constexpr uint32_t FULLMASK_B = 0xffffffffu;

// Transposed warp reduction of 16 per-lane values: after the call, lane L holds the
// full 32-lane sum of v[idx] with idx = ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1)
// (both lanes 2k and 2k+1 hold the same sum).
__device__ __forceinline__ float warp_reduce16_transposed(float (&v)[16], int lane) {
    {
        const bool hi = lane & 16;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            float send = hi ? v[i] : v[i + 8];
            float keep = hi ? v[i + 8] : v[i];
            v[i] = keep + __shfl_xor_sync(FULLMASK_B, send, 16);
        }
    }
    {
        const bool hi = lane & 8;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float send = hi ? v[i] : v[i + 4];
            float keep = hi ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(FULLMASK_B, send, 8);
        }
    }
    {
        const bool hi = lane & 4;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            float send = hi ? v[i] : v[i + 2];
            float keep = hi ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(FULLMASK_B, send, 4);
        }
    }
    {
        const bool hi = lane & 2;
        float send = hi ? v[0] : v[1];
        float keep = hi ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(FULLMASK_B, send, 2);
    }
    v[0] += __shfl_xor_sync(FULLMASK_B, v[0], 1);
    return v[0];
}

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULLMASK_B, x, o);
    return x;
}

__global__ void __launch_bounds__(TILE_PIX)
surfel_render_bwd(const uint2* __restrict__ ranges, const SplatRec* __restrict__ recs, int W, int H,
                  int gx, const float* __restrict__ bg, const float* __restrict__ final_T,
                  const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix,
                  const float* __restrict__ dL_dothers, float* __restrict__ gacc) {
    __shared__ __align__(128) SplatRec sbuf[2][BWD_BATCH];
    __shared__ __align__(8) uint64_t full_bar[2];
    __shared__ int s_maxlast;

    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wx0 = (warp & 1) * 8, wy0 = (warp >> 1) * 4;
    const int lx = wx0 + (lane & 7), ly = wy0 + (lane >> 3);
    const int px = tx * TILE + lx, py = ty * TILE + ly;
    const bool inside = px < W && py < H;
    const float fx = (float)lx, fy = (float)ly;
    const size_t N = (size_t)W * H;
    const size_t pid = (size_t)py * W + px;

    const uint2 range = ranges[tile];
    const int n = (int)(range.y - range.x);
    const SplatRec* src = recs + range.x;

    const int last = inside ? (int)n_contrib[pid] : 0;  // entries [0, last) contribute
    const int medpos = inside ? (int)n_contrib[pid + N] - 1 : -1;
    const int wlast = __reduce_max_sync(FULLMASK_B, last);
    if (threadIdx.x == 0) {
        s_maxlast = 0;
        mbar_init(&full_bar[0], 1);
        mbar_init(&full_bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (lane == 0 && wlast > 0) atomicMax(&s_maxlast, wlast);
    __syncthreads();
    const int maxlast = min(s_maxlast, n);
    if (maxlast <= 0) return;
    const int nb = (maxlast + BWD_BATCH - 1) / BWD_BATCH;  // batches [0, nb), walked from nb-1 down

    auto batch_count = [&](int b) { return min(BWD_BATCH, n - b * BWD_BATCH); };
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2 && nb - 1 - i >= 0; i++) {
            int b = nb - 1 - i;
            uint32_t bytes = (uint32_t)(batch_count(b) * sizeof(SplatRec));
            mbar_expect_tx(&full_bar[i], bytes);
            bulk_g2s(&sbuf[i][0], src + b * BWD_BATCH, bytes, &full_bar[i]);
        }
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
    const float bg0 = __ldg(bg), bg1 = __ldg(bg + 1), bg2 = __ldg(bg + 2);
    const float bg_dot_dpixel = bg0 * dpx0 + bg1 * dpx1 + bg2 * dpx2;
    const float MSCALE = FAR_N / (FAR_N - NEAR_N);
    const float DMD = (FAR_N * NEAR_N) / (FAR_N - NEAR_N);

    // running state of the reverse walk
    float T = T_final;
    float last_alpha = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, ar0 = 0.f, ar1 = 0.f, ar2 = 0.f;
    float last_depth = 0.f, accum_depth_rec = 0.f, accum_alpha_rec = 0.f;
    float ln0 = 0.f, ln1 = 0.f, ln2 = 0.f, an0 = 0.f, an1 = 0.f, an2 = 0.f, last_dL_dT = 0.f;

    for (int it = 0; it < nb; it++) {
        const int b = nb - 1 - it;
        const int stage = it & 1;
        const uint32_t parity = (uint32_t)((it >> 1) & 1);
        const int cnt = batch_count(b);
        const int base = b * BWD_BATCH;
        if (base < wlast) {
            mbar_wait(&full_bar[stage], parity);
            const SplatRec* sb = sbuf[stage];
            for (int c0 = ((cnt - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {
                if (base + c0 >= wlast) continue;
                const int e = c0 + lane;
                uint32_t bits = BOUNDS_EMPTY;
                if (e < cnt && base + e < wlast) bits = __float_as_uint(sb[e].cb.w);
                const int bx0 = bits & 15, bx1 = (bits >> 4) & 15, by0 = (bits >> 8) & 15, by1 = (bits >> 12) & 15;
                const bool hit = !(bits & BOUNDS_EMPTY) && bx0 <= wx0 + 7 && bx1 >= wx0 && by0 <= wy0 + 3 && by1 >= wy0;
                uint32_t m = __ballot_sync(FULLMASK_B, hit);
                while (m) {
                    const int bit = 31 - __clz(m);
                    m &= ~(1u << bit);
                    const int j = c0 + bit;
                    const int pos = base + j;  // 0-based list position == reference `contributor`
                    const float4* r = reinterpret_cast<const float4*>(sb + j);
                    const float4 tu = r[0], tv = r[1], tw = r[2];
                    bool valid = pos < last;
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
                    const bool ray = rho3d <= rho2d;
                    const float c_d = ray ? (s0 * tw.x + s1 * tw.y) + tw.z : tw.z;
                    valid = valid && !(c_d < NEAR_N);
                    const float power = -0.5f * rho;
                    valid = valid && !(power > 0.0f);
                    const float G = __expf(power);
                    const float alpha = fminf(ALPHA_MAX, tw.w * G);
                    valid = valid && !(alpha < ALPHA_MIN);
                    if (!__any_sync(FULLMASK_B, valid)) continue;

                    const float4 ng = r[3], cb = r[4];
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) v[i] = 0.f;
                    float gm0 = 0.f, gm1 = 0.f;
                    bool lowpass = false;
                    if (valid) {
                        T = __fdividef(T, 1.0f - alpha);
                        const float w = alpha * T;
                        float dL_dalpha = 0.f;
                        // colour
                        ar0 = last_alpha * lc0 + (1.f - last_alpha) * ar0; lc0 = cb.x;
                        ar1 = last_alpha * lc1 + (1.f - last_alpha) * ar1; lc1 = cb.y;
                        ar2 = last_alpha * lc2 + (1.f - last_alpha) * ar2; lc2 = cb.z;
                        dL_dalpha += (cb.x - ar0) * dpx0 + (cb.y - ar1) * dpx1 + (cb.z - ar2) * dpx2;
                        v[9] = w * dpx0; v[10] = w * dpx1; v[11] = w * dpx2;
                        // depth distortion + median depth
                        float dL_dz = 0.f;
                        const float icd = __frcp_rn(c_d);
                        const float m_d = MSCALE * (1.0f - NEAR_N * icd);
                        const float dmd_dd = DMD * icd * icd;
                        if (pos == medpos) dL_dz += dL_dmedian_depth;
                        const float dL_dweight = (final_D2 + m_d * m_d * final_A - 2.0f * m_d * final_D) * dL_dreg;
                        dL_dalpha += dL_dweight - last_dL_dT;
                        last_dL_dT = dL_dweight * alpha + (1.0f - alpha) * last_dL_dT;
                        const float dL_dmd = 2.0f * w * (m_d * final_A - final_D) * dL_dreg;
                        dL_dz += dL_dmd * dmd_dd;
                        // expected depth, alpha
                        accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                        last_depth = c_d;
                        dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
                        accum_alpha_rec = last_alpha + (1.f - last_alpha) * accum_alpha_rec;
                        dL_dalpha += (1.0f - accum_alpha_rec) * dL_daccum;
                        // normals (incl. the every-splat median-normal term, quirk Q1)
                        an0 = last_alpha * ln0 + (1.f - last_alpha) * an0; ln0 = ng.x;
                        an1 = last_alpha * ln1 + (1.f - last_alpha) * an1; ln1 = ng.y;
                        an2 = last_alpha * ln2 + (1.f - last_alpha) * an2; ln2 = ng.z;
                        dL_dalpha += (ng.x - an0) * dn0 + (ng.y - an1) * dn1 + (ng.z - an2) * dn2;
                        v[12] = w * dn0 + dmn0; v[13] = w * dn1 + dmn1; v[14] = w * dn2 + dmn2;

                        dL_dalpha *= T;
                        last_alpha = alpha;
                        dL_dalpha += (-T_final * __frcp_rn(1.0f - alpha)) * bg_dot_dpixel;
                        const float dL_dG = tw.w * dL_dalpha;
                        dL_dz += w * dL_ddepth;
                        if (ray) {
                            const float dls0 = dL_dG * -G * s0 + dL_dz * tw.x;
                            const float dls1 = dL_dG * -G * s1 + dL_dz * tw.y;
                            const float a0 = dls0 * ip, a1 = dls1 * ip;
                            const float a2 = -(a0 * s0 + a1 * s1);
                            // dL_dk = l x dL_dp ; dL_dl = dL_dp x k
                            const float dk0 = l1 * a2 - l2 * a1, dk1 = l2 * a0 - l0 * a2, dk2 = l0 * a1 - l1 * a0;
                            const float dl0 = a1 * kz - a2 * ky, dl1 = a2 * kx - a0 * kz, dl2 = a0 * ky - a1 * kx;
                            v[0] = -dk0; v[1] = -dk1; v[2] = -dk2;
                            v[3] = -dl0; v[4] = -dl1; v[5] = -dl2;
                            // gradient w.r.t. the tile-local Tw (mapped back to global below)
                            v[6] = fx * dk0 + fy * dl0 + dL_dz * s0;
                            v[7] = fx * dk1 + fy * dl1 + dL_dz * s1;
                            v[8] = fx * dk2 + fy * dl2 + dL_dz;
                        } else {
                            gm0 = dL_dG * (-G * FILTER_INV_SQUARE * d0);
                            gm1 = dL_dG * (-G * FILTER_INV_SQUARE * d1);
                            v[8] = dL_dz;
                            lowpass = true;
                        }
                        v[15] = G * dL_dalpha;
                    }
                    const bool any_lowpass = __any_sync(FULLMASK_B, lowpass);
                    const float red = warp_reduce16_transposed(v, lane);
                    const uint32_t g = __float_as_uint(ng.w);
                    float* acc = gacc + (size_t)g * GACC_STRIDE;
                    const int vi = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    // Tu' = Tu - ox*Tw, Tv' = Tv - oy*Tw  =>  dTw = dTw' - ox*dTu' - oy*dTv'
                    const float ox = (float)(tx * TILE), oy = (float)(ty * TILE);
                    // lanes holding dTw'[c] (vi = 6+c) need dTu'[c] (vi = c) and dTv'[c] (vi = 3+c)
                    // source lane for value index q: lane bits (4,3,2,1) = q
                    const int c = vi - 6;
                    const float su = __shfl_sync(FULLMASK_B, red, ((c & 3) << 1) & 31);
                    const float sv = __shfl_sync(FULLMASK_B, red, (((c + 3) & 15) << 1) & 31);
                    float outv = red;
                    if (vi >= 6 && vi <= 8) outv = red - ox * su - oy * sv;
                    if ((lane & 1) == 0) atomicAdd(acc + vi, outv);
                    if (any_lowpass) {
                        const float r0 = warp_sum(gm0), r1 = warp_sum(gm1);
                        if (lane == 0) { atomicAdd(acc + 16, r0); atomicAdd(acc + 17, r1); }
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && it + 2 < nb) {
            const int b2 = nb - 1 - (it + 2);
            uint32_t bytes = (uint32_t)(batch_count(b2) * sizeof(SplatRec));
            fence_proxy_async();
            mbar_expect_tx(&full_bar[stage], bytes);
            bulk_g2s(&sbuf[stage][0], src + b2 * BWD_BATCH, bytes, &full_bar[stage]);
        }
    }
}

}  // namespace gsr
