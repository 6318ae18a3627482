// ssim.cu -- fused SSIM loss (forward + backward) with the 11x11 Gaussian window, for sm_100a.
//
// Result contract = VanillaScene._ssim / .ssim (/root/reference/gssr/scene/vanilla_scene.py:32-61): depthwise
// zero-padded 11x11 Gaussian window (sigma 1.5, outer product of the normalised 1-D kernel),
//   mu1 = G*x, mu2 = G*y, s11 = G*x^2 - mu1^2, s22 = G*y^2 - mu2^2, s12 = G*(xy) - mu1 mu2,
//   ssim = ((2 mu1 mu2 + C1)(2 s12 + C2)) / ((mu1^2 + mu2^2 + C1)(s11 + s22 + C2)),  C1 = 0.01^2, C2 = 0.03^2,
// averaged over every pixel and channel.  The reference spends 5 depthwise conv2d (121 taps each) + ~15
// elementwise kernels in the forward and 3 more convolutions + ~30 kernels in autograd's backward (3.0 ms per
// iteration at 1600x1060, as much as the backward rasterizer); the window is separable, so:
//
// B200 design: one CTA per 32x32 output tile and channel.  The 42x42 input halo of x and y is staged in shared
// memory once, a horizontal 11-tap pass produces the five running quantities for 42 rows, a vertical pass
// finishes them, and the same thread evaluates the SSIM value plus the three partial derivatives the backward
// needs (d ssim / d mu1, / d G*x^2, / d G*xy).  The backward is the adjoint: the three derivative maps are
// convolved with the same (symmetric) window and combined with x and y per pixel.  HBM traffic: 24 + 60 B per
// pixel-channel instead of ~50 full-image passes.  Tile sums are written per CTA and reduced by the caller
// (deterministic, no float atomics).
#include "common.cuh"
#include "../../include/gsr_b200.h"

namespace gsr {

constexpr int SSIM_T = 32;                 // output tile edge
constexpr int SSIM_R = 5;                  // window radius (11 taps)
constexpr int SSIM_IN = SSIM_T + 2 * SSIM_R;   // 42
constexpr int SSIM_THREADS = 256;

struct SsimWindow { float g[11]; };

__device__ __forceinline__ float tile_load(const float* __restrict__ img, int H, int W, int y, int x) {
    return (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(img + (size_t)y * W + x) : 0.f;
}

__global__ void __launch_bounds__(SSIM_THREADS)
ssim_forward_kernel(int H, int W, const float* __restrict__ img1, const float* __restrict__ img2, SsimWindow win,
                    float* __restrict__ dmu1, float* __restrict__ dE11, float* __restrict__ dE12,
                    float* __restrict__ tile_sums) {
    __shared__ float sx[SSIM_IN][SSIM_IN + 1], sy[SSIM_IN][SSIM_IN + 1];
    __shared__ float hq[5][SSIM_IN][SSIM_T];
    __shared__ float wsum[SSIM_THREADS / 32];
    const int c = blockIdx.z;
    const size_t plane = (size_t)H * W;
    const float* x = img1 + c * plane;
    const float* y = img2 + c * plane;
    const int ox = blockIdx.x * SSIM_T, oy = blockIdx.y * SSIM_T;
    for (int i = threadIdx.x; i < SSIM_IN * SSIM_IN; i += SSIM_THREADS) {
        const int r = i / SSIM_IN, q = i - r * SSIM_IN;
        sx[r][q] = tile_load(x, H, W, oy + r - SSIM_R, ox + q - SSIM_R);
        sy[r][q] = tile_load(y, H, W, oy + r - SSIM_R, ox + q - SSIM_R);
    }
    __syncthreads();
    // horizontal pass: 42 rows x 32 columns
    for (int i = threadIdx.x; i < SSIM_IN * SSIM_T; i += SSIM_THREADS) {
        const int r = i / SSIM_T, q = i - r * SSIM_T;
        float a = 0.f, b = 0.f, aa = 0.f, bb = 0.f, ab = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = win.g[k], u = sx[r][q + k], v = sy[r][q + k];
            const float gu = g * u, gv = g * v;
            a += gu; b += gv; aa = fmaf(gu, u, aa); bb = fmaf(gv, v, bb); ab = fmaf(gu, v, ab);
        }
        hq[0][r][q] = a; hq[1][r][q] = b; hq[2][r][q] = aa; hq[3][r][q] = bb; hq[4][r][q] = ab;
    }
    __syncthreads();
    // vertical pass + SSIM + derivatives: 4 outputs per thread (rows ty, ty+8, ty+16, ty+24)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    float local = 0.f;
#pragma unroll
    for (int o = 0; o < 4; o++) {
        const int r = ty + 8 * o;
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = win.g[k];
            mu1 = fmaf(g, hq[0][r + k][tx], mu1);
            mu2 = fmaf(g, hq[1][r + k][tx], mu2);
            e11 = fmaf(g, hq[2][r + k][tx], e11);
            e22 = fmaf(g, hq[3][r + k][tx], e22);
            e12 = fmaf(g, hq[4][r + k][tx], e12);
        }
        const int py = oy + r, px = ox + tx;
        if (py < H && px < W) {
            const float mu1s = mu1 * mu1, mu2s = mu2 * mu2, mu12 = mu1 * mu2;
            const float s11 = e11 - mu1s, s22 = e22 - mu2s, s12 = e12 - mu12;
            const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2, B1 = mu1s + mu2s + C1, B2 = s11 + s22 + C2;
            const float iB1 = 1.f / B1, iB2 = 1.f / B2;
            const float S = (A1 * A2) * (iB1 * iB2);
            local += S;
            const size_t pid = c * plane + (size_t)py * W + px;
            dmu1[pid] = 2.f * mu2 * (A2 - A1) * (iB1 * iB2) - 2.f * mu1 * S * (iB1 - iB2);
            dE11[pid] = -S * iB2;
            dE12[pid] = 2.f * A1 * (iB1 * iB2);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if (tx == 0) wsum[ty] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < SSIM_THREADS / 32; i++) t += wsum[i];
        tile_sums[((size_t)c * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
    }
}

// dL/dx = scale * ( G*dmu1 + 2 x (G*dE11) + y (G*dE12) ),  scale = dL/dssim_mean / (C H W)
__global__ void __launch_bounds__(SSIM_THREADS)
ssim_backward_kernel(int H, int W, const float* __restrict__ img1, const float* __restrict__ img2, SsimWindow win,
                     const float* __restrict__ dmu1, const float* __restrict__ dE11, const float* __restrict__ dE12,
                     const float* __restrict__ upstream, float inv_count, float* __restrict__ dL_dimg1) {
    __shared__ float sm[3][SSIM_IN][SSIM_IN + 1];
    __shared__ float hq[3][SSIM_IN][SSIM_T];
    const int c = blockIdx.z;
    const size_t plane = (size_t)H * W;
    const float* m0 = dmu1 + c * plane;
    const float* m1 = dE11 + c * plane;
    const float* m2 = dE12 + c * plane;
    const int ox = blockIdx.x * SSIM_T, oy = blockIdx.y * SSIM_T;
    for (int i = threadIdx.x; i < SSIM_IN * SSIM_IN; i += SSIM_THREADS) {
        const int r = i / SSIM_IN, q = i - r * SSIM_IN;
        const int yy = oy + r - SSIM_R, xx = ox + q - SSIM_R;
        sm[0][r][q] = tile_load(m0, H, W, yy, xx);
        sm[1][r][q] = tile_load(m1, H, W, yy, xx);
        sm[2][r][q] = tile_load(m2, H, W, yy, xx);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SSIM_IN * SSIM_T; i += SSIM_THREADS) {
        const int r = i / SSIM_T, q = i - r * SSIM_T;
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = win.g[k];
            a = fmaf(g, sm[0][r][q + k], a); b = fmaf(g, sm[1][r][q + k], b); d = fmaf(g, sm[2][r][q + k], d);
        }
        hq[0][r][q] = a; hq[1][r][q] = b; hq[2][r][q] = d;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float scale = __ldg(upstream) * inv_count;
#pragma unroll
    for (int o = 0; o < 4; o++) {
        const int r = ty + 8 * o;
        float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
        for (int k = 0; k < 11; k++) {
            const float g = win.g[k];
            a = fmaf(g, hq[0][r + k][tx], a); b = fmaf(g, hq[1][r + k][tx], b); d = fmaf(g, hq[2][r + k][tx], d);
        }
        const int py = oy + r, px = ox + tx;
        if (py < H && px < W) {
            const size_t pid = c * plane + (size_t)py * W + px;
            const float x = __ldg(img1 + pid), y = __ldg(img2 + pid);
            dL_dimg1[pid] = scale * (a + 2.f * x * b + y * d);
        }
    }
}

}  // namespace gsr

extern "C" size_t gsr_ssim_tile_count(int channels, int height, int width) {
    using namespace gsr;
    if (channels <= 0 || height <= 0 || width <= 0) return 0;
    return (size_t)channels * ((height + SSIM_T - 1) / SSIM_T) * ((width + SSIM_T - 1) / SSIM_T);
}

extern "C" int gsr_ssim_forward(int channels, int height, int width, const float* img1, const float* img2,
                                const float* window11, float* dmu1, float* dE11, float* dE12, float* tile_sums,
                                void* stream_v) {
    using namespace gsr;
    if (channels <= 0 || height <= 0 || width <= 0 || !img1 || !img2 || !window11 || !dmu1 || !dE11 || !dE12 || !tile_sums ||
        channels > 65535) {
        set_error("gsr_ssim_forward: invalid argument");
        return GSR_E_INVALID;
    }
    SsimWindow w;
    for (int i = 0; i < 11; i++) w.g[i] = window11[i];
    dim3 grid((width + SSIM_T - 1) / SSIM_T, (height + SSIM_T - 1) / SSIM_T, channels);
    ssim_forward_kernel<<<grid, SSIM_THREADS, 0, (cudaStream_t)stream_v>>>(height, width, img1, img2, w, dmu1, dE11, dE12, tile_sums);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}

extern "C" int gsr_ssim_backward(int channels, int height, int width, const float* img1, const float* img2,
                                 const float* window11, const float* dmu1, const float* dE11, const float* dE12,
                                 const float* upstream_dev, float* dL_dimg1, void* stream_v) {
    using namespace gsr;
    if (channels <= 0 || height <= 0 || width <= 0 || !img1 || !img2 || !window11 || !dmu1 || !dE11 || !dE12 || !upstream_dev ||
        !dL_dimg1 || channels > 65535) {
        set_error("gsr_ssim_backward: invalid argument");
        return GSR_E_INVALID;
    }
    SsimWindow w;
    for (int i = 0; i < 11; i++) w.g[i] = window11[i];
    dim3 grid((width + SSIM_T - 1) / SSIM_T, (height + SSIM_T - 1) / SSIM_T, channels);
    const float inv_count = 1.0f / ((float)channels * (float)height * (float)width);
    ssim_backward_kernel<<<grid, SSIM_THREADS, 0, (cudaStream_t)stream_v>>>(height, width, img1, img2, w, dmu1, dE11, dE12,
                                                                             upstream_dev, inv_count, dL_dimg1);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}
