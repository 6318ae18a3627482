// surfel_post.cu -- fused image-space post-processing of the 2DGS allmap (forward + backward), for sm_100a.
//
// Result contract = the block of TwoDGSScene.render that follows the rasterizer call
// (/root/reference/gssr/scene/twodgs_scene.py:88-117) with depth_to_normal / depths_to_points
// (/root/reference/gssr/utils/point_utils.py:9-37):
//   render_normal = allmap[2:5] rotated to world space with world_view_transform[:3,:3].T          (:92-93)
//   depth_median  = nan_to_num(allmap[5]),  depth_expected = nan_to_num(allmap[0] / allmap[1])        (:96-102)
//   surf_depth    = depth_expected (1 - depth_ratio) + depth_ratio depth_median                       (:110)
//   surf_normal   = normalize(cross(P[r+1,c] - P[r-1,c], P[r,c+1] - P[r,c-1])) * allmap[1].detach(),  (:113-116)
//                   P = surf_depth * rays_d + rays_o, zero on the one-pixel border (point_utils.py:27-37)
// and autograd's backward of exactly that graph w.r.t. allmap for arbitrary upstream gradients on the three
// outputs.  The reference spends ~45 small torch kernels (meshgrid, stack, two 3x3 inverses, three matmuls over
// the H*W x 3 point list, cross, normalize, slicing copies) per iteration on it, and as many again in backward.
//
// B200 design: four streaming per-pixel kernels, every operand read through the read-only path (neighbour taps
// hit L1/L2), all outputs fully written (no memsets): forward A (depth select + normal rotation), forward B
// (finite-difference normal), backward A (per-pixel adjoint of the cross/normalize -> two 3-vectors),
// backward B (gather of the four neighbours' adjoints, ray dot, division / nan_to_num adjoint, normal rotation).
// The only deliberate difference: where allmap[1] == 0 the reference's backward yields NaN for allmap[0:2]
// (0/0 in the division adjoint; never consumed by the rasterizer because such pixels have no contributor) --
// this implementation writes 0 there.
#include "common.cuh"
#include "../../include/gsr_b200.h"

namespace gsr {

struct PostCam {            // loaded once per thread from a 24-float device array
    float K[9];             // rays_d = [x, y, 1] @ K   (K = intrins.inverse().T @ c2w[:3,:3].T), row-major
    float o[3];             // rays_o = c2w[:3,3]
    float R[9];             // world_view_transform[:3,:3] row-major: render_normal[c] = sum_k n_view[k] * R[c][k]
};

__device__ __forceinline__ PostCam load_cam(const float* __restrict__ cam) {
    PostCam c;
#pragma unroll
    for (int i = 0; i < 9; i++) c.K[i] = __ldg(cam + i);
#pragma unroll
    for (int i = 0; i < 3; i++) c.o[i] = __ldg(cam + 9 + i);
#pragma unroll
    for (int i = 0; i < 9; i++) c.R[i] = __ldg(cam + 12 + i);
    return c;
}

__device__ __forceinline__ float nan_to_num0(float v) {        // torch.nan_to_num(v, 0, 0): nan -> 0, +inf -> 0, -inf -> lowest
    if (v != v) return 0.f;
    if (v == INFINITY) return 0.f;
    if (v == -INFINITY) return -3.4028234663852886e38f;
    return v;
}
__device__ __forceinline__ bool finite_f(float v) { return fabsf(v) <= 3.4028234663852886e38f; }

__device__ __forceinline__ float3 ray_dir(const PostCam& c, int x, int y) {
    const float fx = (float)x, fy = (float)y;
    return make_float3(fx * c.K[0] + fy * c.K[3] + c.K[6], fx * c.K[1] + fy * c.K[4] + c.K[7],
                       fx * c.K[2] + fy * c.K[5] + c.K[8]);
}
__device__ __forceinline__ float3 point_at(const PostCam& c, const float* __restrict__ depth, int W, int x, int y) {
    const float d = __ldg(depth + (size_t)y * W + x);
    const float3 r = ray_dir(c, x, y);
    return make_float3(d * r.x + c.o[0], d * r.y + c.o[1], d * r.z + c.o[2]);
}

// forward A: surf_depth (1,H,W), render_normal (3,H,W)
__global__ void __launch_bounds__(256)
post_fwd_a(int H, int W, const float* __restrict__ allmap, const float* __restrict__ cam, float depth_ratio,
           float* __restrict__ render_normal, float* __restrict__ surf_depth) {
    const size_t N = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= N) return;
    const PostCam c = load_cam(cam);
    const float a0 = __ldg(allmap + p), a1 = __ldg(allmap + N + p);
    const float n0 = __ldg(allmap + 2 * N + p), n1 = __ldg(allmap + 3 * N + p), n2 = __ldg(allmap + 4 * N + p);
    const float a5 = __ldg(allmap + 5 * N + p);
    render_normal[p] = n0 * c.R[0] + n1 * c.R[1] + n2 * c.R[2];
    render_normal[N + p] = n0 * c.R[3] + n1 * c.R[4] + n2 * c.R[5];
    render_normal[2 * N + p] = n0 * c.R[6] + n1 * c.R[7] + n2 * c.R[8];
    const float dmed = nan_to_num0(a5), dexp = nan_to_num0(a0 / a1);
    surf_depth[p] = dexp * (1.f - depth_ratio) + depth_ratio * dmed;
}

// forward B: surf_normal (3,H,W) from surf_depth
__global__ void __launch_bounds__(256)
post_fwd_b(int H, int W, const float* __restrict__ allmap, const float* __restrict__ cam,
           const float* __restrict__ surf_depth, float* __restrict__ surf_normal) {
    const size_t N = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= N) return;
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    float3 n = make_float3(0.f, 0.f, 0.f);
    if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
        const PostCam c = load_cam(cam);
        const float3 pu = point_at(c, surf_depth, W, x, y - 1), pd = point_at(c, surf_depth, W, x, y + 1);
        const float3 pl = point_at(c, surf_depth, W, x - 1, y), pr = point_at(c, surf_depth, W, x + 1, y);
        const float3 dx = make_float3(pd.x - pu.x, pd.y - pu.y, pd.z - pu.z);      // points[2:, 1:-1] - points[:-2, 1:-1]
        const float3 dy = make_float3(pr.x - pl.x, pr.y - pl.y, pr.z - pl.z);      // points[1:-1, 2:] - points[1:-1, :-2]
        const float3 cr = cross3(dx, dy);
        const float len = fmaxf(sqrtf(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z), 1e-12f);   // F.normalize eps
        const float alpha = __ldg(allmap + N + p);
        const float s = alpha / len;
        n = make_float3(cr.x * s, cr.y * s, cr.z * s);
    }
    surf_normal[p] = n.x; surf_normal[N + p] = n.y; surf_normal[2 * N + p] = n.z;
}

// backward A: adjoints of dx and dy at every pixel (zero on the border): gd (6,H,W)
__global__ void __launch_bounds__(256)
post_bwd_a(int H, int W, const float* __restrict__ allmap, const float* __restrict__ cam,
           const float* __restrict__ surf_depth, const float* __restrict__ g_surf_normal, float* __restrict__ gd) {
    const size_t N = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= N) return;
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    float3 gdx = make_float3(0.f, 0.f, 0.f), gdy = gdx;
    if (x >= 1 && x < W - 1 && y >= 1 && y < H - 1) {
        const PostCam c = load_cam(cam);
        const float3 pu = point_at(c, surf_depth, W, x, y - 1), pd = point_at(c, surf_depth, W, x, y + 1);
        const float3 pl = point_at(c, surf_depth, W, x - 1, y), pr = point_at(c, surf_depth, W, x + 1, y);
        const float3 dx = make_float3(pd.x - pu.x, pd.y - pu.y, pd.z - pu.z);
        const float3 dy = make_float3(pr.x - pl.x, pr.y - pl.y, pr.z - pl.z);
        const float3 cr = cross3(dx, dy);
        const float norm = sqrtf(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
        const float alpha = __ldg(allmap + N + p);
        // upstream on n = cr / max(|cr|, eps):  g = g_surf_normal * alpha (alpha detached)
        const float3 g = make_float3(__ldg(g_surf_normal + p) * alpha, __ldg(g_surf_normal + N + p) * alpha,
                                     __ldg(g_surf_normal + 2 * N + p) * alpha);
        float3 gc;
        if (norm > 1e-12f) {
            const float inv = 1.f / norm;
            const float3 n = make_float3(cr.x * inv, cr.y * inv, cr.z * inv);
            const float ng = n.x * g.x + n.y * g.y + n.z * g.z;
            gc = make_float3((g.x - n.x * ng) * inv, (g.y - n.y * ng) * inv, (g.z - n.z * ng) * inv);
        } else {
            gc = make_float3(g.x * 1e12f, g.y * 1e12f, g.z * 1e12f);     // clamp_min(eps) branch of F.normalize
        }
        gdx = cross3(dy, gc);      // d (dx x dy).gc / d dx
        gdy = cross3(gc, dx);      // d (dx x dy).gc / d dy
    }
    gd[p] = gdx.x; gd[N + p] = gdx.y; gd[2 * N + p] = gdx.z;
    gd[3 * N + p] = gdy.x; gd[4 * N + p] = gdy.y; gd[5 * N + p] = gdy.z;
}

__device__ __forceinline__ float3 ld3(const float* __restrict__ base, size_t N, size_t q) {
    return make_float3(__ldg(base + q), __ldg(base + N + q), __ldg(base + 2 * N + q));
}

// backward B: dL/dallmap (11,H,W), fully written
__global__ void __launch_bounds__(256)
post_bwd_b(int H, int W, const float* __restrict__ allmap, const float* __restrict__ cam, float depth_ratio,
           const float* __restrict__ gd, const float* __restrict__ g_render_normal,
           const float* __restrict__ g_surf_depth, float* __restrict__ dL_dallmap) {
    const size_t N = (size_t)H * W;
    const size_t p = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (p >= N) return;
    const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
    const PostCam c = load_cam(cam);
    // dL/dP(p): p is the "+row" tap of (y-1), the "-row" tap of (y+1), the "+col" tap of (x-1), the "-col" tap of (x+1)
    float3 gp = make_float3(0.f, 0.f, 0.f);
    if (y >= 1) { const float3 t = ld3(gd, N, p - W); gp.x += t.x; gp.y += t.y; gp.z += t.z; }
    if (y < H - 1) { const float3 t = ld3(gd, N, p + W); gp.x -= t.x; gp.y -= t.y; gp.z -= t.z; }
    if (x >= 1) { const float3 t = ld3(gd + 3 * N, N, p - 1); gp.x += t.x; gp.y += t.y; gp.z += t.z; }
    if (x < W - 1) { const float3 t = ld3(gd + 3 * N, N, p + 1); gp.x -= t.x; gp.y -= t.y; gp.z -= t.z; }
    const float3 r = ray_dir(c, x, y);
    const float gdepth = (g_surf_depth ? __ldg(g_surf_depth + p) : 0.f) + gp.x * r.x + gp.y * r.y + gp.z * r.z;
    const float a0 = __ldg(allmap + p), a1 = __ldg(allmap + N + p), a5 = __ldg(allmap + 5 * N + p);
    const float q = a0 / a1;
    float g0 = 0.f, g1 = 0.f, g5 = 0.f;
    if (finite_f(q) && a1 != 0.f) {                    // nan_to_num passes the gradient only where its input is finite
        const float ge = gdepth * (1.f - depth_ratio);
        g0 = ge / a1;
        g1 = -ge * q / a1;
    }
    if (finite_f(a5)) g5 = gdepth * depth_ratio;
    float gn0 = 0.f, gn1 = 0.f, gn2 = 0.f;
    if (g_render_normal) {
        const float3 g = ld3(g_render_normal, N, p);
        gn0 = g.x * c.R[0] + g.y * c.R[3] + g.z * c.R[6];
        gn1 = g.x * c.R[1] + g.y * c.R[4] + g.z * c.R[7];
        gn2 = g.x * c.R[2] + g.y * c.R[5] + g.z * c.R[8];
    }
    dL_dallmap[p] = g0; dL_dallmap[N + p] = g1;
    dL_dallmap[2 * N + p] = gn0; dL_dallmap[3 * N + p] = gn1; dL_dallmap[4 * N + p] = gn2;
    dL_dallmap[5 * N + p] = g5;
#pragma unroll
    for (int k = 6; k < 11; k++) dL_dallmap[k * N + p] = 0.f;
}

}  // namespace gsr

extern "C" int gsr_surfel_post_forward(int height, int width, const float* allmap, const float* cam21, float depth_ratio,
                                       float* render_normal, float* surf_depth, float* surf_normal, void* stream_v) {
    using namespace gsr;
    if (height <= 0 || width <= 0 || !allmap || !cam21 || !render_normal || !surf_depth || !surf_normal) {
        set_error("gsr_surfel_post_forward: invalid argument");
        return GSR_E_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream_v;
    const size_t N = (size_t)height * width;
    const unsigned blocks = (unsigned)((N + 255) / 256);
    post_fwd_a<<<blocks, 256, 0, s>>>(height, width, allmap, cam21, depth_ratio, render_normal, surf_depth);
    post_fwd_b<<<blocks, 256, 0, s>>>(height, width, allmap, cam21, surf_depth, surf_normal);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}

extern "C" int gsr_surfel_post_backward(int height, int width, const float* allmap, const float* cam21, float depth_ratio,
                                        const float* surf_depth, const float* g_render_normal, const float* g_surf_depth,
                                        const float* g_surf_normal, float* scratch6, float* dL_dallmap, void* stream_v) {
    using namespace gsr;
    if (height <= 0 || width <= 0 || !allmap || !cam21 || !surf_depth || !scratch6 || !dL_dallmap) {
        set_error("gsr_surfel_post_backward: invalid argument");
        return GSR_E_INVALID;
    }
    cudaStream_t s = (cudaStream_t)stream_v;
    const size_t N = (size_t)height * width;
    const unsigned blocks = (unsigned)((N + 255) / 256);
    if (g_surf_normal) post_bwd_a<<<blocks, 256, 0, s>>>(height, width, allmap, cam21, surf_depth, g_surf_normal, scratch6);
    else GSR_CUDA_CHECK(cudaMemsetAsync(scratch6, 0, 6 * N * sizeof(float), s));
    post_bwd_b<<<blocks, 256, 0, s>>>(height, width, allmap, cam21, depth_ratio, scratch6, g_render_normal, g_surf_depth, dL_dallmap);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}
