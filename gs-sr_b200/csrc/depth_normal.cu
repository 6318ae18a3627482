// depth_normal.cu -- PGSR's normal_from_depth_image as one streaming kernel pair.
//
// Result contract = /root/reference/gssr/utils/graphics_utils.py:139-146 with offset=None (the only way GS-SR calls
// it, gssr/scene/pgsr_scene.py:227-238,320):
//   depth2point_cam / ndc_2_cam (:79-99): camera-space point of pixel (x, y) = [x z, y z, z] @ inverse(K^T)
//   depth_pcd2normal (:110-137):         n = normalize(cross(P(y, x+1) - P(y, x-1), P(y-1, x) - P(y+1, x))) on the interior,
//                                        zero on the one-pixel border (F.normalize: v / max(|v|, 1e-12)).
// The 3x3 inverse(K^T) is taken by the caller with the same float32 torch op as the reference and handed over as 9
// floats ON THE DEVICE (row-major, Ki[r][c]; no host read-back); the point is formed in the matmul's order
// x z Ki[0][c] + y z Ki[1][c] + z Ki[2][c].
//
// B200: pure HBM streaming.  Forward: 4 B in (the depth map; the 5-point stencil is served by L1/L2) + 12 B out per
// pixel; optional per-pixel weight (PGSR multiplies by the detached alpha, pgsr_scene.py:320) fused in.  Backward: the
// per-pixel adjoints of the two difference vectors (24 B) are written by pass 1 and gathered by pass 2 from the four
// neighbours (no atomics, deterministic), which folds them through the point map into dL/ddepth.
#include "common.cuh"

namespace gsr {

struct DnCam { float k[9]; };   // inverse(K^T), row-major
__device__ __forceinline__ DnCam dn_load(const float* __restrict__ kinv) {
    DnCam c;
#pragma unroll
    for (int i = 0; i < 9; i++) c.k[i] = __ldg(kinv + i);
    return c;
}

__device__ __forceinline__ float3 dn_point(const DnCam& c, float x, float y, float z) {
    const float xz = x * z, yz = y * z;
    return make_float3(fmaf(z, c.k[6], fmaf(yz, c.k[3], xz * c.k[0])), fmaf(z, c.k[7], fmaf(yz, c.k[4], xz * c.k[1])),
                       fmaf(z, c.k[8], fmaf(yz, c.k[5], xz * c.k[2])));
}

__global__ void __launch_bounds__(256)
depth_normal_fwd(int H, int W, const float* __restrict__ depth, const float* __restrict__ kinv, const float* __restrict__ weight,
                 float* __restrict__ normal /* (3,H,W) */) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const DnCam cam = dn_load(kinv);
    const size_t N = (size_t)W * H, pid = (size_t)y * W + x;
    float3 n = make_float3(0.f, 0.f, 0.f);
    if (x > 0 && x < W - 1 && y > 0 && y < H - 1) {
        const float3 r = dn_point(cam, (float)(x + 1), (float)y, __ldg(depth + pid + 1));
        const float3 l = dn_point(cam, (float)(x - 1), (float)y, __ldg(depth + pid - 1));
        const float3 t = dn_point(cam, (float)x, (float)(y - 1), __ldg(depth + pid - W));
        const float3 b = dn_point(cam, (float)x, (float)(y + 1), __ldg(depth + pid + W));
        const float3 u = make_float3(r.x - l.x, r.y - l.y, r.z - l.z), v = make_float3(t.x - b.x, t.y - b.y, t.z - b.z);
        const float3 c = cross3(u, v);
        const float inv = 1.0f / fmaxf(sqrtf(dot3(c, c)), 1e-12f);
        n = make_float3(c.x * inv, c.y * inv, c.z * inv);
        if (weight != nullptr) { const float w = __ldg(weight + pid); n.x *= w; n.y *= w; n.z *= w; }
    }
    normal[pid] = n.x; normal[pid + N] = n.y; normal[pid + 2 * N] = n.z;
}

// pass 1: adjoints of u = right - left and v = top - bottom at every pixel (zero on the border)
__global__ void __launch_bounds__(256)
depth_normal_bwd_adjoint(int H, int W, const float* __restrict__ depth, const float* __restrict__ kinv,
                         const float* __restrict__ weight, const float* __restrict__ g_normal, float* __restrict__ adj /* (6,H,W) */) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const DnCam cam = dn_load(kinv);
    const size_t N = (size_t)W * H, pid = (size_t)y * W + x;
    float3 du = make_float3(0.f, 0.f, 0.f), dv = du;
    if (x > 0 && x < W - 1 && y > 0 && y < H - 1) {
        const float3 r = dn_point(cam, (float)(x + 1), (float)y, __ldg(depth + pid + 1));
        const float3 l = dn_point(cam, (float)(x - 1), (float)y, __ldg(depth + pid - 1));
        const float3 t = dn_point(cam, (float)x, (float)(y - 1), __ldg(depth + pid - W));
        const float3 b = dn_point(cam, (float)x, (float)(y + 1), __ldg(depth + pid + W));
        const float3 u = make_float3(r.x - l.x, r.y - l.y, r.z - l.z), v = make_float3(t.x - b.x, t.y - b.y, t.z - b.z);
        const float3 c = cross3(u, v);
        const float len = sqrtf(dot3(c, c));
        float3 g = make_float3(__ldg(g_normal + pid), __ldg(g_normal + pid + N), __ldg(g_normal + pid + 2 * N));
        if (weight != nullptr) { const float w = __ldg(weight + pid); g.x *= w; g.y *= w; g.z *= w; }
        float3 gc;                                     // adjoint of c -> c / max(|c|, eps)
        if (len > 1e-12f) {
            const float inv = 1.0f / len;
            const float3 nh = make_float3(c.x * inv, c.y * inv, c.z * inv);
            const float gd = dot3(g, nh);
            gc = make_float3((g.x - nh.x * gd) * inv, (g.y - nh.y * gd) * inv, (g.z - nh.z * gd) * inv);
        } else {
            gc = make_float3(g.x * 1e12f, g.y * 1e12f, g.z * 1e12f);
        }
        du = cross3(v, gc);                            // c = u x v:  dL/du = v x gc,  dL/dv = gc x u
        dv = cross3(gc, u);
    }
    adj[pid] = du.x; adj[pid + N] = du.y; adj[pid + 2 * N] = du.z;
    adj[pid + 3 * N] = dv.x; adj[pid + 4 * N] = dv.y; adj[pid + 5 * N] = dv.z;
}

// pass 2: P(y, x) is the right point of (y, x-1), the left point of (y, x+1), the top point of (y+1, x) and the bottom
// point of (y-1, x); dL/dP folds through  P = z (x Ki[0] + y Ki[1] + Ki[2])  into dL/dz
__global__ void __launch_bounds__(256)
depth_normal_bwd_gather(int H, int W, const float* __restrict__ kinv, const float* __restrict__ adj, float* __restrict__ g_depth) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const DnCam cam = dn_load(kinv);
    const size_t N = (size_t)W * H, pid = (size_t)y * W + x;
    float3 gp = make_float3(0.f, 0.f, 0.f);
    if (x > 0) { gp.x += __ldg(adj + pid - 1); gp.y += __ldg(adj + pid - 1 + N); gp.z += __ldg(adj + pid - 1 + 2 * N); }
    if (x < W - 1) { gp.x -= __ldg(adj + pid + 1); gp.y -= __ldg(adj + pid + 1 + N); gp.z -= __ldg(adj + pid + 1 + 2 * N); }
    if (y < H - 1) { gp.x += __ldg(adj + pid + W + 3 * N); gp.y += __ldg(adj + pid + W + 4 * N); gp.z += __ldg(adj + pid + W + 5 * N); }
    if (y > 0) { gp.x -= __ldg(adj + pid - W + 3 * N); gp.y -= __ldg(adj + pid - W + 4 * N); gp.z -= __ldg(adj + pid - W + 5 * N); }
    const float fx = (float)x, fy = (float)y;
    const float3 ray = make_float3(fmaf(fx, cam.k[0], fmaf(fy, cam.k[3], cam.k[6])), fmaf(fx, cam.k[1], fmaf(fy, cam.k[4], cam.k[7])),
                                   fmaf(fx, cam.k[2], fmaf(fy, cam.k[5], cam.k[8])));
    g_depth[pid] = dot3(gp, ray);
}

cudaError_t launch_depth_normal_fwd(int H, int W, const float* depth, const float* kinv, const float* weight, float* normal,
                                    cudaStream_t s) {
    const dim3 grid((W + 31) / 32, (H + 7) / 8);
    depth_normal_fwd<<<grid, 256, 0, s>>>(H, W, depth, kinv, weight, normal);
    return cudaGetLastError();
}

cudaError_t launch_depth_normal_bwd(int H, int W, const float* depth, const float* kinv, const float* weight,
                                    const float* g_normal, float* scratch6, float* g_depth, cudaStream_t s) {
    const dim3 grid((W + 31) / 32, (H + 7) / 8);
    depth_normal_bwd_adjoint<<<grid, 256, 0, s>>>(H, W, depth, kinv, weight, g_normal, scratch6);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    depth_normal_bwd_gather<<<grid, 256, 0, s>>>(H, W, kinv, scratch6, g_depth);
    return cudaGetLastError();
}

}  // namespace gsr
