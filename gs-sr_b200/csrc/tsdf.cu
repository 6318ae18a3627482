// tsdf.cu -- multi-view TSDF / colour fusion of rendered depth + RGB maps on a set of sample
// points (the voxel lattice that GS-SR hands to marching cubes), for sm_100a.
//
// Result contract = GaussianExtractor.extract_mesh_unbounded's in-repo torch rule
// (/root/reference/gssr/utils/mesh_utils.py):
//   contract / uncontract                :187-193
//   compute_sdf_perframe                 :195-207  project with full_proj_transform (row-vector),
//                                                  mask_proj = |ndc.xy| < 1 & z > 0, bilinear grid_sample
//                                                  (border padding, align_corners=True) of depth and RGB,
//                                                  sdf = depth - z
//   compute_unbounded_tsdf               :209-246  adaptive truncation 5*voxel/(2 - min(|x|,1.9)) outside
//                                                  the unit ball, running mean
//                                                  tsdf <- (tsdf*w + clamp(sdf/trunc,-1,1))/(w+1) (same for
//                                                  RGB) over the views with sdf > -trunc, w0 = 1, tsdf0 = 1
// The reference runs ~25 full-array torch passes PER VIEW over the 16.7 M-point chunk
// (cat, matmul, divide, masks, two grid_samples, masked gathers/scatters of tsdf / rgb / weights).
//
// B200 design: ONE pass over the samples for ALL views.  One thread owns one sample (consecutive
// threads = consecutive lattice points along the fastest axis, so a warp's projections fall on
// neighbouring pixels and its 4-tap gathers hit the same few 128-byte lines); tsdf, weight and RGB
// stay in registers across the view loop and are written once; the per-view constants (projection,
// image size, map pointers: 96 bytes) are staged 64 views at a time into shared memory and read as
// warp-uniform broadcasts; depth / RGB maps are read through the read-only path and live in the
// 126 MB L2 (6.8 MB per 1600x1060 depth map).  HBM traffic per sample: 12 B in, 4 (+12) B out,
// independent of the number of views.
#include "common.cuh"
#include "../../include/gsr_b200.h"

namespace gsr {

constexpr int TSDF_VCHUNK = 64;
constexpr int TSDF_THREADS = 256;

struct __align__(16) TsdfViewDev {
    float m[16];          // full_proj_transform, row-major (4,4), row-vector convention: clip = [x y z 1] @ M
    int W, H;
    int pad0, pad1;
    const float* depth;   // (H, W)
    const float* rgb;     // (3, H, W) or NULL
};
static_assert(sizeof(TsdfViewDev) == 96, "TsdfViewDev must stay 96 bytes (matches gsr_tsdf_view)");

// torch grid_sample, bilinear, padding_mode='border', align_corners=True
// (aten/src/ATen/native/cuda/GridSampler.cuh: grid_sampler_unnormalize, clip_coordinates).
struct BilinearTaps {
    int x0, y0, x1, y1;
    float nw, ne, sw, se;
    bool x1_in, y1_in;
};

__device__ __forceinline__ BilinearTaps bilinear_taps(float gx, float gy, int W, int H) {
    BilinearTaps t;
    float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
    float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
    ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    t.x0 = (int)fx0; t.y0 = (int)fy0; t.x1 = t.x0 + 1; t.y1 = t.y0 + 1;
    const float fx1 = fx0 + 1.f, fy1 = fy0 + 1.f;
    t.nw = (fx1 - ix) * (fy1 - iy);
    t.ne = (ix - fx0) * (fy1 - iy);
    t.sw = (fx1 - ix) * (iy - fy0);
    t.se = (ix - fx0) * (iy - fy0);
    t.x1_in = t.x1 < W;
    t.y1_in = t.y1 < H;
    return t;
}

__device__ __forceinline__ float sample_plane(const float* __restrict__ img, int W, const BilinearTaps& t) {
    const float* r0 = img + (size_t)t.y0 * W;
    float out = __ldg(r0 + t.x0) * t.nw;
    if (t.x1_in) out += __ldg(r0 + t.x1) * t.ne;
    if (t.y1_in) {
        const float* r1 = r0 + W;
        out += __ldg(r1 + t.x0) * t.sw;
        if (t.x1_in) out += __ldg(r1 + t.x1) * t.se;
    }
    return out;
}

// GRID: the samples are the lattice points  origin + (ix, iy, iz) * voxel_size  of a bounded nx x ny x nz volume (x fastest),
// generated from the thread index instead of read from memory; fixed truncation `grid_trunc`; a view is skipped where its
// sampled depth is <= 0 (masked background, mesh_utils.py:160-162) or > depth_trunc (the depth_trunc of
// RGBDImage.create_from_color_and_depth, :165-170).  This is the bounded fusion of extract_mesh_bounded (:138-179) with
// the reference's own torch integration rule in place of Open3D's ScalableTSDFVolume (absent dependency, parity unpinned).
//
// GRID traversal: thread blocks own 8 x 8 x 4 bricks (a warp = an 8 x 4 patch of one z slice) and walk the lattice one
// 64^3 macro-block after the other.  In lattice order (one z slice after the other) every slice projects onto most of every
// view's map, 32 maps are 217 MB > L2, and each slice re-streamed them from DRAM: 512^3 x 32 views took 27 ms, DRAM-bound
// on traffic ~50x the algorithmic bytes.  The blocks running at one time now cover a compact piece of space whose
// projections (a few 100 KB per view) stay in L2; a map region is re-fetched at most once per macro-block along its rays.
struct TsdfGrid { float ox, oy, oz, trunc, depth_trunc; int nx, ny, nz, mgx, mgy; };
constexpr int TSDF_MACRO = 64;                       // macro-block edge in voxels
constexpr int TSDF_BRICKS_PER_MACRO = (TSDF_MACRO / 8) * (TSDF_MACRO / 8) * (TSDF_MACRO / 4);   // 1024

template <bool RGB, bool GRID>
__global__ void __launch_bounds__(TSDF_THREADS)
tsdf_fuse_kernel(long long n, const float* __restrict__ samples, int contracted, float cx, float cy, float cz,
                 float radius, float voxel_size, int nviews, const TsdfViewDev* __restrict__ views, int init,
                 float* __restrict__ tsdf_io, float* __restrict__ weight_io, float* __restrict__ rgb_io, const TsdfGrid grid) {
    __shared__ TsdfViewDev sviews[TSDF_VCHUNK];
    long long i = (long long)blockIdx.x * TSDF_THREADS + threadIdx.x;
    bool live = i < n;

    float x = 0.f, y = 0.f, z = 0.f, trunc = 5.f * voxel_size;
    float tsdf = 1.f, w = 1.f, r = 0.f, g = 0.f, b = 0.f;
    if (GRID) {
        const unsigned macro = blockIdx.x / TSDF_BRICKS_PER_MACRO, brick = blockIdx.x % TSDF_BRICKS_PER_MACRO;
        const unsigned mx = macro % (unsigned)grid.mgx, mt = macro / (unsigned)grid.mgx;
        const unsigned my = mt % (unsigned)grid.mgy, mz = mt / (unsigned)grid.mgy;
        const int ix = (int)(mx * TSDF_MACRO + (brick & 7u) * 8u + (threadIdx.x & 7u));
        const int iy = (int)(my * TSDF_MACRO + ((brick >> 3) & 7u) * 8u + ((threadIdx.x >> 3) & 7u));
        const int iz = (int)(mz * TSDF_MACRO + (brick >> 6) * 4u + (threadIdx.x >> 6));
        live = ix < grid.nx && iy < grid.ny && iz < grid.nz;
        i = ((long long)iz * grid.ny + iy) * grid.nx + ix;
        x = fmaf((float)ix, voxel_size, grid.ox);
        y = fmaf((float)iy, voxel_size, grid.oy);
        z = fmaf((float)iz, voxel_size, grid.oz);
        trunc = grid.trunc;
    }
    if (live) {
        if (!GRID) { x = samples[3 * i]; y = samples[3 * i + 1]; z = samples[3 * i + 2]; }
        if (!GRID && contracted) {
            // mesh_utils.py:215-219 (adaptive truncation) and :190-193, :248-250 (uncontract, unnormalize)
            const float mag = sqrtf(x * x + y * y + z * z);
            if (mag > 1.f) trunc *= 1.f / (2.f - fminf(mag, 1.9f));
            if (!(mag < 1.f)) {
                const float s = 1.f / (2.f - mag);
                x = s * (x / mag); y = s * (y / mag); z = s * (z / mag);
            }
            x = x * radius + cx; y = y * radius + cy; z = z * radius + cz;
        }
        if (!init) {
            tsdf = tsdf_io[i]; w = weight_io[i];
            if (RGB) { r = rgb_io[3 * i]; g = rgb_io[3 * i + 1]; b = rgb_io[3 * i + 2]; }
        }
    }

    const float itrunc = __frcp_rn(trunc);
    for (int v0 = 0; v0 < nviews; v0 += TSDF_VCHUNK) {
        const int nv = min(TSDF_VCHUNK, nviews - v0);
        __syncthreads();
        {   // stage nv view descriptors (96 B each = 6 x 16 B) with coalesced 128-bit copies
            const float4* src = reinterpret_cast<const float4*>(views + v0);
            float4* dst = reinterpret_cast<float4*>(sviews);
            for (int k = threadIdx.x; k < nv * 6; k += TSDF_THREADS) dst[k] = __ldg(src + k);
        }
        __syncthreads();
        if (!live) continue;
        for (int v = 0; v < nv; v++) {
            const TsdfViewDev& V = sviews[v];
            const float hx = x * V.m[0] + y * V.m[4] + z * V.m[8] + V.m[12];
            const float hy = x * V.m[1] + y * V.m[5] + z * V.m[9] + V.m[13];
            const float hw = x * V.m[3] + y * V.m[7] + z * V.m[11] + V.m[15];
            // one correctly rounded reciprocal instead of two IEEE divisions (u, t within 1 ulp of hx / hw, hy / hw: a
            // 1e-4 pixel shift of the bilinear sample at most; the kernel is instruction-bound)
            const float ihw = __frcp_rn(hw);
            const float u = hx * ihw, t = hy * ihw;
            const bool mask_proj = (u > -1.f) && (u < 1.f) && (t > -1.f) && (t < 1.f) && (hw > 0.f);
            if (!mask_proj) continue;
            const BilinearTaps taps = bilinear_taps(u, t, V.W, V.H);
            const float dsamp = sample_plane(V.depth, V.W, taps);
            if (GRID && (!(dsamp > 0.f) || dsamp > grid.depth_trunc)) continue;
            const float sdf = dsamp - hw;
            if (!(sdf > -trunc)) continue;
            const float s = fminf(fmaxf(sdf * itrunc, -1.f), 1.f);
            const float wp = w + 1.f, iwp = __frcp_rn(wp);
            tsdf = (tsdf * w + s) * iwp;
            if (RGB) {
                const size_t plane = (size_t)V.W * V.H;
                r = (r * w + sample_plane(V.rgb, V.W, taps)) * iwp;
                g = (g * w + sample_plane(V.rgb + plane, V.W, taps)) * iwp;
                b = (b * w + sample_plane(V.rgb + 2 * plane, V.W, taps)) * iwp;
            }
            w = wp;
        }
    }
    if (live) {
        tsdf_io[i] = tsdf;
        if (weight_io) weight_io[i] = w;
        if (RGB) { rgb_io[3 * i] = r; rgb_io[3 * i + 1] = g; rgb_io[3 * i + 2] = b; }
    }
}

}  // namespace gsr

extern "C" int gsr_tsdf_fuse(long long n, const float* samples, int contracted, const float* center, float radius,
                             float voxel_size, int nviews, const gsr_tsdf_view* views, int init, float* tsdf,
                             float* weights, float* rgb, void* stream_v) {
    using namespace gsr;
    static_assert(sizeof(gsr_tsdf_view) == sizeof(TsdfViewDev), "gsr_tsdf_view layout");
    cudaStream_t s = (cudaStream_t)stream_v;
    if (n < 0 || nviews < 0 || (n > 0 && (!samples || !tsdf)) || (nviews > 0 && !views) || !(voxel_size > 0.f) ||
        (contracted && !(radius > 0.f)) || (!init && !weights)) {
        set_error("gsr_tsdf_fuse: invalid argument");
        return GSR_E_INVALID;
    }
    if (n == 0) return GSR_OK;
    const float cx = center ? center[0] : 0.f, cy = center ? center[1] : 0.f, cz = center ? center[2] : 0.f;
    const long long blocks = (n + TSDF_THREADS - 1) / TSDF_THREADS;
    if (blocks > 0x7fffffffLL) { set_error("gsr_tsdf_fuse: too many samples (%lld)", n); return GSR_E_OVERFLOW; }
    const TsdfViewDev* vd = reinterpret_cast<const TsdfViewDev*>(views);
    const TsdfGrid nogrid = {0.f, 0.f, 0.f, 0.f, 0.f, 0, 0, 0, 1, 1};
    if (rgb)
        tsdf_fuse_kernel<true, false><<<(unsigned)blocks, TSDF_THREADS, 0, s>>>(n, samples, contracted, cx, cy, cz, radius,
                                                                                voxel_size, nviews, vd, init, tsdf, weights, rgb, nogrid);
    else
        tsdf_fuse_kernel<false, false><<<(unsigned)blocks, TSDF_THREADS, 0, s>>>(n, samples, contracted, cx, cy, cz, radius,
                                                                                 voxel_size, nviews, vd, init, tsdf, weights, nullptr, nogrid);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}

extern "C" int gsr_tsdf_integrate_grid(int nx, int ny, int nz, const float* origin, float voxel_size, float sdf_trunc,
                                       float depth_trunc, int nviews, const gsr_tsdf_view* views, int init, float* tsdf,
                                       float* weights, float* rgb, void* stream_v) {
    using namespace gsr;
    cudaStream_t s = (cudaStream_t)stream_v;
    if (nx <= 0 || ny <= 0 || nz <= 0 || !origin || !(voxel_size > 0.f) || !(sdf_trunc > 0.f) || !(depth_trunc > 0.f) || nviews < 0 ||
        (nviews > 0 && !views) || !tsdf || !weights) {
        set_error("gsr_tsdf_integrate_grid: invalid argument");
        return GSR_E_INVALID;
    }
    static_assert(TSDF_THREADS == 8 * 8 * 4, "a thread block is one 8 x 8 x 4 brick");
    const long long n = (long long)nx * ny * nz;
    const int mgx = (nx + TSDF_MACRO - 1) / TSDF_MACRO, mgy = (ny + TSDF_MACRO - 1) / TSDF_MACRO, mgz = (nz + TSDF_MACRO - 1) / TSDF_MACRO;
    const long long blocks = (long long)mgx * mgy * mgz * TSDF_BRICKS_PER_MACRO;
    if (blocks > 0x7fffffffLL) { set_error("gsr_tsdf_integrate_grid: volume too large (%lld voxels)", n); return GSR_E_OVERFLOW; }
    const TsdfGrid g = {origin[0], origin[1], origin[2], sdf_trunc, depth_trunc, nx, ny, nz, mgx, mgy};
    const TsdfViewDev* vd = reinterpret_cast<const TsdfViewDev*>(views);
    if (rgb)
        tsdf_fuse_kernel<true, true><<<(unsigned)blocks, TSDF_THREADS, 0, s>>>(n, nullptr, 0, 0.f, 0.f, 0.f, 1.f, voxel_size, nviews, vd,
                                                                               init, tsdf, weights, rgb, g);
    else
        tsdf_fuse_kernel<false, true><<<(unsigned)blocks, TSDF_THREADS, 0, s>>>(n, nullptr, 0, 0.f, 0.f, 0.f, 1.f, voxel_size, nviews, vd,
                                                                                init, tsdf, weights, nullptr, g);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}
