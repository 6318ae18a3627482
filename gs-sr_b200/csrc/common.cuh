// common.cuh -- shared definitions for libgsr_b200 (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#ifndef __CUDA_ARCH__
#define GSR_HOST_ONLY 1
#endif
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgsr_b200 is written for sm_100a (B200) only"
#endif

namespace gsr {

constexpr int TILE = 16;             // reference BLOCK_X/BLOCK_Y (S/cuda_rasterizer/config.h:16-17)
constexpr int TILE_PIX = TILE * TILE;
constexpr float NEAR_N = 0.2f;       // S/cuda_rasterizer/auxiliary.h:39-42
constexpr float FAR_N = 100.0f;
constexpr float FILTER_SIZE = 0.707106f;
constexpr float FILTER_INV_SQUARE = 2.0f;
constexpr float ALPHA_MIN = 1.0f / 255.0f;
constexpr float ALPHA_MAX = 0.99f;
constexpr float T_EPS = 0.0001f;

// ---- error plumbing (no exceptions across the C ABI) ----------------------
void set_error(const char* fmt, ...);
#define GSR_CUDA_CHECK(expr)                                                          \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            gsr::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,       \
                           cudaGetErrorString(_e));                                   \
            return GSR_E_CUDA;                                                        \
        }                                                                             \
    } while (0)

// ---- HBM layouts -----------------------------------------------------------
// Per-Gaussian geometry record written by the forward preprocess: one 64-byte
// line per Gaussian (four 128-bit stores / loads, always 16-byte aligned).
struct __align__(16) GeomRec {
    float4 tu;  // Tu.xyz (x*w row of the splat->pixel homography), aabb centre x
    float4 tv;  // Tv.xyz,                                          aabb centre y
    float4 tw;  // Tw.xyz,                                          opacity
    float4 nd;  // view-space normal (facing the camera),           tau (>= 0) or -(tau+1): "always evaluate"
};

// Per-Gaussian culling record (forward preprocess -> tile binning / record build).
// The conic q (see cull.cuh) is expressed in a frame shifted to the splat's rounded
// screen centre (sx, sy) so that its float32 coefficients stay well conditioned.
struct __align__(16) CullRec {
    float4 q0;  // xx, xy, yy, bx
    float4 q1;  // by, c0, disc radius^2 (= tau/2), tau
    float4 q2;  // sx, sy, mode (0 exact, 1 always evaluate, 2 never), unused
};
constexpr int CULL_EXACT = 0, CULL_ALWAYS = 1, CULL_NEVER = 2;

// Per-(Gaussian, tile) record stream, materialised in sorted (tile, depth, index) order as
// SIX float4 planes (SoA of 16-byte words): each tile's list is a contiguous run in every
// plane, pulled into shared memory with cp.async.bulk, and lane j's 128-bit read of entry j
// is bank-conflict free.  All geometry is in TILE-LOCAL pixel coordinates (origin = the
// tile's first pixel): with Tu' = Tu - ox Tw, Tv' = Tv - oy Tw,
//     p(x,y) = k x l = a x + b y + c,  a = Tv' x Tw,  b = Tw x Tu',  c = Tu' x Tv',
//     s = p.xy / p.z,   depth = s.Tw.xy + Tw.z = det(T) / p.z.
//   plane 0 (QA): a.xyz, cx'           plane 3 (QD): det(T), tau, Tw.z, index | flag<<31
//   plane 1 (QB): b.xyz, cy'           plane 4 (PN): normal.xyz, colour.r
//   plane 2 (QC): c.xyz, opacity       plane 5 (PC): colour.g, colour.b, sx', sy' (moment origin)
constexpr int REC_PLANES = 6;
constexpr uint32_t REC_FLAG_ALWAYS = 0x80000000u;   // conic is not an ellipse: evaluate on every pixel
// When P < 2^23 the Gaussian index needs 23 bits and bits 23..30 of the same record word carry, per warp block of the
// tile (8 blocks of 8x4 pixels), "the forward blended this entry on at least one pixel of the block" (set with
// red.global.or by surfel_render_fwd, read back through the record stream by surfel_render_bwd, which then walks
// exactly the contributing entries instead of repeating the cull test).
constexpr int REC_USED_SHIFT = 23;
constexpr uint32_t REC_INDEX_MASK_USED = (1u << REC_USED_SHIFT) - 1u;

// Per-Gaussian gradient accumulator filled by the backward render (float atomics after the
// in-warp reduction), consumed by the backward preprocess.  80 bytes:
//   [0..2] M0 = sum dL/dp      [3..5] MX = sum x~ dL/dp    [6..8] MY = sum y~ dL/dp   (x~,y~ from the moment origin)
//   [9] dL/d det(T)            [10..12] dL/dcolour         [13..15] dL/dnormal
//   [16] dL/dopacity           [17] dL/dTw.z (low-pass branch)   [18..19] dL/d(filter centre)
constexpr int GACC_STRIDE = 20;

// ---- workspace carving (128-byte aligned sub-allocations of one blob) -------
struct Carver {
    char* p;
    size_t used = 0;
    explicit Carver(char* base) : p(base) {}
    template <typename T>
    T* take(size_t count) {
        size_t off = (used + 127) & ~size_t(127);
        used = off + count * sizeof(T);
        return p ? reinterpret_cast<T*>(p + off) : nullptr;
    }
};

// Per-tile atomic counters (count pass of the preprocess, cursors of the scatter) are spaced one 128-byte line apart:
// the L2 atomic unit serialises per line, and 32 neighbouring tiles sharing a line made the two passes wait on
// ~30 k queued atomics per line (scatter_keys: 88 % long-scoreboard stalls, 12 % issue).
constexpr int TILE_CTR_STRIDE = 32;
constexpr uint32_t MASK_RETEST = 0xffffffffu;   // rect larger than 32 tiles: scatter re-runs the tile test

struct GeomWs {      // geometryBuffer
    GeomRec* geom;       // P
    CullRec* cull;       // P
    float* depths;       // P  view depth (sort key bits)
    uint32_t* masks;     // P  bit k set: k-th tile (row-major) of the getRect rectangle is reachable
    float* rgb;          // 3P (SH path only; else unused)
    uint8_t* clamped;    // 3P
    int* flags;          // [0] prefiltered violation, [1..] reserved
    static size_t carve(GeomWs& w, char* base, int P);
};

struct ImageWs {     // imageBuffer
    float* final_T;      // 3N: T, M1, M2   (S/cuda_rasterizer/forward.cu:429-437)
    uint32_t* n_contrib; // 2N: last contributor, median contributor
    uint32_t* tile_count;   // tiles x TILE_CTR_STRIDE: counter of tile t at [t * TILE_CTR_STRIDE]
    uint32_t* tile_offset;  // tiles + 1: tile t owns list entries [offset[t], offset[t+1])
    uint32_t* tile_cursor;  // tiles x TILE_CTR_STRIDE: scatter cursors, same spacing
    uint32_t* total;        // [0] = R, [1] = prefiltered-violation flag (written by tile_scan, read back by the host)
    // nT / nC: per-pixel float / uint32 planes (surfel 3 + 2, EWA 1 + 1)
    static size_t carve(ImageWs& w, char* base, int W, int H, int nT = 3, int nC = 2);
};

struct BinWs {       // binningBuffer
    uint64_t* keys;      // R  (depth_bits << 32 | index), bucketed by tile
    float4* planes;      // REC_PLANES x Rpad float4 (plane stride = Rpad)
    size_t plane_stride; // Rpad = R rounded up to a multiple of 8 (keeps every plane 128-byte aligned)
    float* gacc;         // P * GACC_STRIDE (backward accumulators; lives here so forward owns one blob)
    static size_t carve(BinWs& w, char* base, int64_t R, int P, int nplanes = REC_PLANES, int gacc_stride = GACC_STRIDE);
};

// Per-view constants.  view/proj/campos stay DEVICE pointers exactly as the reference
// API hands them over (no host read-back, no extra sync); kernels load the 35 floats
// through the read-only path (L1 broadcast).
struct ViewParams {
    const float* view;    // (4,4) row-vector convention, element (r,c) of the column-vector matrix at [4c+r]
    const float* proj;    // full projection = view @ proj
    const float* campos;  // (3) or NULL
    int W, H, gx, gy;
    float scale_modifier;
};

// ---- small device math ------------------------------------------------------
#ifdef __CUDACC__
// Gaussian handled by this thread in the per-Gaussian binning kernels (256 threads per CTA): warp w of CTA b takes the 32
// CONSECUTIVE Gaussians of global warp w * gridDim.x + b, so loads stay coalesced while a contiguous run of big splats
// (e.g. the children a densification step appends at the end of the arrays) is spread over as many CTAs as it has
// warps -- with thread i <-> Gaussian i, 500 screen-filling splats in a row were the work of two CTAs (2.4 ms).
__device__ __forceinline__ int spread_gaussian_index() {
    return (int)(((threadIdx.x >> 5) * gridDim.x + blockIdx.x) * 32u + (threadIdx.x & 31u));
}
__device__ __forceinline__ float3 xform43(const float* __restrict__ m, float3 p) {
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
                       m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
                       m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14]);
}
__device__ __forceinline__ float3 xformvec43(const float* __restrict__ m, float3 p) {
    return make_float3(m[0] * p.x + m[4] * p.y + m[8] * p.z,
                       m[1] * p.x + m[5] * p.y + m[9] * p.z,
                       m[2] * p.x + m[6] * p.y + m[10] * p.z);
}
__device__ __forceinline__ float3 xformvec43T(const float* __restrict__ m, float3 p) {
    return make_float3(m[0] * p.x + m[1] * p.y + m[2] * p.z,
                       m[4] * p.x + m[5] * p.y + m[6] * p.z,
                       m[8] * p.x + m[9] * p.y + m[10] * p.z);
}
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

// Origin of the backward's moment frame along one axis: the splat's screen centre, clamped to
// the image and rounded (exactly reproducible in build_records and preprocess_bwd).
__device__ __forceinline__ float moment_origin(float c, int extent) {
    return rintf(fminf(fmaxf(c, 0.f), (float)(extent - 1)));
}

// S/cuda_rasterizer/auxiliary.h:69-79 (getRect); r is the integer radius.
__device__ __forceinline__ void get_rect(float px, float py, int r, int gx, int gy, int& x0, int& y0,
                                         int& x1, int& y1) {
    x0 = min(gx, max(0, (int)((px - r) / TILE)));
    y0 = min(gy, max(0, (int)((py - r) / TILE)));
    x1 = min(gx, max(0, (int)((px + r + TILE - 1) / TILE)));
    y1 = min(gy, max(0, (int)((py + r + TILE - 1) / TILE)));
}
#endif  // __CUDACC__

}  // namespace gsr
