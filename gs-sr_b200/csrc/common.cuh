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
    float4 nd;  // view-space normal (facing the camera),           view depth
};

// Per-(Gaussian, tile) record, materialised in sorted (tile, depth, index) order
// so that each tile's list is one contiguous 80-byte-stride stream that a CTA
// pulls into shared memory with cp.async.bulk.  Homography rows and the filter
// centre are re-expressed in TILE-LOCAL pixel coordinates (origin = tile's first pixel).
struct __align__(16) SplatRec {
    float4 tu;   // Tu - ox*Tw,   cx - ox
    float4 tv;   // Tv - oy*Tw,   cy - oy
    float4 tw;   // Tw,           opacity
    float4 ng;   // normal.xyz,   Gaussian index (int bits)
    float4 cb;   // colour rgb,   packed conservative pixel bounds (int bits, see pack_bounds)
};
static_assert(sizeof(SplatRec) == 80, "SplatRec must be 80 bytes");

// Per-Gaussian gradient accumulator filled by the backward render (float atomics
// after in-warp reduction), consumed by the backward preprocess.  80 bytes.
//   [0..8] dL/dT (Tu,Tv,Tw)  [9..11] dL/dcolour  [12..14] dL/dnormal  [15] dL/dopacity
//   [16..17] dL/d(filter centre)  [18..19] pad
constexpr int GACC_STRIDE = 20;

// bounds: local pixel range [xmin,xmax] x [ymin,ymax] (each 0..15) outside of which
// the splat provably cannot reach alpha >= 1/255; bit 16 set = empty.
__host__ __device__ inline uint32_t pack_bounds(int xmin, int xmax, int ymin, int ymax) {
    return (uint32_t)xmin | ((uint32_t)xmax << 4) | ((uint32_t)ymin << 8) | ((uint32_t)ymax << 12);
}
constexpr uint32_t BOUNDS_EMPTY = 1u << 16;
constexpr uint32_t BOUNDS_FULL = 0u | (15u << 4) | (0u << 8) | (15u << 12);

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

struct GeomWs {      // geometryBuffer
    GeomRec* geom;       // P
    float4* cbox;        // P  conservative contribution box, global pixel coords (xmin,xmax,ymin,ymax)
    uint32_t* tiles;     // P  tiles touched
    uint32_t* offsets;   // P  inclusive scan
    float* rgb;          // 3P (SH path only; else unused)
    uint8_t* clamped;    // 3P
    int* flags;          // [0] prefiltered violation, [1..] reserved
    char* scan_tmp;
    size_t scan_tmp_bytes;
    static size_t carve(GeomWs& w, char* base, int P, size_t scan_tmp_bytes);
};

struct ImageWs {     // imageBuffer
    float* final_T;      // 3N: T, M1, M2   (S/cuda_rasterizer/forward.cu:429-437)
    uint32_t* n_contrib; // 2N: last contributor, median contributor
    uint2* ranges;       // tiles
    static size_t carve(ImageWs& w, char* base, int W, int H);
};

struct BinWs {       // binningBuffer
    uint64_t* keys_unsorted;
    uint64_t* keys;
    uint32_t* vals_unsorted;
    uint32_t* vals;
    SplatRec* recs;      // R
    float* gacc;         // P * GACC_STRIDE (backward accumulators; lives here so forward owns one blob)
    char* sort_tmp;
    size_t sort_tmp_bytes;
    static size_t carve(BinWs& w, char* base, int64_t R, int P, size_t sort_tmp_bytes);
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
