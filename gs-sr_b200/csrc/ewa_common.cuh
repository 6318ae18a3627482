// ewa_common.cuh -- layouts shared by the EWA-splat kernels: the 3DGS rasterizer
// (diff_gaussian_rasterization, reference G/ = submodules/diff-gaussian-rasterization) and its PGSR
// superset (diff_plane_rasterization, L/ = submodules/diff-plane-rasterization: + all_map blend, plane
// depth, out_observe, dL_dmean2D_abs), plus scaffold_filter's visible_filter (F/).
#pragma once
#include "common.cuh"

namespace gsr {

constexpr int NUM_ALL_MAP = 5;   // L/cuda_rasterizer/config.h:16

// Per-Gaussian record written by the forward preprocess (32 bytes, two 128-bit words).
struct __align__(16) EwaGeom {
    float4 a;   // screen x, y (pixels), conic.x, conic.y          (G/forward.cu:209-252)
    float4 b;   // conic.z, opacity, tau (contribution cutoff, see cull.cuh), mode (CULL_*)
};

// Record planes of the sorted per-tile lists (float4 SoA like the surfel planes, tile-local xy):
//   plane 0: x', y', conic.x, conic.y       plane 2: colour r, g, b, all_map[4]
//   plane 1: conic.z, opacity, tau, idx|flag plane 3: all_map[0..3]   (render_geo only)
constexpr int EWA_PLANES = 3, EWA_PLANES_GEO = 4;

// Per-Gaussian backward accumulator (16 floats = one 64-byte line):
//   [0..1] dL/dmean2D (already scaled by 0.5 W, 0.5 H)   [2..4] dL/dconic x, y, w   [5] dL/dopacity
//   [6..8] dL/dcolour   [9..10] |dL/dmean2D| sums (plane)   [11..15] dL/dall_map (plane, render_geo)
constexpr int EWA_GACC = 16;

struct EwaGeomWs {   // geometryBuffer
    EwaGeom* geom;       // P
    CullRec* cull;       // P
    float* depths;       // P
    uint32_t* masks;     // P
    float* rgb;          // 3P (SH path)
    uint8_t* clamped;    // 3P
    int* flags;          // [0] prefiltered violation
    static size_t carve(EwaGeomWs& w, char* base, int P);
};

}  // namespace gsr
