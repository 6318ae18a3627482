// oracle/ref_driver_raster.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C driver around the UNMODIFIED reference rasterizer sources.  It is
// compiled together with the reference's own
//   submodules/<variant>/cuda_rasterizer/{forward,backward,rasterizer_impl}.cu
// (where they lie under /root/reference; nothing is copied into this repo)
// into oracle/_ref/libref_<variant>.so by oracle/build_ref.sh.
//
// It replaces only the torch glue of the reference
// (submodules/diff-surfel-rasterization/rasterize_points.cu:39-255, which
// needs ~5 min of torch headers to compile) with raw device pointers, so the
// reference kernels can be driven from ctypes / a C harness.  The three
// scratch "chunk" buffers the reference resizes through std::function
// callbacks (rasterize_points.cu:31-37) are cudaMalloc'ed here and kept in a
// handle until ref_free().
//
// One source, three variants, selected by -DREF_VARIANT_{SURFEL,GAUSSIAN,PLANE}
// because every reference submodule defines the same CudaRasterizer::Rasterizer
// symbol (one .so per variant).
//
// Users: tests/ (parity checker), bench.py --impl reference.  Never imported
// by the product path.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <cuda_runtime.h>
#include "cuda_rasterizer/config.h"
#include "cuda_rasterizer/rasterizer.h"

namespace {
struct Chunk {
    char* ptr = nullptr;
    size_t cap = 0;
    char* get(size_t n) {
        if (n > cap) {
            if (ptr) cudaFree(ptr);
            size_t want = n + (n >> 2) + 256;
            if (cudaMalloc(&ptr, want) != cudaSuccess) { ptr = nullptr; cap = 0; return nullptr; }
            cap = want;
        }
        return ptr;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};
struct RefHandle {
    Chunk geom, binning, img;
    int num_rendered = 0;
};
}  // namespace

extern "C" {

void* ref_create() { return new RefHandle(); }

void ref_free(void* h) {
    RefHandle* r = (RefHandle*)h;
    if (!r) return;
    r->geom.release(); r->binning.release(); r->img.release();
    delete r;
}

int ref_num_rendered(void* h) { return ((RefHandle*)h)->num_rendered; }

int ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present) {
    if (P == 0) return 0;
    CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
    return (int)cudaGetLastError();
}

#if defined(REF_VARIANT_SURFEL)
// mirrors RasterizeGaussiansCUDA (S/rasterize_points.cu:39-135); outputs must
// be zero-filled by the caller as torch::full(...,0) does there.
int ref_forward(void* h, int P, int D, int M, const float* bg, int W, int H,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* transMat_precomp,
                const float* viewmatrix, const float* projmatrix, const float* campos,
                float tan_fovx, float tan_fovy, int prefiltered,
                float* out_color, float* out_others, int* radii, int debug) {
    RefHandle* r = (RefHandle*)h;
    r->num_rendered = 0;
    if (P == 0) return 0;
    std::function<char*(size_t)> g = [r](size_t n) { return r->geom.get(n); };
    std::function<char*(size_t)> b = [r](size_t n) { return r->binning.get(n); };
    std::function<char*(size_t)> i = [r](size_t n) { return r->img.get(n); };
    try {
        r->num_rendered = CudaRasterizer::Rasterizer::forward(
            g, b, i, P, D, M, bg, W, H, means3D, shs, colors_precomp, opacities, scales,
            scale_modifier, rotations, transMat_precomp, viewmatrix, projmatrix, campos,
            tan_fovx, tan_fovy, prefiltered != 0, out_color, out_others, radii, debug != 0);
    } catch (...) { return -1; }
    return (int)cudaGetLastError();
}

// mirrors RasterizeGaussiansBackwardCUDA (S/rasterize_points.cu:137-234);
// all dL_* outputs must be zero-filled by the caller (torch::zeros there).
int ref_backward(void* h, int P, int D, int M, const float* bg, int W, int H,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* scales, float scale_modifier, const float* rotations,
                 const float* transMat_precomp, const float* viewmatrix,
                 const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
                 const int* radii, const float* dL_dpix, const float* dL_dothers,
                 float* dL_dmean2D, float* dL_dnormal, float* dL_dopacity, float* dL_dcolor,
                 float* dL_dmean3D, float* dL_dtransMat, float* dL_dsh, float* dL_dscale,
                 float* dL_drot, int debug) {
    RefHandle* r = (RefHandle*)h;
    if (P == 0) return 0;
    try {
        CudaRasterizer::Rasterizer::backward(
            P, D, M, r->num_rendered, bg, W, H, means3D, shs, colors_precomp, scales,
            scale_modifier, rotations, transMat_precomp, viewmatrix, projmatrix, campos,
            tan_fovx, tan_fovy, radii, r->geom.ptr, r->binning.ptr, r->img.ptr, dL_dpix,
            dL_dothers, dL_dmean2D, dL_dnormal, dL_dopacity, dL_dcolor, dL_dmean3D,
            dL_dtransMat, dL_dsh, dL_dscale, dL_drot, debug != 0);
    } catch (...) { return -1; }
    return (int)cudaGetLastError();
}
#endif  // REF_VARIANT_SURFEL

#if defined(REF_VARIANT_GAUSSIAN)
// mirrors RasterizeGaussiansCUDA (G/rasterize_points.cu:35-113); outputs zero-filled by the caller.
int ref_forward(void* h, int P, int D, int M, const float* bg, int W, int H,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp,
                const float* viewmatrix, const float* projmatrix, const float* campos,
                float tan_fovx, float tan_fovy, int prefiltered,
                float* out_color, int* radii, int debug) {
    RefHandle* r = (RefHandle*)h;
    r->num_rendered = 0;
    if (P == 0) return 0;
    std::function<char*(size_t)> g = [r](size_t n) { return r->geom.get(n); };
    std::function<char*(size_t)> b = [r](size_t n) { return r->binning.get(n); };
    std::function<char*(size_t)> i = [r](size_t n) { return r->img.get(n); };
    try {
        r->num_rendered = CudaRasterizer::Rasterizer::forward(
            g, b, i, P, D, M, bg, W, H, means3D, shs, colors_precomp, opacities, scales,
            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
            tan_fovx, tan_fovy, prefiltered != 0, out_color, radii, debug != 0);
    } catch (...) { return -1; }
    return (int)cudaGetLastError();
}

// mirrors RasterizeGaussiansBackwardCUDA (G/rasterize_points.cu:115-197); dL_* zero-filled by the caller.
int ref_backward(void* h, int P, int D, int M, const float* bg, int W, int H,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* scales, float scale_modifier, const float* rotations,
                 const float* cov3D_precomp, const float* viewmatrix,
                 const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
                 const int* radii, const float* dL_dpix,
                 float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                 float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
                 float* dL_drot, int debug) {
    RefHandle* r = (RefHandle*)h;
    if (P == 0) return 0;
    try {
        CudaRasterizer::Rasterizer::backward(
            P, D, M, r->num_rendered, bg, W, H, means3D, shs, colors_precomp, scales,
            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
            tan_fovx, tan_fovy, radii, r->geom.ptr, r->binning.ptr, r->img.ptr, dL_dpix,
            dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh,
            dL_dscale, dL_drot, debug != 0);
    } catch (...) { return -1; }
    return (int)cudaGetLastError();
}
#endif  // REF_VARIANT_GAUSSIAN

#if defined(REF_VARIANT_PLANE)
// mirrors RasterizeGaussiansCUDA (L/rasterize_points.cu:35-125); outputs zero-filled by the caller.
int ref_forward(void* h, int P, int D, int M, const float* bg, int W, int H,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp, const float* all_map,
                const float* viewmatrix, const float* projmatrix, const float* campos,
                float tan_fovx, float tan_fovy, int prefiltered,
                float* out_color, int* radii, int* out_observe, float* out_all_map,
                float* out_plane_depth, int render_geo, int debug) {
    RefHandle* r = (RefHandle*)h;
    r->num_rendered = 0;
    if (P == 0) return 0;
    std::function<char*(size_t)> g = [r](size_t n) { return r->geom.get(n); };
    std::function<char*(size_t)> b = [r](size_t n) { return r->binning.get(n); };
    std::function<char*(size_t)> i = [r](size_t n) { return r->img.get(n); };
    try {
        r->num_rendered = CudaRasterizer::Rasterizer::forward(
            g, b, i, P, D, M, bg, W, H, means3D, shs, colors_precomp, opacities, scales,
            scale_modifier, rotations, cov3D_precomp, all_map, viewmatrix, projmatrix, campos,
            tan_fovx, tan_fovy, prefiltered != 0, out_color, radii, out_observe, out_all_map,
            out_plane_depth, render_geo != 0, debug != 0);
    } catch (...) { return -1; }
    return (int)cudaGetLastError();
}

// mirrors RasterizeGaussiansBackwardCUDA (L/rasterize_points.cu:127-231); dL_* zero-filled by the caller.
int ref_backward(void* h, int P, int D, int M, const float* bg, const float* all_map_pixels, int W, int H,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* all_maps, const float* scales, float scale_modifier,
                 const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                 const float* projmatrix, const float* campos, float tan_fovx, float tan_fovy,
                 const int* radii, const float* dL_dpix, const float* dL_dout_all_map,
                 const float* dL_dout_plane_depth, float* dL_dmean2D, float* dL_dmean2D_abs,
                 float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                 float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                 float* dL_dall_map, int render_geo, int debug) {
    RefHandle* r = (RefHandle*)h;
    if (P == 0) return 0;
    try {
        CudaRasterizer::Rasterizer::backward(
            P, D, M, r->num_rendered, bg, all_map_pixels, W, H, means3D, shs, colors_precomp, all_maps,
            scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos,
            tan_fovx, tan_fovy, radii, r->geom.ptr, r->binning.ptr, r->img.ptr, dL_dpix,
            dL_dout_all_map, dL_dout_plane_depth, dL_dmean2D, dL_dmean2D_abs, dL_dconic, dL_dopacity,
            dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, dL_dall_map,
            render_geo != 0, debug != 0);
    } catch (...) { return -1; }
    return (int)cudaGetLastError();
}
#endif  // REF_VARIANT_PLANE

}  // extern "C"
