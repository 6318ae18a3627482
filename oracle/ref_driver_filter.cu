// oracle/ref_driver_filter.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C driver around the UNMODIFIED reference scaffold-filter sources
// (submodules/scaffold-filter/cuda_rasterizer/{forward,rasterizer_impl}.cu, compiled where
// they lie under /root/reference by oracle/build_ref.sh into oracle/_ref/libref_filter.so).
// Replaces only the torch glue RasterizeGaussiansfilterCUDA (F/rasterize_points.cu:220-284).
// Users: tests/ and tests/golden/make_golden.py.  Never loaded by the product path.
#include <cstdint>
#include <functional>
#include <cuda_runtime.h>
#include "cuda_rasterizer/config.h"
#include "cuda_rasterizer/rasterizer.h"

extern "C" int ref_visible_filter(int P, int W, int H, const float* means3D, const float* scales,
                                  float scale_modifier, const float* rotations, const float* cov3D_precomp,
                                  const float* viewmatrix, const float* projmatrix, float tan_fovx,
                                  float tan_fovy, int prefiltered, int* radii, int debug) {
    if (P == 0) return 0;
    char *g = nullptr, *b = nullptr, *i = nullptr;
    auto mk = [](char** slot) {
        return std::function<char*(size_t)>([slot](size_t n) {
            if (*slot) cudaFree(*slot);
            *slot = nullptr;
            cudaMalloc(slot, n + 256);
            return *slot;
        });
    };
    int rc = 0;
    try {
        CudaRasterizer::Rasterizer::visible_filter(mk(&g), mk(&b), mk(&i), P, 0, W, H, means3D, scales,
                                                   scale_modifier, rotations, cov3D_precomp, viewmatrix,
                                                   projmatrix, tan_fovx, tan_fovy, prefiltered != 0, radii,
                                                   debug != 0);
        cudaDeviceSynchronize();
        rc = (int)cudaGetLastError();
    } catch (...) { rc = -1; }
    if (g) cudaFree(g);
    if (b) cudaFree(b);
    if (i) cudaFree(i);
    return rc;
}
