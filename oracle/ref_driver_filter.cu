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
    // Scratch as the reference's torch glue provides it: tensors resized through torch's caching allocator, i.e. no
    // cudaMalloc / cudaFree and no synchronisation per call.  Grow-only static buffers stand in for the allocator (a
    // cudaMalloc + cudaFree per call would be this driver's cost, not the reference's: it made the reference arm of
    // bench.py's visible_filter entry 1.2 - 95 ms depending on the box).
    static char* slots[3] = {nullptr, nullptr, nullptr};
    static size_t caps[3] = {0, 0, 0};
    auto mk = [](int k) {
        return std::function<char*(size_t)>([k](size_t n) {
            if (n + 256 > caps[k]) {
                if (slots[k]) cudaFree(slots[k]);
                slots[k] = nullptr;
                caps[k] = 0;
                if (cudaMalloc(&slots[k], n + 256) == cudaSuccess) caps[k] = n + 256;
            }
            return slots[k];
        });
    };
    int rc = 0;
    try {
        CudaRasterizer::Rasterizer::visible_filter(mk(0), mk(1), mk(2), P, 0, W, H, means3D, scales,
                                                   scale_modifier, rotations, cov3D_precomp, viewmatrix,
                                                   projmatrix, tan_fovx, tan_fovy, prefiltered != 0, radii,
                                                   debug != 0);
        rc = (int)cudaGetLastError();          // launch errors; the caller synchronises when it reads the result
    } catch (...) { rc = -1; }
    return rc;
}
