// sh.cuh -- real spherical-harmonics colour (degree <= 3) and its gradient.
//
// Same basis, sign convention, +0.5 offset and clamp-at-zero rule as the
// reference (S/cuda_rasterizer/forward.cu:20-71, backward.cu:20-139), written
// as "evaluate the 16 basis polynomials (and their gradients) once, then dot
// with the coefficients" so the forward and the backward share one table.
// Coefficient layout (P, M, 3) like vanilla_gaussian.py:262-265.
#pragma once
#include "common.cuh"

namespace gsr {

__device__ __forceinline__ int sh_count(int deg) { return (deg + 1) * (deg + 1); }

// b[k] for k < (deg+1)^2 at unit direction (x,y,z)
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float* b) {
    b[0] = 0.28209479177387814f;
    if (deg < 1) return;
    const float c1 = 0.4886025119029199f;
    b[1] = -c1 * y; b[2] = c1 * z; b[3] = -c1 * x;
    if (deg < 2) return;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    b[4] = 1.0925484305920792f * xy;
    b[5] = -1.0925484305920792f * yz;
    b[6] = 0.31539156525252005f * (2.0f * zz - xx - yy);
    b[7] = -1.0925484305920792f * xz;
    b[8] = 0.5462742152960396f * (xx - yy);
    if (deg < 3) return;
    b[9] = -0.5900435899266435f * y * (3.0f * xx - yy);
    b[10] = 2.890611442640554f * xy * z;
    b[11] = -0.4570457994644658f * y * (4.0f * zz - xx - yy);
    b[12] = 0.3731763325901154f * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
    b[13] = -0.4570457994644658f * x * (4.0f * zz - xx - yy);
    b[14] = 1.445305721320277f * z * (xx - yy);
    b[15] = -0.5900435899266435f * x * (xx - 3.0f * yy);
}

// gradient of b[k] w.r.t. (x,y,z), k >= 1
__device__ __forceinline__ void sh_basis_grad(int deg, float x, float y, float z, float3* g) {
    g[0] = make_float3(0.f, 0.f, 0.f);
    if (deg < 1) return;
    const float c1 = 0.4886025119029199f;
    g[1] = make_float3(0.f, -c1, 0.f); g[2] = make_float3(0.f, 0.f, c1); g[3] = make_float3(-c1, 0.f, 0.f);
    if (deg < 2) return;
    const float a = 1.0925484305920792f, c26 = 0.31539156525252005f, c28 = 0.5462742152960396f;
    g[4] = make_float3(a * y, a * x, 0.f);
    g[5] = make_float3(0.f, -a * z, -a * y);
    g[6] = make_float3(-2.f * c26 * x, -2.f * c26 * y, 4.f * c26 * z);
    g[7] = make_float3(-a * z, 0.f, -a * x);
    g[8] = make_float3(2.f * c28 * x, -2.f * c28 * y, 0.f);
    if (deg < 3) return;
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    const float k0 = -0.5900435899266435f, k1 = 2.890611442640554f, k2 = -0.4570457994644658f,
                k3 = 0.3731763325901154f, k5 = 1.445305721320277f;
    g[9] = make_float3(k0 * 6.f * xy, k0 * 3.f * (xx - yy), 0.f);
    g[10] = make_float3(k1 * yz, k1 * xz, k1 * xy);
    g[11] = make_float3(k2 * -2.f * xy, k2 * (4.f * zz - xx - 3.f * yy), k2 * 8.f * yz);
    g[12] = make_float3(k3 * -6.f * xz, k3 * -6.f * yz, k3 * 3.f * (2.f * zz - xx - yy));
    g[13] = make_float3(k2 * (4.f * zz - 3.f * xx - yy), k2 * -2.f * xy, k2 * 8.f * xz);
    g[14] = make_float3(k5 * 2.f * xz, k5 * -2.f * yz, k5 * (xx - yy));
    g[15] = make_float3(k0 * 3.f * (xx - yy), k0 * -6.f * xy, 0.f);
}

__device__ __forceinline__ float3 sh_to_rgb(int deg, int M, float3 mean, float3 campos,
                                            const float* __restrict__ sh, uint8_t* clamped) {
    float3 dir = make_float3(mean.x - campos.x, mean.y - campos.y, mean.z - campos.z);
    float inv = 1.0f / sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    float b[16];
    sh_basis(deg, dir.x * inv, dir.y * inv, dir.z * inv, b);
    int n = min(sh_count(deg), M);
    float3 c = make_float3(0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k >= n) break;
        c.x += b[k] * sh[3 * k + 0];      // plain loads: `sh` may point to global OR shared memory (sh.cu)
        c.y += b[k] * sh[3 * k + 1];
        c.z += b[k] * sh[3 * k + 2];
    }
    c.x += 0.5f; c.y += 0.5f; c.z += 0.5f;
    clamped[0] = c.x < 0.f; clamped[1] = c.y < 0.f; clamped[2] = c.z < 0.f;
    return make_float3(fmaxf(c.x, 0.f), fmaxf(c.y, 0.f), fmaxf(c.z, 0.f));
}

// Backward, in place (sh.cu): `c` holds the 3*M coefficients of one Gaussian on entry and
// dL/dsh (zeros above the active degree) on exit; returns the mean gradient through the view direction.
__device__ __forceinline__ float3 sh_to_rgb_bwd_inplace(int deg, int M, float3 mean, float3 campos, float* c,
                                                        const uint8_t* __restrict__ clamped, float3 dL_dcolor) {
    float3 d0 = make_float3(mean.x - campos.x, mean.y - campos.y, mean.z - campos.z);
    float sum2 = d0.x * d0.x + d0.y * d0.y + d0.z * d0.z;
    float inv = 1.0f / sqrtf(sum2);
    float x = d0.x * inv, y = d0.y * inv, z = d0.z * inv;
    float3 dc = make_float3(clamped[0] ? 0.f : dL_dcolor.x, clamped[1] ? 0.f : dL_dcolor.y,
                            clamped[2] ? 0.f : dL_dcolor.z);
    float b[16];
    float3 g[16];
    sh_basis(deg, x, y, z, b);
    sh_basis_grad(deg, x, y, z, g);
    int n = min(sh_count(deg), M);
    float3 ddir = make_float3(0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k >= n) break;
        float s = c[3 * k + 0] * dc.x + c[3 * k + 1] * dc.y + c[3 * k + 2] * dc.z;
        c[3 * k + 0] = b[k] * dc.x;
        c[3 * k + 1] = b[k] * dc.y;
        c[3 * k + 2] = b[k] * dc.z;
        ddir.x += g[k].x * s; ddir.y += g[k].y * s; ddir.z += g[k].z * s;
    }
    for (int j = 3 * n; j < 3 * M; j++) c[j] = 0.f;
    float inv32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
    return make_float3(
        ((sum2 - d0.x * d0.x) * ddir.x - d0.y * d0.x * ddir.y - d0.z * d0.x * ddir.z) * inv32,
        (-d0.x * d0.y * ddir.x + (sum2 - d0.y * d0.y) * ddir.y - d0.z * d0.y * ddir.z) * inv32,
        (-d0.x * d0.z * ddir.x - d0.y * d0.z * ddir.y + (sum2 - d0.z * d0.z) * ddir.z) * inv32);
}

}  // namespace gsr
