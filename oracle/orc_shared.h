/* oracle/orc_shared.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Helpers shared by the CPU restatements (surfel_oracle.c, gauss_oracle.c): the reference's
 * auxiliary.h transforms / getRect and its SH colour evaluation, which are identical in all
 * rasterizer submodules (S/ = diff-surfel-rasterization, G/ = diff-gaussian-rasterization).
 * Everything is static: each translation unit gets its own copy. */
#ifndef ORC_SHARED_H_
#define ORC_SHARED_H_
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORACLE_DOUBLE
typedef double real;
#define R(x) x
#define rsqrt_r(x) (1.0 / sqrt(x))
#define sqrt_r sqrt
#define exp_r exp
#define ceil_r ceil
#else
typedef float real;
#define R(x) x##f
#define rsqrt_r(x) (1.0f / sqrtf(x))
#define sqrt_r sqrtf
#define exp_r expf
#define ceil_r ceilf
#endif

#define BLOCK_X 16
#define BLOCK_Y 16
#define NEAR_N R(0.2)
#define FAR_N R(100.0)
#define FILTER_SIZE R(0.707106)
#define FILTER_INV_SQUARE R(2.0)

static const real SH_C0 = R(0.28209479177387814);
static const real SH_C1 = R(0.4886025119029199);
static const real SH_C2[5] = {R(1.0925484305920792), R(-1.0925484305920792), R(0.31539156525252005),
                              R(-1.0925484305920792), R(0.5462742152960396)};
static const real SH_C3[7] = {R(-0.5900435899266435), R(2.890611442640554), R(-0.4570457994644658),
                              R(0.3731763325901154), R(-0.4570457994644658), R(1.445305721320277),
                              R(-0.5900435899266435)};

/* S/aux:81-110 */
static void xform43(const real* m, const real* p, real* o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2] + m[12];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2] + m[13];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2] + m[14];
}
static __attribute__((unused)) void xformvec43(const real* m, const real* p, real* o) {
    o[0] = m[0] * p[0] + m[4] * p[1] + m[8] * p[2];
    o[1] = m[1] * p[0] + m[5] * p[1] + m[9] * p[2];
    o[2] = m[2] * p[0] + m[6] * p[1] + m[10] * p[2];
}
/* S/aux:112-120 */
static __attribute__((unused)) void xformvec43T(const real* m, const real* p, real* o) {
    o[0] = m[0] * p[0] + m[1] * p[1] + m[2] * p[2];
    o[1] = m[4] * p[0] + m[5] * p[1] + m[6] * p[2];
    o[2] = m[8] * p[0] + m[9] * p[1] + m[10] * p[2];
}

/* CUDA float->int conversion saturates and maps NaN to 0 (cvt.rzi.s32.f32). */
static int cvt_rzi(real v) {
    if (v != v) return 0;
    if (v >= R(2147483648.0)) return 2147483647;
    if (v <= R(-2147483648.0)) return (-2147483647 - 1);
    return (int)v;
}
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* S/aux:69-79 getRect */
static void get_rect(const real* p, int max_radius, int gx, int gy, int* rmin, int* rmax) {
    rmin[0] = imin(gx, imax(0, cvt_rzi((p[0] - max_radius) / BLOCK_X)));
    rmin[1] = imin(gy, imax(0, cvt_rzi((p[1] - max_radius) / BLOCK_Y)));
    rmax[0] = imin(gx, imax(0, cvt_rzi((p[0] + max_radius + BLOCK_X - 1) / BLOCK_X)));
    rmax[1] = imin(gy, imax(0, cvt_rzi((p[1] + max_radius + BLOCK_Y - 1) / BLOCK_Y)));
}

/* S/fwd:20-71 computeColorFromSH */
static void color_from_sh(int idx, int deg, int M, const real* means, const real* campos,
                          const real* shs, uint8_t* clamped, real* out) {
    real dir[3] = {means[3 * idx] - campos[0], means[3 * idx + 1] - campos[1], means[3 * idx + 2] - campos[2]};
    real len = sqrt_r(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    real x = dir[0] / len, y = dir[1] / len, z = dir[2] / len;
    const real* sh = shs + (size_t)idx * M * 3;
    for (int c = 0; c < 3; c++) {
#define SH(k) sh[(k)*3 + c]
        real r = SH_C0 * SH(0);
        if (deg > 0) {
            r = r - SH_C1 * y * SH(1) + SH_C1 * z * SH(2) - SH_C1 * x * SH(3);
            if (deg > 1) {
                real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * SH(4) + SH_C2[1] * yz * SH(5) +
                    SH_C2[2] * (R(2.0) * zz - xx - yy) * SH(6) + SH_C2[3] * xz * SH(7) +
                    SH_C2[4] * (xx - yy) * SH(8);
                if (deg > 2) {
                    r = r + SH_C3[0] * y * (R(3.0) * xx - yy) * SH(9) + SH_C3[1] * xy * z * SH(10) +
                        SH_C3[2] * y * (R(4.0) * zz - xx - yy) * SH(11) +
                        SH_C3[3] * z * (R(2.0) * zz - R(3.0) * xx - R(3.0) * yy) * SH(12) +
                        SH_C3[4] * x * (R(4.0) * zz - xx - yy) * SH(13) +
                        SH_C3[5] * z * (xx - yy) * SH(14) + SH_C3[6] * x * (xx - R(3.0) * yy) * SH(15);
                }
            }
        }
#undef SH
        r += R(0.5);
        clamped[3 * idx + c] = (r < 0);
        out[c] = r > 0 ? r : 0;
    }
}

static real* to_real(const float* src, size_t n) {
    if (!src || n == 0) return NULL;
    real* d = (real*)malloc(n * sizeof(real));
    for (size_t i = 0; i < n; i++) d[i] = (real)src[i];
    return d;
}

typedef struct { uint64_t key; uint32_t idx; } KeyIdx;
static int cmp_keyidx(const void* a, const void* b) {
    const KeyIdx* x = (const KeyIdx*)a; const KeyIdx* y = (const KeyIdx*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    /* stable radix sort of pairs emitted in ascending Gaussian index (S/impl:70-111,304-309) */
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* S/bwd:20-139 SH backward.  dL_dmeans is accumulated (+=). */
static void color_from_sh_bwd(int deg, const real* mean, const real* campos, const real* sh,
                              const uint8_t* clamped, const real* dL_dcolor, real* dL_dmean,
                              real* dsh) {
    real dir_orig[3] = {mean[0] - campos[0], mean[1] - campos[1], mean[2] - campos[2]};
    real len = sqrt_r(dir_orig[0] * dir_orig[0] + dir_orig[1] * dir_orig[1] + dir_orig[2] * dir_orig[2]);
    real x = dir_orig[0] / len, y = dir_orig[1] / len, z = dir_orig[2] / len;
    real dRGB[3];
    for (int c = 0; c < 3; c++) dRGB[c] = dL_dcolor[c] * (clamped[c] ? 0 : 1);
    real ddir[3] = {0, 0, 0};
    for (int c = 0; c < 3; c++) {
#define SH(k) sh[(k)*3 + c]
#define DSH(k) dsh[(k)*3 + c]
        real dx = 0, dy = 0, dz = 0;
        DSH(0) = SH_C0 * dRGB[c];
        if (deg > 0) {
            DSH(1) = (-SH_C1 * y) * dRGB[c]; DSH(2) = (SH_C1 * z) * dRGB[c]; DSH(3) = (-SH_C1 * x) * dRGB[c];
            dx = -SH_C1 * SH(3); dy = -SH_C1 * SH(1); dz = SH_C1 * SH(2);
            if (deg > 1) {
                real xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                DSH(4) = (SH_C2[0] * xy) * dRGB[c]; DSH(5) = (SH_C2[1] * yz) * dRGB[c];
                DSH(6) = (SH_C2[2] * (R(2.) * zz - xx - yy)) * dRGB[c];
                DSH(7) = (SH_C2[3] * xz) * dRGB[c]; DSH(8) = (SH_C2[4] * (xx - yy)) * dRGB[c];
                dx += SH_C2[0] * y * SH(4) + SH_C2[2] * R(2.) * -x * SH(6) + SH_C2[3] * z * SH(7) + SH_C2[4] * R(2.) * x * SH(8);
                dy += SH_C2[0] * x * SH(4) + SH_C2[1] * z * SH(5) + SH_C2[2] * R(2.) * -y * SH(6) + SH_C2[4] * R(2.) * -y * SH(8);
                dz += SH_C2[1] * y * SH(5) + SH_C2[2] * R(2.) * R(2.) * z * SH(6) + SH_C2[3] * x * SH(7);
                if (deg > 2) {
                    DSH(9) = (SH_C3[0] * y * (R(3.) * xx - yy)) * dRGB[c];
                    DSH(10) = (SH_C3[1] * xy * z) * dRGB[c];
                    DSH(11) = (SH_C3[2] * y * (R(4.) * zz - xx - yy)) * dRGB[c];
                    DSH(12) = (SH_C3[3] * z * (R(2.) * zz - R(3.) * xx - R(3.) * yy)) * dRGB[c];
                    DSH(13) = (SH_C3[4] * x * (R(4.) * zz - xx - yy)) * dRGB[c];
                    DSH(14) = (SH_C3[5] * z * (xx - yy)) * dRGB[c];
                    DSH(15) = (SH_C3[6] * x * (xx - R(3.) * yy)) * dRGB[c];
                    dx += SH_C3[0] * SH(9) * R(3.) * R(2.) * xy + SH_C3[1] * SH(10) * yz + SH_C3[2] * SH(11) * R(-2.) * xy +
                          SH_C3[3] * SH(12) * R(-3.) * R(2.) * xz + SH_C3[4] * SH(13) * (R(-3.) * xx + R(4.) * zz - yy) +
                          SH_C3[5] * SH(14) * R(2.) * xz + SH_C3[6] * SH(15) * R(3.) * (xx - yy);
                    dy += SH_C3[0] * SH(9) * R(3.) * (xx - yy) + SH_C3[1] * SH(10) * xz +
                          SH_C3[2] * SH(11) * (R(-3.) * yy + R(4.) * zz - xx) + SH_C3[3] * SH(12) * R(-3.) * R(2.) * yz +
                          SH_C3[4] * SH(13) * R(-2.) * xy + SH_C3[5] * SH(14) * R(-2.) * yz + SH_C3[6] * SH(15) * R(-3.) * R(2.) * xy;
                    dz += SH_C3[1] * SH(10) * xy + SH_C3[2] * SH(11) * R(4.) * R(2.) * yz +
                          SH_C3[3] * SH(12) * R(3.) * (R(2.) * zz - xx - yy) + SH_C3[4] * SH(13) * R(4.) * R(2.) * xz +
                          SH_C3[5] * SH(14) * (xx - yy);
                }
            }
        }
#undef SH
#undef DSH
        ddir[0] += dx * dRGB[c]; ddir[1] += dy * dRGB[c]; ddir[2] += dz * dRGB[c];
    }
    /* dnormvdv, S/aux:130-140 */
    real* v = dir_orig;
    real sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    real invsum32 = R(1.0) / sqrt_r(sum2 * sum2 * sum2);
    dL_dmean[0] += ((+sum2 - v[0] * v[0]) * ddir[0] - v[1] * v[0] * ddir[1] - v[2] * v[0] * ddir[2]) * invsum32;
    dL_dmean[1] += (-v[0] * v[1] * ddir[0] + (sum2 - v[1] * v[1]) * ddir[1] - v[2] * v[1] * ddir[2]) * invsum32;
    dL_dmean[2] += (-v[0] * v[2] * ddir[0] - v[1] * v[2] * ddir[1] + (sum2 - v[2] * v[2]) * ddir[2]) * invsum32;
}

static void atomic_add_d(double* a, double v) {
#pragma omp atomic
    *a += v;
}

#endif /* ORC_SHARED_H_ */
