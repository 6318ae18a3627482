/* oracle/surfel_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, OpenMP over tiles) of the reference 2DGS surfel
 * rasterizer, forward and backward:
 *   submodules/diff-surfel-rasterization/cuda_rasterizer/forward.cu
 *   submodules/diff-surfel-rasterization/cuda_rasterizer/backward.cu
 *   submodules/diff-surfel-rasterization/cuda_rasterizer/rasterizer_impl.cu
 *   submodules/diff-surfel-rasterization/cuda_rasterizer/auxiliary.h
 * (abbreviated S/fwd, S/bwd, S/impl, S/aux below; each function cites the
 * lines it follows).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product path (gs-sr_b200/) never
 * does.
 *
 * Pinning status: the reference ships no golden vectors.
 * This restatement is pinned against outputs of the reference CUDA kernels
 * themselves (oracle/_ref/libref_surfel.so, built by oracle/build_ref.sh from
 * the unmodified sources and run on a B200); the captured vectors live in
 * tests/golden/ with the generating script tests/golden/make_golden.py.
 *
 * Arithmetic: per-pixel and per-Gaussian math is float32 like the reference
 * (`real` = float); per-Gaussian gradient sums, which the reference forms
 * with float atomicAdd in arbitrary order, are accumulated here in double so
 * the oracle is the order-independent value both GPU implementations
 * approximate.  Build with -DORACLE_DOUBLE for an all-double variant.
 */
#include "orc_shared.h"

typedef struct {
    int P, D, M, W, H, gx, gy;
    int has_sh, has_scales; /* what the forward call was given */
    real scale_modifier, tan_fovx, tan_fovy, focal_x, focal_y;
    real view[16], proj[16], campos[3], bg[3];
    /* GeometryState (S/impl:155-170) */
    real* depths;
    uint8_t* clamped;
    int* radii;
    real* xy;             /* P*2 */
    real* transMat;       /* P*9 */
    real* normal_opacity; /* P*4 */
    real* rgb;            /* P*3 */
    uint32_t* tiles_touched;
    /* BinningState */
    int64_t R;
    uint32_t* point_list;
    uint64_t* point_keys;
    /* ImageState (S/impl:172-179) */
    uint32_t* ranges;    /* tiles*2 */
    real* final_T;       /* 3N: T, M1, M2 */
    uint32_t* n_contrib; /* 2N: last contributor, median contributor */
    int64_t k_eval;      /* (pixel,splat) pairs walked in forward (bench statistic) */
    int64_t k_blend;     /* pairs that passed every test and were blended */
} OrcSurfel;

/* S/aux:215-237: rotation matrix from (w,x,y,z), normalised in-kernel.
 * Rm[c][r] = column c, row r (glm column-major). */
static void quat_to_rotmat(const real* q, real Rm[3][3]) {
    real s = rsqrt_r(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    real w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
    Rm[0][0] = R(1.) - R(2.) * (y * y + z * z);
    Rm[0][1] = R(2.) * (x * y + w * z);
    Rm[0][2] = R(2.) * (x * z - w * y);
    Rm[1][0] = R(2.) * (x * y - w * z);
    Rm[1][1] = R(1.) - R(2.) * (x * x + z * z);
    Rm[1][2] = R(2.) * (y * z + w * x);
    Rm[2][0] = R(2.) * (x * z + w * y);
    Rm[2][1] = R(2.) * (y * z - w * x);
    Rm[2][2] = R(1.) - R(2.) * (x * x + y * y);
}

/* S/aux:240-284 (v_R[c][r] column-major, result in (w,x,y,z) order) */
static void quat_to_rotmat_vjp(const real* q, real vR[3][3], real* vq) {
    real s = rsqrt_r(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    real w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
    vq[0] = R(2.) * (x * (vR[1][2] - vR[2][1]) + y * (vR[2][0] - vR[0][2]) + z * (vR[0][1] - vR[1][0]));
    vq[1] = R(2.) * (R(-2.) * x * (vR[1][1] + vR[2][2]) + y * (vR[0][1] + vR[1][0]) +
                     z * (vR[0][2] + vR[2][0]) + w * (vR[1][2] - vR[2][1]));
    vq[2] = R(2.) * (x * (vR[0][1] + vR[1][0]) - R(2.) * y * (vR[0][0] + vR[2][2]) +
                     z * (vR[1][2] + vR[2][1]) + w * (vR[2][0] - vR[0][2]));
    vq[3] = R(2.) * (x * (vR[0][2] + vR[2][0]) + y * (vR[1][2] + vR[2][1]) -
                     R(2.) * z * (vR[0][0] + vR[1][1]) + w * (vR[0][1] - vR[1][0]));
}

/* P = world2ndc * ndc2pix as used at S/fwd:99-112 and S/bwd:497-510.
 * Pm[i][r]: column i (0: x*w pixel, 1: y*w pixel, 2: w), row r (world homogeneous index). */
static void build_P(const real* proj, int W, int H, real Pm[3][4]) {
    for (int r = 0; r < 4; r++) {
        /* world2ndc (math) X(r,c) = proj[4r+c]; ndc2pix columns per S/fwd:106-110 */
        real x = proj[4 * r + 0], y = proj[4 * r + 1], w = proj[4 * r + 3];
        Pm[0][r] = x * ((real)W / R(2.0)) + w * ((real)(W - 1) / R(2.0));
        Pm[1][r] = y * ((real)H / R(2.0)) + w * ((real)(H - 1) / R(2.0));
        Pm[2][r] = w;
    }
}

/* S/fwd:75-115 compute_transmat (bwd_order = 0) and its re-evaluation at
 * S/bwd:484-512 (bwd_order = 1).  T[i][j]: i = output coordinate (x*w, y*w, w),
 * j = (u axis, v axis, centre); stored transMat[3i+j] as at S/fwd:195-198.
 * The forward multiplies (splat2world^T * world2ndc) * ndc2pix, the backward
 * splat2world^T * (world2ndc * ndc2pix); both association orders are kept. */
static void compute_transmat(const real* p, const real* scale, real mod, const real* rot,
                             const real* proj, const real* view, int W, int H, int bwd_order,
                             real T[3][3], real* normal) {
    real Rm[3][3];
    quat_to_rotmat(rot, Rm);
    real L0[3], L1[3], L2[3];
    for (int r = 0; r < 3; r++) {
        L0[r] = Rm[0][r] * (mod * scale[0]);
        L1[r] = Rm[1][r] * (mod * scale[1]);
        L2[r] = Rm[2][r];
    }
    if (bwd_order) {
        real Pm[3][4];
        build_P(proj, W, H, Pm);
        for (int i = 0; i < 3; i++) {
            T[i][0] = L0[0] * Pm[i][0] + L0[1] * Pm[i][1] + L0[2] * Pm[i][2];
            T[i][1] = L1[0] * Pm[i][0] + L1[1] * Pm[i][1] + L1[2] * Pm[i][2];
            T[i][2] = p[0] * Pm[i][0] + p[1] * Pm[i][1] + p[2] * Pm[i][2] + Pm[i][3];
        }
    } else {
        const real* vec[3] = {L0, L1, p};
        const real hw[3] = {R(0.0), R(0.0), R(1.0)};
        for (int j = 0; j < 3; j++) {
            real clip[4];
            for (int c = 0; c < 4; c++)
                clip[c] = vec[j][0] * proj[c] + vec[j][1] * proj[4 + c] + vec[j][2] * proj[8 + c] + hw[j] * proj[12 + c];
            T[0][j] = clip[0] * ((real)W / R(2.0)) + clip[3] * ((real)(W - 1) / R(2.0));
            T[1][j] = clip[1] * ((real)H / R(2.0)) + clip[3] * ((real)(H - 1) / R(2.0));
            T[2][j] = clip[3];
        }
    }
    xformvec43(view, L2, normal);
}

/* S/fwd:119-145 compute_aabb */
static int compute_aabb(real T[3][3], real cutoff, real* center, real* extent) {
    real t[3] = {cutoff * cutoff, cutoff * cutoff, R(-1.0)};
    real d = t[0] * T[2][0] * T[2][0] + t[1] * T[2][1] * T[2][1] + t[2] * T[2][2] * T[2][2];
    if (d == R(0.0)) return 0;
    real f[3] = {(R(1.) / d) * t[0], (R(1.) / d) * t[1], (R(1.) / d) * t[2]};
    real p0 = f[0] * T[0][0] * T[2][0] + f[1] * T[0][1] * T[2][1] + f[2] * T[0][2] * T[2][2];
    real p1 = f[0] * T[1][0] * T[2][0] + f[1] * T[1][1] * T[2][1] + f[2] * T[1][2] * T[2][2];
    real h0 = p0 * p0 - (f[0] * T[0][0] * T[0][0] + f[1] * T[0][1] * T[0][1] + f[2] * T[0][2] * T[0][2]);
    real h1 = p1 * p1 - (f[0] * T[1][0] * T[1][0] + f[1] * T[1][1] * T[1][1] + f[2] * T[1][2] * T[1][2]);
    real m0 = h0 > R(1e-4) ? h0 : R(1e-4), m1 = h1 > R(1e-4) ? h1 : R(1e-4); /* max(1e-4, h0); NaN -> 1e-4 as CUDA fmaxf */
    if (h0 != h0) m0 = R(1e-4);
    if (h1 != h1) m1 = R(1e-4);
    center[0] = p0;
    center[1] = p1;
    extent[0] = sqrt_r(m0);
    extent[1] = sqrt_r(m1);
    return 1;
}

void orc_surfel_free(OrcSurfel* o) {
    if (!o) return;
    free(o->depths); free(o->clamped); free(o->radii); free(o->xy); free(o->transMat);
    free(o->normal_opacity); free(o->rgb); free(o->tiles_touched); free(o->point_list);
    free(o->point_keys); free(o->ranges); free(o->final_T); free(o->n_contrib);
    free(o);
}

/* Forward: S/impl:198-342 (host orchestration), S/fwd:148-251 (preprocess),
 * S/impl:70-138 (binning), S/fwd:256-448 (render).
 * tile_stride > 1 renders only tiles with (tx % stride == 0 && ty % stride == 0)
 * (bounded CPU-baseline sample); other pixels are left untouched.
 * Returns a handle (NULL on error, *err: 1 = prefiltered trap). */
OrcSurfel* orc_surfel_forward(int P, int D, int M, const float* bg, int W, int H,
                              const float* means3D_f, const float* shs_f, const float* colors_f,
                              const float* opac_f, const float* scales_f, float scale_modifier,
                              const float* rot_f, const float* transMat_precomp_f,
                              const float* view_f, const float* proj_f, const float* campos_f,
                              float tan_fovx, float tan_fovy, int prefiltered, int tile_stride,
                              float* out_color, float* out_others, int* out_radii, int* err) {
    *err = 0;
    OrcSurfel* o = (OrcSurfel*)calloc(1, sizeof(OrcSurfel));
    o->P = P; o->D = D; o->M = M; o->W = W; o->H = H;
    o->gx = (W + BLOCK_X - 1) / BLOCK_X; o->gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    o->has_sh = shs_f != NULL; o->has_scales = scales_f != NULL;
    o->scale_modifier = scale_modifier; o->tan_fovx = tan_fovx; o->tan_fovy = tan_fovy;
    /* S/impl:223-224 (float32 in the reference) */
    o->focal_y = (real)((float)H / (2.0f * tan_fovy));
    o->focal_x = (real)((float)W / (2.0f * tan_fovx));
    for (int i = 0; i < 16; i++) { o->view[i] = view_f[i]; o->proj[i] = proj_f[i]; }
    for (int i = 0; i < 3; i++) { o->campos[i] = campos_f[i]; o->bg[i] = bg[i]; }
    const int N = W * H, ntiles = o->gx * o->gy;
    size_t Pn = P > 0 ? (size_t)P : 1;
    o->depths = (real*)calloc(Pn, sizeof(real));
    o->clamped = (uint8_t*)calloc(Pn * 3, 1);
    o->radii = (int*)calloc(Pn, sizeof(int));
    o->xy = (real*)calloc(Pn * 2, sizeof(real));
    o->transMat = (real*)calloc(Pn * 9, sizeof(real));
    o->normal_opacity = (real*)calloc(Pn * 4, sizeof(real));
    o->rgb = (real*)calloc(Pn * 3, sizeof(real));
    o->tiles_touched = (uint32_t*)calloc(Pn, sizeof(uint32_t));
    o->ranges = (uint32_t*)calloc((size_t)ntiles * 2, sizeof(uint32_t));
    o->final_T = (real*)calloc((size_t)N * 3, sizeof(real));
    o->n_contrib = (uint32_t*)calloc((size_t)N * 2, sizeof(uint32_t));

    real* means = to_real(means3D_f, (size_t)P * 3);
    real* shs = to_real(shs_f, (size_t)P * M * 3);
    real* colors = to_real(colors_f, (size_t)P * 3);
    real* opac = to_real(opac_f, (size_t)P);
    real* scales = to_real(scales_f, (size_t)P * 2);
    real* rots = to_real(rot_f, (size_t)P * 4);
    real* Tpre = to_real(transMat_precomp_f, (size_t)P * 9);
    int trap = 0;

    /* ---- preprocess, S/fwd:148-251 ---- */
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        real pv[3];
        xform43(o->view, means + 3 * idx, pv); /* in_frustum, S/aux:187-212 */
        if (pv[2] <= R(0.2)) { if (prefiltered) trap = 1; continue; }
        real T[3][3], normal[3];
        if (!Tpre) {
            compute_transmat(means + 3 * idx, scales + 2 * idx, o->scale_modifier, rots + 4 * idx,
                             o->proj, o->view, W, H, 0, T, normal);
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o->transMat[9 * idx + 3 * i + j] = T[i][j];
        } else {
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[i][j] = Tpre[9 * idx + 3 * i + j];
            normal[0] = 0; normal[1] = 0; normal[2] = 1;
        }
        real cosv = -(pv[0] * normal[0] + pv[1] * normal[1] + pv[2] * normal[2]); /* DUAL_VISIABLE S/fwd:209-214 */
        if (cosv == 0) continue;
        real mult = cosv > 0 ? R(1.) : R(-1.);
        normal[0] *= mult; normal[1] *= mult; normal[2] *= mult;
        real center[2], extent[2];
        if (!compute_aabb(T, R(3.0), center, extent)) continue;
        real mx = extent[0] > extent[1] ? extent[0] : extent[1];
        real cf = R(3.0) * FILTER_SIZE;
        real radius = ceil_r(mx > cf ? mx : cf);
        int rmin[2], rmax[2];
        get_rect(center, cvt_rzi(radius), o->gx, o->gy, rmin, rmax);
        if ((rmax[0] - rmin[0]) * (rmax[1] - rmin[1]) == 0) continue;
        if (!colors) color_from_sh(idx, D, M, means, o->campos, shs, o->clamped, o->rgb + 3 * idx);
        o->depths[idx] = pv[2];
        o->radii[idx] = cvt_rzi(radius);
        o->xy[2 * idx] = center[0]; o->xy[2 * idx + 1] = center[1];
        o->normal_opacity[4 * idx + 0] = normal[0]; o->normal_opacity[4 * idx + 1] = normal[1];
        o->normal_opacity[4 * idx + 2] = normal[2]; o->normal_opacity[4 * idx + 3] = opac[idx];
        o->tiles_touched[idx] = (uint32_t)((rmax[1] - rmin[1]) * (rmax[0] - rmin[0]));
    }
    if (trap) { *err = 1; }
    if (out_radii) for (int i = 0; i < P; i++) out_radii[i] = o->radii[i];

    /* ---- binning, S/impl:276-320 ---- */
    int64_t Rn = 0;
    for (int i = 0; i < P; i++) Rn += o->tiles_touched[i];
    o->R = Rn;
    KeyIdx* ki = (KeyIdx*)malloc((size_t)(Rn > 0 ? Rn : 1) * sizeof(KeyIdx));
    {
        int64_t off = 0;
        for (int idx = 0; idx < P; idx++) {
            if (o->radii[idx] <= 0) continue;
            int rmin[2], rmax[2];
            get_rect(o->xy + 2 * idx, o->radii[idx], o->gx, o->gy, rmin, rmax);
            float df = (float)o->depths[idx];
            uint32_t dbits; memcpy(&dbits, &df, 4);
            for (int y = rmin[1]; y < rmax[1]; y++)
                for (int x = rmin[0]; x < rmax[0]; x++) {
                    ki[off].key = ((uint64_t)(uint32_t)(y * o->gx + x) << 32) | dbits;
                    ki[off].idx = (uint32_t)idx;
                    off++;
                }
        }
    }
    qsort(ki, (size_t)Rn, sizeof(KeyIdx), cmp_keyidx);
    o->point_list = (uint32_t*)malloc((size_t)(Rn > 0 ? Rn : 1) * 4);
    o->point_keys = (uint64_t*)malloc((size_t)(Rn > 0 ? Rn : 1) * 8);
    for (int64_t i = 0; i < Rn; i++) { o->point_list[i] = ki[i].idx; o->point_keys[i] = ki[i].key; }
    free(ki);
    for (int64_t i = 0; i < Rn; i++) { /* identifyTileRanges S/impl:116-138 */
        uint32_t cur = (uint32_t)(o->point_keys[i] >> 32);
        if (i == 0) o->ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(o->point_keys[i - 1] >> 32);
            if (cur != prev) { o->ranges[2 * prev + 1] = (uint32_t)i; o->ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == Rn - 1) o->ranges[2 * cur + 1] = (uint32_t)Rn;
    }

    /* ---- render, S/fwd:256-448 ---- */
    const real* feat = colors ? colors : o->rgb;
    const real* TM = Tpre ? Tpre : o->transMat;
    int64_t k_eval = 0, k_blend = 0;
    if (tile_stride < 1) tile_stride = 1;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : k_eval, k_blend)
    for (int tile = 0; tile < ntiles; tile++) {
        int tx = tile % o->gx, ty = tile / o->gx;
        if (tx % tile_stride || ty % tile_stride) continue;
        uint32_t r0 = o->ranges[2 * tile], r1 = o->ranges[2 * tile + 1];
        for (int ly = 0; ly < BLOCK_Y; ly++)
            for (int lx = 0; lx < BLOCK_X; lx++) {
                int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
                if (px >= W || py >= H) continue;
                int pix_id = W * py + px;
                real pixf[2] = {(real)px, (real)py};
                real T = 1, C[3] = {0, 0, 0}, Nn[3] = {0, 0, 0}, Dd = 0, M1 = 0, M2 = 0, distortion = 0;
                real median_depth = 0, median_contributor = -1, median_normal[3] = {0, 0, 0};
                int surf_idx = -1;
                uint32_t contributor = 0, last_contributor = 0;
                for (uint32_t e = r0; e < r1; e++) {
                    contributor++;
                    uint32_t g = o->point_list[e];
                    const real* Tu = TM + 9 * g; const real* Tv = Tu + 3; const real* Tw = Tu + 6;
                    real k[3] = {pixf[0] * Tw[0] - Tu[0], pixf[0] * Tw[1] - Tu[1], pixf[0] * Tw[2] - Tu[2]};
                    real l[3] = {pixf[1] * Tw[0] - Tv[0], pixf[1] * Tw[1] - Tv[1], pixf[1] * Tw[2] - Tv[2]};
                    real p[3] = {k[1] * l[2] - k[2] * l[1], k[2] * l[0] - k[0] * l[2], k[0] * l[1] - k[1] * l[0]};
                    if (p[2] == 0) continue;
                    real s[2] = {p[0] / p[2], p[1] / p[2]};
                    real rho3d = s[0] * s[0] + s[1] * s[1];
                    real d[2] = {o->xy[2 * g] - pixf[0], o->xy[2 * g + 1] - pixf[1]};
                    real rho2d = FILTER_INV_SQUARE * (d[0] * d[0] + d[1] * d[1]);
                    real rho = rho3d < rho2d ? rho3d : rho2d;
                    real depth = (rho3d <= rho2d) ? (s[0] * Tw[0] + s[1] * Tw[1]) + Tw[2] : Tw[2];
                    if (depth < NEAR_N) continue;
                    const real* no = o->normal_opacity + 4 * g;
                    real power = R(-0.5) * rho;
                    if (power > 0) continue;
                    real alpha = no[3] * exp_r(power);
                    if (alpha > R(0.99)) alpha = R(0.99);
                    if (alpha < R(1.0) / R(255.0)) continue;
                    real test_T = T * (1 - alpha);
                    if (test_T < R(0.0001)) break; /* done = true */
                    real w = alpha * T;
                    real A = 1 - T;
                    real m = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N / depth);
                    distortion += (m * m * A + M2 - 2 * m * M1) * w;
                    Dd += depth * w; M1 += m * w; M2 += m * m * w;
                    if (T > R(0.5)) {
                        median_depth = depth; surf_idx = (int)g;
                        median_normal[0] = no[0]; median_normal[1] = no[1]; median_normal[2] = no[2];
                        median_contributor = (real)contributor;
                    }
                    for (int c = 0; c < 3; c++) Nn[c] += no[c] * w;
                    for (int c = 0; c < 3; c++) C[c] += feat[3 * g + c] * w;
                    T = test_T;
                    last_contributor = contributor;
                    k_blend++;
                }
                k_eval += contributor;
                o->final_T[pix_id] = T;
                o->n_contrib[pix_id] = last_contributor;
                /* float -1 -> uint32 saturates to 0 on the GPU */
                o->n_contrib[pix_id + N] = median_contributor < 0 ? 0u : (uint32_t)median_contributor;
                o->final_T[pix_id + N] = M1;
                o->final_T[pix_id + 2 * N] = M2;
                if (out_color) for (int c = 0; c < 3; c++) out_color[c * N + pix_id] = (float)(C[c] + T * o->bg[c]);
                if (out_others) {
                    out_others[pix_id + 0 * N] = (float)Dd;
                    out_others[pix_id + 1 * N] = (float)(1 - T);
                    for (int c = 0; c < 3; c++) out_others[pix_id + (2 + c) * N] = (float)Nn[c];
                    out_others[pix_id + 5 * N] = (float)median_depth;
                    out_others[pix_id + 6 * N] = (float)distortion;
                    out_others[pix_id + 7 * N] = (float)surf_idx;
                    for (int c = 0; c < 3; c++) out_others[pix_id + (8 + c) * N] = (float)median_normal[c];
                }
            }
    }
    o->k_eval = k_eval;
    o->k_blend = k_blend;
    free(means); free(shs); free(colors); free(opac); free(scales); free(rots); free(Tpre);
    return o;
}

/* accessors for tests */
int64_t orc_surfel_num_rendered(OrcSurfel* o) { return o->R; }
int64_t orc_surfel_k_eval(OrcSurfel* o) { return o->k_eval; }
int64_t orc_surfel_k_blend(OrcSurfel* o) { return o->k_blend; }
void orc_surfel_get_geom(OrcSurfel* o, float* depths, float* xy, float* transMat, float* normal_opacity,
                         float* rgb, int* tiles_touched) {
    for (int i = 0; i < o->P; i++) {
        if (depths) depths[i] = (float)o->depths[i];
        if (xy) { xy[2 * i] = (float)o->xy[2 * i]; xy[2 * i + 1] = (float)o->xy[2 * i + 1]; }
        if (transMat) for (int j = 0; j < 9; j++) transMat[9 * i + j] = (float)o->transMat[9 * i + j];
        if (normal_opacity) for (int j = 0; j < 4; j++) normal_opacity[4 * i + j] = (float)o->normal_opacity[4 * i + j];
        if (rgb) for (int j = 0; j < 3; j++) rgb[3 * i + j] = (float)o->rgb[3 * i + j];
        if (tiles_touched) tiles_touched[i] = (int)o->tiles_touched[i];
    }
}
void orc_surfel_get_binning(OrcSurfel* o, uint32_t* point_list, uint32_t* ranges) {
    if (point_list) memcpy(point_list, o->point_list, (size_t)o->R * 4);
    if (ranges) memcpy(ranges, o->ranges, (size_t)o->gx * o->gy * 8);
}
void orc_surfel_get_image_state(OrcSurfel* o, float* final_T, uint32_t* n_contrib) {
    size_t N = (size_t)o->W * o->H;
    if (final_T) for (size_t i = 0; i < 3 * N; i++) final_T[i] = (float)o->final_T[i];
    if (n_contrib) memcpy(n_contrib, o->n_contrib, 2 * N * 4);
}

/* Backward: S/impl:346-448, S/bwd:143-447 (render), S/bwd:450-637 (preprocess).
 * All outputs are float arrays the caller owns (fully written here, zeros for
 * radii == 0 as torch::zeros does at S/rasterize_points.cu:188-196).
 * dL_dnormal (P,3) and dL_dmean2D_raw (P,2: the screen-space accumulators before
 * the densification hack overwrites them) are extra taps for tests; may be NULL. */
void orc_surfel_backward(OrcSurfel* o, const float* means3D_f, const float* shs_f,
                         const float* colors_f, const float* scales_f, const float* rot_f,
                         const float* transMat_precomp_f, const float* dL_dpix_f,
                         const float* dL_dothers_f, int tile_stride, float* dL_dmean2D,
                         float* dL_dcolors, float* dL_dopacity, float* dL_dmean3D,
                         float* dL_dtransMat, float* dL_dsh, float* dL_dscales, float* dL_drots,
                         float* dL_dnormal_out, float* dL_dmean2D_raw) {
    const int P = o->P, W = o->W, H = o->H, N = W * H, M = o->M, ntiles = o->gx * o->gy;
    real* means = to_real(means3D_f, (size_t)P * 3);
    real* shs = to_real(shs_f, (size_t)P * M * 3);
    real* colors = to_real(colors_f, (size_t)P * 3);
    real* scales = to_real(scales_f, (size_t)P * 2);
    real* rots = to_real(rot_f, (size_t)P * 4);
    real* Tpre = to_real(transMat_precomp_f, (size_t)P * 9);
    real* dpix = to_real(dL_dpix_f, (size_t)3 * N);
    real* doth = to_real(dL_dothers_f, (size_t)11 * N);
    const real* feat = colors ? colors : o->rgb;
    const real* TM = Tpre ? Tpre : o->transMat;
    size_t Pn = P > 0 ? (size_t)P : 1;
    /* order-independent accumulators for what the reference builds with atomicAdd */
    double* aT = (double*)calloc(Pn * 9, 8);
    double* aM2 = (double*)calloc(Pn * 2, 8);
    double* aN = (double*)calloc(Pn * 3, 8);
    double* aO = (double*)calloc(Pn, 8);
    double* aC = (double*)calloc(Pn * 3, 8);
    if (tile_stride < 1) tile_stride = 1;

    /* ---- render backward, S/bwd:143-447 ---- */
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < ntiles; tile++) {
        int tx = tile % o->gx, ty = tile / o->gx;
        if (tx % tile_stride || ty % tile_stride) continue;
        uint32_t r0 = o->ranges[2 * tile], r1 = o->ranges[2 * tile + 1];
        int toDo = (int)(r1 - r0);
        for (int ly = 0; ly < BLOCK_Y; ly++)
            for (int lx = 0; lx < BLOCK_X; lx++) {
                int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
                if (px >= W || py >= H) continue;
                int pix_id = W * py + px;
                real pixf[2] = {(real)px, (real)py};
                const real T_final = o->final_T[pix_id];
                real T = T_final;
                uint32_t contributor = (uint32_t)toDo;
                const int last_contributor = (int)o->n_contrib[pix_id];
                real accum_rec[3] = {0, 0, 0}, dL_dpixel[3];
                const int median_contributor = (int)o->n_contrib[pix_id + N];
                real dL_ddepth = doth[0 * N + pix_id], dL_daccum = doth[1 * N + pix_id], dL_dreg = doth[6 * N + pix_id];
                real dL_dnormal2D[3], dL_dmedian_normal2D[3];
                for (int i = 0; i < 3; i++) dL_dnormal2D[i] = doth[(2 + i) * N + pix_id];
                real dL_dmedian_depth = doth[5 * N + pix_id];
                for (int i = 0; i < 3; i++) dL_dmedian_normal2D[i] = doth[(8 + i) * N + pix_id];
                real last_depth = 0, last_normal[3] = {0, 0, 0}, accum_depth_rec = 0, accum_alpha_rec = 0;
                real accum_normal_rec[3] = {0, 0, 0};
                const real final_D = o->final_T[pix_id + N], final_D2 = o->final_T[pix_id + 2 * N];
                const real final_A = 1 - T_final;
                real last_dL_dT = 0;
                for (int i = 0; i < 3; i++) dL_dpixel[i] = dpix[i * N + pix_id];
                real last_alpha = 0, last_color[3] = {0, 0, 0};
                for (int e = (int)r1 - 1; e >= (int)r0; e--) {
                    contributor--;
                    if (contributor >= (uint32_t)last_contributor) continue;
                    uint32_t g = o->point_list[e];
                    const real* Tu = TM + 9 * g; const real* Tv = Tu + 3; const real* Tw = Tu + 6;
                    real k[3] = {pixf[0] * Tw[0] - Tu[0], pixf[0] * Tw[1] - Tu[1], pixf[0] * Tw[2] - Tu[2]};
                    real l[3] = {pixf[1] * Tw[0] - Tv[0], pixf[1] * Tw[1] - Tv[1], pixf[1] * Tw[2] - Tv[2]};
                    real p[3] = {k[1] * l[2] - k[2] * l[1], k[2] * l[0] - k[0] * l[2], k[0] * l[1] - k[1] * l[0]};
                    if (p[2] == 0) continue;
                    real s[2] = {p[0] / p[2], p[1] / p[2]};
                    real rho3d = s[0] * s[0] + s[1] * s[1];
                    real d[2] = {o->xy[2 * g] - pixf[0], o->xy[2 * g + 1] - pixf[1]};
                    real rho2d = FILTER_INV_SQUARE * (d[0] * d[0] + d[1] * d[1]);
                    real rho = rho3d < rho2d ? rho3d : rho2d;
                    real c_d = (rho3d <= rho2d) ? (s[0] * Tw[0] + s[1] * Tw[1]) + Tw[2] : Tw[2];
                    if (c_d < NEAR_N) continue;
                    const real* no = o->normal_opacity + 4 * g;
                    real opa = no[3];
                    real power = R(-0.5) * rho;
                    if (power > 0) continue;
                    const real G = exp_r(power);
                    real alpha = opa * G;
                    if (alpha > R(0.99)) alpha = R(0.99);
                    if (alpha < R(1.0) / R(255.0)) continue;
                    T = T / (R(1.) - alpha);
                    const real dchannel_dcolor = alpha * T;
                    real dL_dalpha = 0;
                    for (int ch = 0; ch < 3; ch++) {
                        const real c = feat[3 * g + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (R(1.) - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum_rec[ch]) * dL_dpixel[ch];
                        atomic_add_d(&aC[3 * g + ch], (double)(dchannel_dcolor * dL_dpixel[ch]));
                    }
                    real dL_dz = 0, dL_dweight = 0;
                    const real m_d = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N / c_d);
                    const real dmd_dd = (FAR_N * NEAR_N) / ((FAR_N - NEAR_N) * c_d * c_d);
                    if (contributor == (uint32_t)(median_contributor - 1)) dL_dz += dL_dmedian_depth;
                    dL_dweight += (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_dreg;
                    dL_dalpha += dL_dweight - last_dL_dT;
                    last_dL_dT = dL_dweight * alpha + (1 - alpha) * last_dL_dT;
                    const real dL_dmd = R(2.0) * (T * alpha) * (m_d * final_A - final_D) * dL_dreg;
                    dL_dz += dL_dmd * dmd_dd;
                    accum_depth_rec = last_alpha * last_depth + (R(1.) - last_alpha) * accum_depth_rec;
                    last_depth = c_d;
                    dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
                    accum_alpha_rec = last_alpha * R(1.0) + (R(1.) - last_alpha) * accum_alpha_rec;
                    dL_dalpha += (1 - accum_alpha_rec) * dL_daccum;
                    for (int ch = 0; ch < 3; ch++) {
                        accum_normal_rec[ch] = last_alpha * last_normal[ch] + (R(1.) - last_alpha) * accum_normal_rec[ch];
                        last_normal[ch] = no[ch];
                        dL_dalpha += (no[ch] - accum_normal_rec[ch]) * dL_dnormal2D[ch];
                        /* S/bwd:379-381: both adds, the second for EVERY contributing splat (quirk Q1) */
                        atomic_add_d(&aN[3 * g + ch], (double)(alpha * T * dL_dnormal2D[ch]));
                        atomic_add_d(&aN[3 * g + ch], (double)dL_dmedian_normal2D[ch]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    real bg_dot_dpixel = 0;
                    for (int i = 0; i < 3; i++) bg_dot_dpixel += o->bg[i] * dL_dpixel[i];
                    dL_dalpha += (-T_final / (R(1.) - alpha)) * bg_dot_dpixel;
                    const real dL_dG = opa * dL_dalpha;
                    dL_dz += alpha * T * dL_ddepth;
                    if (rho3d <= rho2d) {
                        const real dL_ds[2] = {dL_dG * -G * s[0] + dL_dz * Tw[0], dL_dG * -G * s[1] + dL_dz * Tw[1]};
                        const real dz_dTw[3] = {s[0], s[1], R(1.0)};
                        const real dsx_pz = dL_ds[0] / p[2], dsy_pz = dL_ds[1] / p[2];
                        const real dL_dp[3] = {dsx_pz, dsy_pz, -(dsx_pz * s[0] + dsy_pz * s[1])};
                        const real dL_dk[3] = {l[1] * dL_dp[2] - l[2] * dL_dp[1], l[2] * dL_dp[0] - l[0] * dL_dp[2], l[0] * dL_dp[1] - l[1] * dL_dp[0]};
                        const real dL_dl[3] = {dL_dp[1] * k[2] - dL_dp[2] * k[1], dL_dp[2] * k[0] - dL_dp[0] * k[2], dL_dp[0] * k[1] - dL_dp[1] * k[0]};
                        for (int c = 0; c < 3; c++) {
                            atomic_add_d(&aT[9 * g + 0 + c], (double)(-dL_dk[c]));
                            atomic_add_d(&aT[9 * g + 3 + c], (double)(-dL_dl[c]));
                            atomic_add_d(&aT[9 * g + 6 + c], (double)(pixf[0] * dL_dk[c] + pixf[1] * dL_dl[c] + dL_dz * dz_dTw[c]));
                        }
                    } else {
                        const real dG_ddelx = -G * FILTER_INV_SQUARE * d[0];
                        const real dG_ddely = -G * FILTER_INV_SQUARE * d[1];
                        atomic_add_d(&aM2[2 * g + 0], (double)(dL_dG * dG_ddelx));
                        atomic_add_d(&aM2[2 * g + 1], (double)(dL_dG * dG_ddely));
                        atomic_add_d(&aT[9 * g + 8], (double)dL_dz);
                    }
                    atomic_add_d(&aO[g], (double)(G * dL_dalpha));
                }
            }
    }

    /* ---- preprocess backward, S/bwd:582-637 with compute_transmat_aabb S/bwd:450-580 ---- */
    /* S/bwd:614-615: W,H recomputed through float32 focal*tan*2 and truncated (quirk Q3) */
    const float fx32 = (float)W / (2.0f * (float)o->tan_fovx), fy32 = (float)H / (2.0f * (float)o->tan_fovy);
    const int Wb = (int)(fx32 * (float)o->tan_fovx * 2), Hb = (int)(fy32 * (float)o->tan_fovy * 2);
    real Pm[3][4];
    build_P(o->proj, Wb, Hb, Pm);
#pragma omp parallel for schedule(static)
    for (int idx = 0; idx < P; idx++) {
        real dT[3][3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) dT[i][j] = (real)aT[9 * idx + 3 * i + j];
        real dmean2D[2] = {(real)aM2[2 * idx], (real)aM2[2 * idx + 1]};
        real dmean3D[3] = {0, 0, 0}, dscale[2] = {0, 0}, drot[4] = {0, 0, 0, 0};
        real out_mean2D[3] = {dmean2D[0], dmean2D[1], 0};
        real dcol[3];
        for (int c = 0; c < 3; c++) dcol[c] = (real)aC[3 * idx + c];
        if (dL_dnormal_out) for (int c = 0; c < 3; c++) dL_dnormal_out[3 * idx + c] = (float)aN[3 * idx + c];
        if (dL_dmean2D_raw) { dL_dmean2D_raw[2 * idx] = (float)dmean2D[0]; dL_dmean2D_raw[2 * idx + 1] = (float)dmean2D[1]; }
        if (dL_dsh) for (int j = 0; j < M * 3; j++) dL_dsh[(size_t)idx * M * 3 + j] = 0;
        if (o->radii[idx] > 0) {
            real T[3][3], normal[3] = {0, 0, 0}, Rm[3][3];
            const int precomp = (scales == NULL); /* S/bwd:616 */
            if (precomp) {
                for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) T[i][j] = TM[9 * idx + 3 * i + j];
            } else {
                /* S/bwd:484-512: scale_modifier ignored (quirk Q4), Wb/Hb from Q3 */
                compute_transmat(means + 3 * idx, scales + 2 * idx, R(1.0), rots + 4 * idx, o->proj, o->view, Wb, Hb, 1, T, normal);
                quat_to_rotmat(rots + 4 * idx, Rm);
            }
            int stop = 0;
            if (dmean2D[0] != 0 || dmean2D[1] != 0) { /* S/bwd:522-550 */
                real t[3] = {R(9.0), R(9.0), R(-1.0)};
                real d = t[0] * T[2][0] * T[2][0] + t[1] * T[2][1] * T[2][1] + t[2] * T[2][2] * T[2][2];
                real f[3] = {t[0] * (R(1.0) / d), t[1] * (R(1.0) / d), t[2] * (R(1.0) / d)};
                real dT0[3], dT1[3], dT3[3], dLdf[3];
                for (int c = 0; c < 3; c++) {
                    dT0[c] = dmean2D[0] * f[c] * T[2][c];
                    dT1[c] = dmean2D[1] * f[c] * T[2][c];
                    dT3[c] = dmean2D[0] * f[c] * T[0][c] + dmean2D[1] * f[c] * T[1][c];
                    dLdf[c] = dmean2D[0] * T[0][c] * T[2][c] + dmean2D[1] * T[1][c] * T[2][c];
                }
                real dLdd = (dLdf[0] * f[0] + dLdf[1] * f[1] + dLdf[2] * f[2]) * (R(-1.0) / d);
                for (int c = 0; c < 3; c++) {
                    dT3[c] += dLdd * (t[c] * T[2][c] * R(2.0));
                    dT[0][c] += dT0[c]; dT[1][c] += dT1[c]; dT[2][c] += dT3[c];
                }
                if (precomp) stop = 1; /* dL_dTs written back, S/bwd:538-549 */
            }
            if (precomp) stop = 1;
            if (!stop) {
                /* dL_dM = P * transpose(dL_dT), S/bwd:555 */
                real dM[3][4];
                for (int j = 0; j < 3; j++) for (int r = 0; r < 4; r++)
                    dM[j][r] = Pm[0][r] * dT[0][j] + Pm[1][r] * dT[1][j] + Pm[2][r] * dT[2][j];
                real dn[3] = {(real)aN[3 * idx], (real)aN[3 * idx + 1], (real)aN[3 * idx + 2]};
                real dtn[3];
                xformvec43T(o->view, dn, dtn);
                real pv[3];
                xform43(o->view, means + 3 * idx, pv);
                real cosv = -(pv[0] * normal[0] + pv[1] * normal[1] + pv[2] * normal[2]);
                real mult = cosv > 0 ? R(1.) : R(-1.);
                for (int c = 0; c < 3; c++) dtn[c] *= mult;
                real dRS[3][3], dR[3][3];
                for (int r = 0; r < 3; r++) { dRS[0][r] = dM[0][r]; dRS[1][r] = dM[1][r]; dRS[2][r] = dtn[r]; }
                for (int r = 0; r < 3; r++) {
                    dR[0][r] = dRS[0][r] * scales[2 * idx]; dR[1][r] = dRS[1][r] * scales[2 * idx + 1]; dR[2][r] = dRS[2][r];
                }
                quat_to_rotmat_vjp(rots + 4 * idx, dR, drot);
                dscale[0] = dRS[0][0] * Rm[0][0] + dRS[0][1] * Rm[0][1] + dRS[0][2] * Rm[0][2];
                dscale[1] = dRS[1][0] * Rm[1][0] + dRS[1][1] * Rm[1][1] + dRS[1][2] * Rm[1][2];
                for (int c = 0; c < 3; c++) dmean3D[c] = dM[2][c];
            }
            real dT_out[9];
            /* returned dL_dtransMat: raw render accumulation unless precomp AND the
             * aabb branch ran (only then S/bwd:538-547 stores the updated matrix) */
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
                dT_out[3 * i + j] = (precomp && (dmean2D[0] != 0 || dmean2D[1] != 0)) ? dT[i][j] : (real)aT[9 * idx + 3 * i + j];
            if (shs) { /* S/bwd:630-631 */
                real* dshs_tmp = (real*)calloc((size_t)M * 3, sizeof(real));
                color_from_sh_bwd(o->D, means + 3 * idx, o->campos, shs + (size_t)idx * M * 3,
                                  o->clamped + 3 * idx, dcol, dmean3D, dshs_tmp);
                for (int j = 0; j < M * 3; j++) dL_dsh[(size_t)idx * M * 3 + j] = (float)dshs_tmp[j];
                free(dshs_tmp);
            }
            /* densification hack, S/bwd:633-636 (reads the stored dL_dtransMat and T[8]) */
            real depth = TM[9 * idx + 8];
            out_mean2D[0] = (real)((double)(dT_out[2] * depth) * 0.5 * (double)(float)Wb);
            out_mean2D[1] = (real)((double)(dT_out[5] * depth) * 0.5 * (double)(float)Hb);
            for (int j = 0; j < 9; j++) dL_dtransMat[9 * idx + j] = (float)dT_out[j];
        } else {
            for (int j = 0; j < 9; j++) dL_dtransMat[9 * idx + j] = (float)aT[9 * idx + j];
        }
        for (int c = 0; c < 3; c++) dL_dmean2D[3 * idx + c] = (float)out_mean2D[c];
        for (int c = 0; c < 3; c++) dL_dcolors[3 * idx + c] = (float)dcol[c];
        dL_dopacity[idx] = (float)aO[idx];
        for (int c = 0; c < 3; c++) dL_dmean3D[3 * idx + c] = (float)dmean3D[c];
        for (int c = 0; c < 2; c++) dL_dscales[2 * idx + c] = (float)dscale[c];
        for (int c = 0; c < 4; c++) dL_drots[4 * idx + c] = (float)drot[c];
    }
    free(aT); free(aM2); free(aN); free(aO); free(aC);
    free(means); free(shs); free(colors); free(scales); free(rots); free(Tpre); free(dpix); free(doth);
}

/* S/impl:54-66,141-153 markVisible */
void orc_mark_visible(int P, const float* means3D, const float* view, uint8_t* present) {
    for (int i = 0; i < P; i++) {
        real m[16], p[3] = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]}, pv[3];
        for (int j = 0; j < 16; j++) m[j] = view[j];
        xform43(m, p, pv);
        present[i] = pv[2] > R(0.2);
    }
}
