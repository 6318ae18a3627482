/* gsr_b200.h -- C ABI of libgsr_b200.so, the B200 (sm_100a) rasterizer behind
 * GS-SR's GaussianRasterizer / GaussianRasterizationSettings Python API.
 *
 * Every entry point replaces one function of the reference's native layer
 * (cited per function; S/ = submodules/diff-surfel-rasterization, G/ =
 * diff-gaussian-rasterization, L/ = diff-plane-rasterization, F/ =
 * scaffold-filter, K/ = simple-knn under /root/reference).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host;
 *     float32 / int32, densely packed, row-major, exactly the layouts the
 *     reference's torch glue passes down (S/rasterize_points.cu:108-132);
 *   - a NULL pointer means "not provided", like the empty tensors the
 *     reference wrapper turns into nullptr (S/diff_surfel_rasterization/__init__.py:198-208);
 *   - `stream` is a cudaStream_t; all work is enqueued on it (the reference
 *     uses the legacy default stream);
 *   - the library never allocates device memory: scratch comes from the three
 *     resize callbacks, the same protocol as the reference's
 *     std::function<char*(size_t)> geometryBuffer/binningBuffer/imageBuffer
 *     (S/cuda_rasterizer/rasterizer.h:31-34, S/rasterize_points.cu:31-37).
 *     A callback must return a device pointer valid until the matching
 *     backward call has been enqueued (>= 256-byte aligned), or NULL on failure;
 *   - return value: >= 0 on success (forward: binning layout size >= num_rendered), < 0 = GSR_E_*;
 *     gsr_last_error() returns a thread-local message.  No C++ exception
 *     crosses the boundary.
 */
#ifndef GSR_B200_H_
#define GSR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_ABI_VERSION 1

enum {
    GSR_OK = 0,
    GSR_E_INVALID = -1,     /* bad argument (shape, NULL where required) */
    GSR_E_CUDA = -2,        /* CUDA runtime error, see gsr_last_error() */
    GSR_E_ALLOC = -3,       /* a resize callback returned NULL */
    GSR_E_PREFILTERED = -4, /* prefiltered=1 but a point failed the near cull
                               (the reference __trap()s, S/cuda_rasterizer/auxiliary.h:204-208) */
    GSR_E_OVERFLOW = -5     /* num_rendered does not fit 32-bit indexing */
};

/* Resize callback: return a device buffer of at least `bytes` bytes. */
typedef char* (*gsr_buffer_fn)(void* user, size_t bytes);

int gsr_abi_version(void);
const char* gsr_last_error(void);
/* Name of the GPU architecture the library was compiled for ("sm_100a"). */
const char* gsr_build_arch(void);

/* ---- measurement / debug hooks (no reference counterpart) ------------------ */

/* Per-kernel device timing with cudaEvents recorded on the caller's stream around
 * each launch of the next forward/backward calls (bench.py's roofline numbers).
 * gsr_profile_read() fills GSR_PROF_SLOTS floats (milliseconds, -1 = not run) for the
 * most recent call of each kernel; it blocks on the recorded events.
 * Events are kept per device (enable / read act on the CURRENT device); the on/off switch is
 * process-wide.  A measurement hook, not part of the re-entrant per-stream path; off by default. */
enum {
    GSR_PROF_PREPROCESS_FWD = 0,
    GSR_PROF_SCAN = 1,
    GSR_PROF_DUPLICATE = 2,
    GSR_PROF_SORT = 3,
    GSR_PROF_BUILD_RECORDS = 4,
    GSR_PROF_RENDER_FWD = 5,
    GSR_PROF_RENDER_BWD = 6,
    GSR_PROF_PREPROCESS_BWD = 7,
    GSR_PROF_SLOTS = 16
};
int gsr_profile_enable(int on);
int gsr_profile_read(float* ms_host);

/* Runtime switches for validation.  "no_cull" = 1 disables the conservative
 * contribution boxes so every (pixel, splat) pair of a tile is evaluated like the
 * reference does; results must not change (tests/test_surfel_gpu.py).
 * "no_used_bits" = 1: the backward repeats the cull test instead of reading the forward's marks.
 * "force_capacity" = n > 0: lay the binning buffer out for n entries regardless of history (exercises
 * the re-run taken when num_rendered exceeds the predicted capacity); 0 restores the prediction. */
int gsr_set_option(const char* name, int value);

/* num_rendered WITHOUT a stream sync.  The reference blocks on a device->host copy of num_rendered in
 * the middle of every forward to size its binning buffer (S/cuda_rasterizer/rasterizer_impl.cu:278-285).
 * The *_forward entry points below instead lay the binning buffer out for a capacity predicted from
 * earlier frames of the same (device, rasterizer, resolution), enqueue every kernel of the forward
 * (lists clamped to the capacity), and then wait -- on a pinned host word the scan kernel writes, not on
 * the stream -- for the true count: the GPU never idles, no cudaStreamSynchronize / cudaMemcpy is
 * issued.  If the count exceeds the capacity (or there is no history yet) binning and rendering are
 * enqueued again behind the clamped run with an exactly sized buffer; results are always exact.
 * Consequence for callers: the forward's return value is the number of list entries the binning buffer
 * was LAID OUT for (>= num_rendered).  It must be passed to the matching backward unchanged, which is
 * all the reference wrapper does with it (S/diff_surfel_rasterization/__init__.py:97,121).  The true
 * num_rendered of this host thread's most recent forward: */
int gsr_last_num_rendered(void);

/* CUDA-graph capture (no reference counterpart: the reference forward contains a blocking cudaMemcpy and cannot be
 * captured).  When the stream handed to a *_forward entry point is being captured (cudaStreamIsCapturing), the call
 * records every kernel of the forward exactly once and returns without waiting for num_rendered: the binning buffer
 * is laid out for "capture_margin" percent (gsr_set_option, default 150, minimum 100) of the num_rendered history of
 * earlier EAGER forwards of the same (device, rasterizer, resolution) -- at least one is required, otherwise the call
 * fails with GSR_E_INVALID.  The buffer callbacks run at capture time (with torch: inside the graph's private pool).
 * A replay whose num_rendered outgrows the captured capacity renders a TRUNCATED frame (lists clamped to the
 * capacity; memory safe) and leaves that count in a sticky host word of the capturing thread:
 *   gsr_capture_overflow(reset) returns it (0 = every replay fitted; 0xffffffff = a `prefiltered` violation) and
 *   clears it when reset != 0.  Read it after the replay has completed (any stream / event sync of the caller),
 *   from the host thread that captured the graph; on overflow raise the margin or run an eager forward (which
 *   refreshes the history) and capture again.  Backward entry points have no host interaction and capture as they are. */
unsigned int gsr_capture_overflow(int reset);

/* Decision audit of a finished forward (verification only, no reference counterpart).  Re-walks the record stream
 * kept in the forward's binning / image buffers with the render kernels' own arithmetic and writes, per pixel, the
 * smallest relative distance of any blend decision to its threshold:
 *   surfel  margins (5,H,W): alpha vs 1/255 | T(1-alpha) vs 1e-4 | T vs 0.5 | depth vs 0.2 | rho3d vs rho2d
 *           (the tests of S/cuda_rasterizer/forward.cu:362-389,402)
 *   ewa     margins (3,H,W): alpha vs 1/255 | T(1-alpha) vs 1e-4 | T vs 0.5   (G/forward.cu:336-356, L/forward.cu:381)
 *   info (2,H,W) int: number of blended splats, Gaussian index of the last one (-1 = none)
 *   mismatches (1 int, device): pixels whose replayed final_T / last contributor differ from the forward's (must be 0).
 * Parity tests use it to accept a deviating pixel only when one of its decisions was marginal
 * (tests/test_fullsize_parity_gpu.py).  P, R, width, height (and render_geo) as passed to / returned by the forward. */
int gsr_surfel_audit(int P, int R, int width, int height, char* binning_buffer, char* image_buffer,
                     float* margins, int* info, int* mismatches, void* stream);
int gsr_ewa_audit(int P, int R, int width, int height, int render_geo, char* binning_buffer, char* image_buffer,
                  float* margins, int* info, int* mismatches, void* stream);

/* ---- 2DGS surfel rasterizer: diff_surfel_rasterization ------------------ */

/* Replaces CudaRasterizer::Rasterizer::forward
 * (S/cuda_rasterizer/rasterizer.h:31-56, S/cuda_rasterizer/rasterizer_impl.cu:198-342)
 * as called from RasterizeGaussiansCUDA (S/rasterize_points.cu:39-135).
 *   out_color (3,H,W), out_others (11,H,W), radii (P) are fully written.
 *   Exactly one of shs / colors_precomp and one of (scales+rotations) /
 *   transMat_precomp must be non-NULL.  scales is (P,2) packed.
 * Returns the binning layout size (>= num_rendered, see gsr_last_num_rendered). */
int gsr_surfel_forward(gsr_buffer_fn geometryBuffer, gsr_buffer_fn binningBuffer,
                       gsr_buffer_fn imageBuffer, void* user,
                       int P, int D, int M, const float* background, int width, int height,
                       const float* means3D, const float* shs, const float* colors_precomp,
                       const float* opacities, const float* scales, float scale_modifier,
                       const float* rotations, const float* transMat_precomp,
                       const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                       float tan_fovx, float tan_fovy, int prefiltered,
                       float* out_color, float* out_others, int* radii, int debug, void* stream);

/* Replaces CudaRasterizer::Rasterizer::backward
 * (S/cuda_rasterizer/rasterizer.h:58-88, S/cuda_rasterizer/rasterizer_impl.cu:346-448)
 * as called from RasterizeGaussiansBackwardCUDA (S/rasterize_points.cu:137-234).
 * geom/binning/image buffers are the ones handed out during the forward call,
 * R its return value.  All dL_* outputs are fully written (zeros where the
 * reference leaves its torch::zeros untouched); dL_dnormal (P,3) is scratch the
 * reference also exposes at this level.  dL_dsh may be NULL when M == 0. */
int gsr_surfel_backward(int P, int D, int M, int R, const float* background, int width, int height,
                        const float* means3D, const float* shs, const float* colors_precomp,
                        const float* scales, float scale_modifier, const float* rotations,
                        const float* transMat_precomp, const float* viewmatrix,
                        const float* projmatrix, const float* campos, float tan_fovx,
                        float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer,
                        char* image_buffer, const float* dL_dpix, const float* dL_dothers,
                        float* dL_dmean2D, float* dL_dnormal, float* dL_dopacity, float* dL_dcolor,
                        float* dL_dmean3D, float* dL_dtransMat, float* dL_dsh, float* dL_dscale,
                        float* dL_drot, int debug, void* stream);

/* Replaces CudaRasterizer::Rasterizer::markVisible
 * (S/cuda_rasterizer/rasterizer_impl.cu:141-153; identical in G/ and L/).
 * present: P bytes (bool). */
int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, uint8_t* present, void* stream);


/* ---- 3DGS rasterizer: diff_gaussian_rasterization -------------------------- */

/* Replaces CudaRasterizer::Rasterizer::forward
 * (G/cuda_rasterizer/rasterizer.h:31-53, G/cuda_rasterizer/rasterizer_impl.cu:198-337)
 * as called from RasterizeGaussiansCUDA (G/rasterize_points.cu:35-113).
 *   out_color (3,H,W) and radii (P) are fully written.  Exactly one of shs /
 *   colors_precomp and one of (scales+rotations) / cov3D_precomp must be non-NULL;
 *   scales is (P,3) packed, cov3D_precomp (P,6) upper-triangular.
 * Returns the binning layout size (>= num_rendered, see gsr_last_num_rendered). */
int gsr_gaussian_forward(gsr_buffer_fn geometryBuffer, gsr_buffer_fn binningBuffer,
                         gsr_buffer_fn imageBuffer, void* user,
                         int P, int D, int M, const float* background, int width, int height,
                         const float* means3D, const float* shs, const float* colors_precomp,
                         const float* opacities, const float* scales, float scale_modifier,
                         const float* rotations, const float* cov3D_precomp,
                         const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                         float tan_fovx, float tan_fovy, int prefiltered,
                         float* out_color, int* radii, int debug, void* stream);

/* Replaces CudaRasterizer::Rasterizer::backward
 * (G/cuda_rasterizer/rasterizer.h:55-85, G/cuda_rasterizer/rasterizer_impl.cu:341-433)
 * as called from RasterizeGaussiansBackwardCUDA (G/rasterize_points.cu:115-197).
 * All dL_* outputs are fully written; dL_dconic (P,4: x, y, 0, w) is scratch the
 * reference also exposes at this level and may be NULL, as may dL_dsh (M == 0),
 * dL_dscale and dL_drot (cov3D_precomp given). */
int gsr_gaussian_backward(int P, int D, int M, int R, const float* background, int width, int height,
                          const float* means3D, const float* shs, const float* colors_precomp,
                          const float* scales, float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* viewmatrix,
                          const float* projmatrix, const float* campos, float tan_fovx,
                          float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer,
                          char* image_buffer, const float* dL_dpix, float* dL_dmean2D,
                          float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                          float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                          int debug, void* stream);

/* ---- PGSR plane rasterizer: diff_plane_rasterization ------------------------ */

/* Replaces CudaRasterizer::Rasterizer::forward
 * (L/cuda_rasterizer/rasterizer.h:31-61, L/cuda_rasterizer/rasterizer_impl.cu:200-352)
 * as called from RasterizeGaussiansCUDA (L/rasterize_points.cu:35-125).
 *   all_map (P,5) may be NULL when render_geo == 0.  out_observe (P) int32 is fully
 *   written (count of pixels on which the Gaussian was blended with T > 0.5);
 *   out_all_map (5,H,W) and out_plane_depth (1,H,W) are written when render_geo != 0
 *   and zero-filled otherwise. */
int gsr_plane_forward(gsr_buffer_fn geometryBuffer, gsr_buffer_fn binningBuffer,
                      gsr_buffer_fn imageBuffer, void* user,
                      int P, int D, int M, const float* background, int width, int height,
                      const float* means3D, const float* shs, const float* colors_precomp,
                      const float* opacities, const float* scales, float scale_modifier,
                      const float* rotations, const float* cov3D_precomp, const float* all_map,
                      const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                      float tan_fovx, float tan_fovy, int prefiltered,
                      float* out_color, int* radii, int* out_observe, float* out_all_map,
                      float* out_plane_depth, int render_geo, int debug, void* stream);

/* Replaces CudaRasterizer::Rasterizer::backward
 * (L/cuda_rasterizer/rasterizer.h:63-99, L/cuda_rasterizer/rasterizer_impl.cu:356-462)
 * as called from RasterizeGaussiansBackwardCUDA (L/rasterize_points.cu:127-231).
 * all_map_pixels = the forward's out_all_map.  dL_dmean2D_abs (P,3) and dL_dall_map
 * (P,5) are fully written (zeros when render_geo == 0 for the latter). */
int gsr_plane_backward(int P, int D, int M, int R, const float* background,
                       const float* all_map_pixels, int width, int height,
                       const float* means3D, const float* shs, const float* colors_precomp,
                       const float* all_maps, const float* scales, float scale_modifier,
                       const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                       const float* projmatrix, const float* campos, float tan_fovx,
                       float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer,
                       char* image_buffer, const float* dL_dpix, const float* dL_dout_all_map,
                       const float* dL_dout_plane_depth, float* dL_dmean2D, float* dL_dmean2D_abs,
                       float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D,
                       float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                       float* dL_dall_map, int render_geo, int debug, void* stream);

/* ---- scaffold_filter ------------------------------------------------------------ */

/* Replaces CudaRasterizer::Rasterizer::visible_filter
 * (F/cuda_rasterizer/rasterizer.h:56-72, F/cuda_rasterizer/rasterizer_impl.cu:340-396)
 * as called from RasterizeGaussiansfilterCUDA (F/rasterize_points.cu:220-284).
 * radii (P) int32 is fully written; needs no scratch (the reference still resizes
 * its geometry and image chunks).  `prefiltered` is accepted for signature parity and ignored
 * (no flag word / read-back on this asynchronous entry point; the reference traps in-kernel). */
int gsr_visible_filter(int P, int width, int height, const float* means3D, const float* scales,
                       float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const float* viewmatrix, const float* projmatrix, float tan_fovx,
                       float tan_fovy, int prefiltered, int* radii, int debug, void* stream);

/* ---- simple_knn ---------------------------------------------------------------------- */

/* Replaces SimpleKNN::knn (K/simple_knn.h:15-19, K/simple_knn.cu:186-221) as called from
 * distCUDA2 (K/spatial.cu:14-25): meanDists[i] = mean of the 3 smallest squared
 * distances from points[i] to the other points.  `workspace` must hold
 * gsr_dist2_knn3_workspace(P) bytes of device memory (the reference cudaMallocs /
 * thrust-allocates per call and synchronises twice; this call is fully asynchronous). */
size_t gsr_dist2_knn3_workspace(int P);
int gsr_dist2_knn3(int P, const float* points, float* meanDists, void* workspace, void* stream);

/* ---- TSDF fusion (mesh extraction) --------------------------------------------------- */

/* One view of the fusion: the camera's full_proj_transform ((4,4) row-major, row-vector
 * convention: clip = [x y z 1] @ M, gssr/cameras/__init__.py:85-88) and its rendered maps
 * (device pointers; depth (H,W), rgb (3,H,W) or NULL when colours are not requested). 96 bytes. */
typedef struct gsr_tsdf_view {
    float full_proj[16];
    int width, height;
    int reserved0, reserved1;
    const float* depth;
    const float* rgb;
} gsr_tsdf_view;

/* Replaces the per-view torch loop of GaussianExtractor.extract_mesh_unbounded
 * (gssr/utils/mesh_utils.py:195-246: compute_sdf_perframe + compute_unbounded_tsdf), the
 * function GS-SR hands to marching cubes as `sdf` (:254, gssr/utils/mcube_utils.py:57-68) and
 * uses to colour the mesh vertices (:275).  For every sample point i (samples (n,3)):
 *   contracted != 0: truncation 5*voxel_size/(2 - min(|x|,1.9)) outside the unit ball and
 *   x <- uncontract(x) * radius + center (:187-193, :215-219, :248-250); else truncation 5*voxel_size;
 *   for each view in order: project, mask_proj, bilinear border/align_corners sample of depth
 *   (and rgb), sdf = depth - z, fused where sdf > -trunc with the running mean of :236-241.
 * init != 0 starts from tsdf = 1, weight = 1, rgb = 0 (the reference's initial state) ; init == 0
 * continues from the values in tsdf / weights / rgb (views streamed in several calls).
 * `views` is a DEVICE array of nviews descriptors.  weights may be NULL when init != 0;
 * rgb == NULL skips the colour fusion (the reference computes and discards it when
 * return_rgb=False).  Asynchronous on `stream`; allocates nothing. */
int gsr_tsdf_fuse(long long n, const float* samples, int contracted, const float* center_host3,
                  float radius, float voxel_size, int nviews, const gsr_tsdf_view* views, int init,
                  float* tsdf, float* weights, float* rgb, void* stream);

/* Bounded TSDF volume: the fusion of GaussianExtractor.extract_mesh_bounded (gssr/utils/mesh_utils.py:138-179) and of the
 * multi-tile extract_mesh_split.py:81-119 on a dense nx x ny x nz lattice  origin + (ix, iy, iz) * voxel_size  (x fastest).
 * The reference delegates this to Open3D's ScalableTSDFVolume (absent third-party dependency, no vectors in the
 * reference: PARITY UNPINNED against Open3D); the integration rule used here is the reference's own torch rule
 * (gsr_tsdf_fuse above: project with full_proj_transform, bilinear depth sample, sdf = depth - z, running mean where
 * sdf > -sdf_trunc, start tsdf = 1 / weight = 1 / rgb = 0) with the two masks of the bounded path: a view is skipped
 * where its sampled depth is <= 0 (masked background, :160-162) or > depth_trunc (:165-170).
 * init != 0 starts a fresh volume, init == 0 continues from tsdf / weights / rgb ((nz,ny,nx) floats, rgb (nz,ny,nx,3) or NULL).
 * Volumes fused on different GPUs (one VastGaussian tile each) combine exactly: sum over ranks of (tsdf*w - 1, w - 1, rgb*w). */
int gsr_tsdf_integrate_grid(int nx, int ny, int nz, const float* origin_host3, float voxel_size, float sdf_trunc,
                            float depth_trunc, int nviews, const gsr_tsdf_view* views, int init, float* tsdf,
                            float* weights, float* rgb, void* stream);

/* Triangle mesh of the level set of a TSDF lattice (marching cubes): the step the reference hands its fused volume to --
 * volume.extract_triangle_mesh() of Open3D's ScalableTSDFVolume (gssr/utils/mesh_utils.py:178, extract_mesh_split.py:119)
 * and skimage.measure.marching_cubes(level=0) on host chunks (gssr/utils/mcube_utils.py:71-80).  Both are absent
 * third-party dependencies: PARITY UNPINNED against them; conventions follow Open3D's extractor (inside = f < level; a cell
 * yields triangles only when all eight corners have weight > min_weight; vertices at f0 / (f0 - f1) of a lattice edge, shared
 * between cells; colours interpolated with the same weight) and the case table is consistent across cell faces, so closed
 * surfaces come out closed.  Lattice as in gsr_tsdf_integrate_grid: (nz, ny, nx) floats, x fastest, point (ix, iy, iz) =
 * origin + (ix, iy, iz) * voxel_size.  Two calls, the caller allocates in between:
 *   gsr_mc_count   classifies the cells into `workspace` (gsr_mc_workspace_bytes(nx, ny, nz) device bytes, 256-byte aligned)
 *                  and returns the vertex / triangle counts; weight == NULL = every corner counts.  Synchronises `stream`.
 *   gsr_mc_emit    writes verts (nverts, 3) f32, colors (nverts, 3) f32 (iff rgb (nz, ny, nx, 3) is given) and faces (ntris, 3)
 *                  i32 from the same tsdf / level / workspace.  Asynchronous.  Output order is lattice order: vertices by
 *                  (owner voxel, axis), faces by (cell, table order), normals towards f > level -- deterministic.
 * Lattices of 2^31 voxels or more, or meshes beyond 32-bit indices, return GSR_E_OVERFLOW: extract them in slabs. */
size_t gsr_mc_workspace_bytes(int nx, int ny, int nz);
int gsr_mc_count(int nx, int ny, int nz, const float* tsdf, const float* weight, float min_weight, float level,
                 void* workspace, long long* nverts_host, long long* ntris_host, void* stream);
int gsr_mc_emit(int nx, int ny, int nz, const float* tsdf, const float* rgb, float level, const float* origin_host3,
                float voxel_size, const void* workspace, float* verts, float* colors, int* faces, void* stream);

/* Mesh clean-up of GS-SR's post_process_mesh (gssr/utils/mesh_utils.py:27-49; Open3D cluster_connected_triangles /
 * remove_triangles_by_mask / remove_unreferenced_vertices / remove_degenerate_triangles on the host; Open3D absent: PARITY
 * UNPINNED, the steps are restated by oracle/mesh_clusters_oracle.py).
 *   gsr_mesh_clusters       connected components (lock-free union-find over the vertices, one thread per triangle):
 *                           vertex_root[v] = smallest vertex id of v's component, tri_root[t] = component of triangle t,
 *                           root_ntris[r] / root_area[r] = triangles / area of the component whose smallest vertex is r
 *                           (0 at every other index; root_area and verts may both be NULL).  Triangles are joined through
 *                           shared vertices (Open3D: shared edges -- the same clusters unless two sheets touch in one point).
 *                           Synchronises `stream`; GSR_E_INVALID if a triangle indexes outside [0, nverts).
 *   gsr_mesh_cluster_sizes  the cluster sizes (non-zero entries of root_ntris) gathered into sizes[0 .. nclusters), any order
 *                           -- what the reference sorts on the host to find its N-th largest cluster (mesh_utils.py:39);
 *                           `capacity` words, the last one is scratch; nclusters may exceed capacity - 1 (then only the
 *                           first capacity - 1 were stored: call again with a larger buffer).  Synchronises `stream`.
 *   gsr_mesh_keep_clusters  keep[t] = root_ntris[tri_root[t]] >= min_triangles  (mesh_utils.py:42: the inverse of
 *                           triangles_to_remove).
 *   gsr_mesh_filter_count / gsr_mesh_filter_emit   drop the triangles with tri_keep[t] == 0, then the vertices no kept triangle
 *                           references, then triangles with a repeated index (the reference's order of calls); survivors
 *                           keep their order.  count returns the new sizes (synchronises), the caller allocates, emit
 *                           writes verts_out / colors_out (iff colors) / faces_out (re-indexed).  workspace:
 *                           gsr_mesh_filter_workspace_bytes(nverts, ntris) device bytes, 256-byte aligned, shared by the two. */
int gsr_mesh_clusters(long long nverts, long long ntris, const float* verts, const int* faces, int* vertex_root, int* tri_root,
                      unsigned int* root_ntris, double* root_area, void* stream);
int gsr_mesh_cluster_sizes(long long nverts, const unsigned int* root_ntris, unsigned int* sizes, long long capacity,
                           long long* nclusters_host, void* stream);
int gsr_mesh_keep_clusters(long long ntris, const int* tri_root, const unsigned int* root_ntris, unsigned int min_triangles,
                           unsigned char* tri_keep, void* stream);
size_t gsr_mesh_filter_workspace_bytes(long long nverts, long long ntris);
int gsr_mesh_filter_count(long long nverts, long long ntris, const int* faces, const unsigned char* tri_keep, void* workspace,
                          long long* nverts_out_host, long long* ntris_out_host, void* stream);
int gsr_mesh_filter_emit(long long nverts, long long ntris, const float* verts, const float* colors, const int* faces,
                         const void* workspace, float* verts_out, float* colors_out, int* faces_out, void* stream);

/* ---- fused SSIM loss (image-space step right after the rasterizer) ---------------------- */

/* Replaces VanillaScene._ssim / .ssim (gssr/scene/vanilla_scene.py:32-61): five depthwise 11x11
 * conv2d + elementwise map in the forward, autograd's three more convolutions in the backward.
 * img1 / img2: (channels, height, width) device arrays; window11: HOST array with the 11 taps of the
 * normalised 1-D Gaussian (the reference's _gaussian(11, 1.5)); the 2-D window is their outer product.
 * Forward writes three derivative maps (same shape as img1: d ssim/d(G*x), d ssim/d(G*x^2), d ssim/d(G*xy))
 * for the backward and one partial sum of the SSIM map per CTA tile into tile_sums
 * (gsr_ssim_tile_count(...) floats); ssim.mean() = sum(tile_sums) / (channels*height*width).
 * Backward: dL_dimg1 = upstream * d ssim.mean()/d img1, upstream_dev = DEVICE scalar (autograd's
 * grad_output), so the call never synchronises. */
size_t gsr_ssim_tile_count(int channels, int height, int width);
int gsr_ssim_forward(int channels, int height, int width, const float* img1, const float* img2,
                     const float* window11_host, float* dmu1, float* dE11, float* dE12, float* tile_sums,
                     void* stream);
int gsr_ssim_backward(int channels, int height, int width, const float* img1, const float* img2,
                      const float* window11_host, const float* dmu1, const float* dE11, const float* dE12,
                      const float* upstream_dev, float* dL_dimg1, void* stream);

/* ---- fused allmap post-processing of the 2DGS scene (image-space step after the rasterizer) -- */

/* Replaces the torch block of TwoDGSScene.render after the rasterizer call
 * (gssr/scene/twodgs_scene.py:88-117) including depth_to_normal / depths_to_points
 * (gssr/utils/point_utils.py:9-37).  allmap: (11,H,W) as returned by the surfel rasterizer.
 * cam21: DEVICE array of 21 floats = K (3x3 row-major, rays_d = [x y 1] @ K with
 * K = intrins.inverse().T @ c2w[:3,:3].T), rays_o (3), R = world_view_transform[:3,:3] (3x3 row-major).
 * Forward writes render_normal (3,H,W; world space), surf_depth (1,H,W), surf_normal (3,H,W; times the
 * detached alpha, zero on the border).  Backward takes upstream gradients on those three outputs (any may be
 * NULL = zero), needs scratch6 = 6*H*W floats, and fully writes dL_dallmap (11,H,W). */
int gsr_surfel_post_forward(int height, int width, const float* allmap, const float* cam21, float depth_ratio,
                            float* render_normal, float* surf_depth, float* surf_normal, void* stream);
int gsr_surfel_post_backward(int height, int width, const float* allmap, const float* cam21, float depth_ratio,
                             const float* surf_depth, const float* g_render_normal, const float* g_surf_depth,
                             const float* g_surf_normal, float* scratch6, float* dL_dallmap, void* stream);

/* ---- PGSR depth -> normal -------------------------------------------------------
 * Replaces normal_from_depth_image(depth, intrinsic, extrinsic, offset=None)
 * (gssr/utils/graphics_utils.py:139-146 -> depth2point_cam :88-99, depth_pcd2normal :110-137) as called by
 * PGSRScene.render_normal (gssr/scene/pgsr_scene.py:227-238,320).
 *   depth (H,W) device; kinv: the 9 floats of inverse(K^T), row-major, on the DEVICE (the caller takes the inverse
 *   with the reference's own float32 torch op; no host read-back); weight (H,W) device or NULL: per-pixel factor fused into the output
 *   (PGSR multiplies by the detached alpha); normal (3,H,W) fully written, zero on the one-pixel border.
 * Backward: dL_dnormal (3,H,W) -> dL_ddepth (H,W) fully written; scratch = 6*H*W floats. */
int gsr_depth_normal_forward(int H, int W, const float* depth, const float* kinv, const float* weight,
                             float* normal, void* stream);
int gsr_depth_normal_backward(int H, int W, const float* depth, const float* kinv, const float* weight,
                              const float* dL_dnormal, float* scratch, float* dL_ddepth, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSR_B200_H_ */
