// abi.cu -- extern "C" entry points of libgsr_b200.so (see include/gsr_b200.h)
// and the host-side orchestration of the surfel pipeline.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <chrono>
#include <mutex>
#include "common.cuh"
#include "ewa_common.cuh"
#include "../../include/gsr_b200.h"

namespace gsr {

// ---- kernels / helpers defined in the other translation units ---------------
__global__ void surfel_preprocess_fwd(int, int, int, const float*, const float2*, const float4*, const float*,
                                      const float*, const float*, const bool, const ViewParams, const bool,
                                      const bool, int*, GeomRec*, CullRec*, float*, uint32_t*, uint32_t*, float*, uint8_t*, int*);
__global__ void surfel_preprocess_bwd(int, int, int, const float*, const float2*, const float4*, const float*,
                                      const bool, const ViewParams, const int, const int, const int*,
                                      const GeomRec*, const uint8_t*, const float*, float*, float*, float*,
                                      float*, float*, float*, float*, float*, float*);
__global__ void mark_visible_kernel(int, const float*, const ViewParams, uint8_t*);
cudaError_t launch_sh_forward(int, int, int, const float*, const float*, const float*, const int*, float*, uint8_t*, cudaStream_t);
cudaError_t launch_sh_backward(int, int, int, const float*, const float*, const float*, const uint8_t*, const int*, const float*,
                               float*, float*, cudaStream_t);
__global__ void tile_scan(int, const uint32_t*, uint32_t*, uint32_t*, uint32_t*, const int*, uint32_t, volatile uint32_t*, uint32_t, int);
__global__ void scatter_keys(int, const float*, const float*, int, const CullRec*, const float*, const int*,
                             const uint32_t*, int, int, uint32_t*, uint64_t*, uint32_t);
__global__ void sort_build_records(const uint32_t*, uint64_t*, const GeomRec*, const float*, int, int, int, float4*,
                                   size_t, int);
template <bool MARK>
__global__ void surfel_render_fwd(const uint32_t*, const float4*, size_t, int, int, int, const float*, float*,
                                  uint32_t*, float*, float*, float4*);
cudaError_t launch_surfel_render_bwd(bool, int, const uint32_t*, const float4*, size_t, int, int, int, const float*, const float*,
                                     const uint32_t*, const float*, const float*, float*, int, cudaStream_t);
int fwd_ctas_per_tile();
template <bool kRadiiOnly>
__global__ void ewa_preprocess_fwd(int, int, int, const float*, const float*, const float4*, const float*, const float*,
                                   const float*, const bool, const ViewParams, const float, const float, const float,
                                   const float, const bool, const bool, int*, EwaGeom*, CullRec*, float*, uint32_t*,
                                   uint32_t*, float*, uint8_t*, int*);
__global__ void ewa_preprocess_bwd(int, int, int, const float*, const float*, const float4*, const float*, const float*,
                                   const ViewParams, const float, const float, const float, const float, const int*,
                                   const uint8_t*, const float*, float*, float*, float*, float*, float*, float*, float*,
                                   float*, float*, float*, float*);
template <bool GEO>
__global__ void ewa_build_records(const uint32_t*, uint64_t*, const EwaGeom*, const float*, const float*, int, float4*, size_t);
template <bool GEO>
__global__ void ewa_render_fwd(const uint32_t*, const float4*, size_t, int, int, int, const float*, float, float, float*,
                               uint32_t*, float*, int*, float*, float*, float4*);
cudaError_t launch_ewa_render_bwd(int, bool, int, const uint32_t*, const float4*, size_t, int, int, int, const float*, float, float,
                                  const float*, const uint32_t*, const float*, const float*, const float*, const float*, float*, int, cudaStream_t);
cudaError_t launch_surfel_audit(int, const uint32_t*, const float4*, size_t, int, int, int, const float*, const uint32_t*, uint32_t,
                                float*, int*, int*, cudaStream_t);
cudaError_t launch_ewa_audit(int, const uint32_t*, const float4*, size_t, int, int, int, const float*, const uint32_t*, uint32_t,
                             float*, int*, int*, cudaStream_t);
cudaError_t launch_depth_normal_fwd(int, int, const float*, const float*, const float*, float*, cudaStream_t);
cudaError_t launch_depth_normal_bwd(int, int, const float*, const float*, const float*, const float*, float*, float*, cudaStream_t);
// ---- error string -------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- workspace layouts ----------------------------------------------------------
size_t GeomWs::carve(GeomWs& w, char* base, int P) {
    Carver c(base);
    size_t n = P > 0 ? (size_t)P : 1;
    w.geom = c.take<GeomRec>(n);
    w.cull = c.take<CullRec>(n);
    w.depths = c.take<float>(n);
    w.masks = c.take<uint32_t>(n);
    w.rgb = c.take<float>(3 * n);
    w.clamped = c.take<uint8_t>(3 * n);
    w.flags = c.take<int>(32);
    return c.used + 256;
}
size_t ImageWs::carve(ImageWs& w, char* base, int W, int H, int nT, int nC) {
    Carver c(base);
    size_t N = (size_t)W * H;
    size_t tiles = (size_t)((W + TILE - 1) / TILE) * ((H + TILE - 1) / TILE);
    w.final_T = c.take<float>(nT * N);
    w.n_contrib = c.take<uint32_t>(nC * N);
    w.tile_count = c.take<uint32_t>(tiles * TILE_CTR_STRIDE);
    w.tile_offset = c.take<uint32_t>(tiles + 1);
    w.tile_cursor = c.take<uint32_t>(tiles * TILE_CTR_STRIDE);
    w.total = c.take<uint32_t>(8);
    return c.used + 256;
}
size_t BinWs::carve(BinWs& w, char* base, int64_t R, int P, int nplanes, int gacc_stride) {
    Carver c(base);
    size_t n = R > 0 ? (size_t)R : 1;
    w.keys = c.take<uint64_t>(n);
    w.plane_stride = (n + 7) & ~size_t(7);
    w.planes = c.take<float4>(w.plane_stride * nplanes);
    w.gacc = c.take<float>((size_t)(P > 0 ? P : 1) * gacc_stride);
    return c.used + 256;
}

size_t EwaGeomWs::carve(EwaGeomWs& w, char* base, int P) {
    Carver c(base);
    size_t n = P > 0 ? (size_t)P : 1;
    w.geom = c.take<EwaGeom>(n);
    w.cull = c.take<CullRec>(n);
    w.depths = c.take<float>(n);
    w.masks = c.take<uint32_t>(n);
    w.rgb = c.take<float>(3 * n);
    w.clamped = c.take<uint8_t>(3 * n);
    w.flags = c.take<int>(32);
    return c.used + 256;
}

static inline char* align256(char* p) {
    return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255));
}

static ViewParams make_view(const float* view, const float* proj, const float* campos, int W, int H,
                            float scale_modifier) {
    ViewParams vc;
    vc.view = view; vc.proj = proj; vc.campos = campos;
    vc.W = W; vc.H = H;
    vc.gx = (W + TILE - 1) / TILE; vc.gy = (H + TILE - 1) / TILE;
    vc.scale_modifier = scale_modifier;
    return vc;
}

// ---- runtime options (debug / measurement only; defaults are the product path) ----
static int g_no_used_bits = 0;   // 1: never use the forward's "blended" bits (the path taken when P >= 2^23)
static int g_dbg = 0;         // developer timing experiments (results invalid when non-zero)
static int g_no_cull = -1;   // 1: contribution boxes disabled (every pair evaluated, as the reference does)
static bool no_cull() {
    if (g_no_cull < 0) { const char* e = getenv("GSR_NO_CULL"); g_no_cull = (e && e[0] == '1') ? 1 : 0; }
    return g_no_cull == 1;
}

// ---- host-side stage timing to stderr (GSR_HOST_TIMING=1; developer aid) ----
struct HostTimer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    char buf[512]; int len = 0;
    HostTimer() {
        static int v = -1;
        if (v < 0) { const char* e = getenv("GSR_HOST_TIMING"); v = (e && e[0] == '1') ? 1 : 0; }
        on = v == 1; t0 = std::chrono::steady_clock::now();
    }
    void mark(const char* name) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        len += snprintf(buf + len, sizeof(buf) - len, " %s=%.0fus", name,
                        std::chrono::duration<double, std::micro>(t1 - t0).count());
        t0 = t1;
    }
    void flush(const char* what) { if (on) fprintf(stderr, "[gsr host] %s:%s\n", what, buf); }
};

// ---- num_rendered without blocking the stream -------------------------------------------------
// The reference sizes its binning buffer from a blocking read-back of num_rendered in the middle of the forward
// (S/cuda_rasterizer/rasterizer_impl.cu:278-285): the GPU idles while the host waits, allocates and launches the rest.
// Here the host lays the buffer out for a capacity predicted from earlier frames of the same (device, rasterizer
// family, resolution), enqueues ALL forward kernels (they clamp their lists to the capacity), and only then waits
// for R, which tile_scan publishes into a pinned, mapped host slot -- by then the GPU is busy with the rest of the
// forward, so nothing idles and no cudaStreamSynchronize / cudaMemcpy is issued.  R <= capacity (the steady state):
// done.  R > capacity, or no history yet: the binning buffer is requested again with the exact size and tile_scan /
// scatter / sort / render are re-enqueued behind the clamped run -- still without a stream sync; outputs are simply
// overwritten.  The value returned to the caller (and passed back to the backward as R) is the capacity the buffer was
// laid out for; the true count is available from gsr_last_num_rendered().
struct RHistEntry { int dev, family, W, H; double hist; uint64_t stamp; };
static std::mutex g_hist_mu;
static RHistEntry g_hist[32];
static int g_hist_n = 0;
static uint64_t g_hist_clock = 0;
static int g_force_cap = 0;                      // option "force_capacity": tests of the overflow path
static int g_capture_margin = 150;               // option "capture_margin": capacity of a captured forward, % of the history
static thread_local int g_true_R = 0;

// CUDA-graph capture.  A forward that is being RECORDED (cudaStreamIsCapturing) cannot wait for num_rendered and cannot
// re-run: it lays the binning buffer out for capture_margin % (default 150) of the history of earlier EAGER forwards of
// the same configuration and enqueues every kernel exactly once; the lists of a replay whose num_rendered outgrows that
// capacity stay clamped (a truncated frame, never an out-of-bounds access) and tile_scan leaves the count in a sticky
// host word that gsr_capture_overflow() reports, so the caller can re-capture.  The reference cannot be captured at all
// (blocking cudaMemcpy in the middle of its forward, S/cuda_rasterizer/rasterizer_impl.cu:282).
static bool stream_capturing(cudaStream_t s) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return st == cudaStreamCaptureStatusActive;
}

static uint32_t predicted_capacity(int dev, int family, int W, int H, bool capturing = false) {
    if (g_force_cap > 0) return (uint32_t)g_force_cap;
    std::lock_guard<std::mutex> lk(g_hist_mu);
    for (int i = 0; i < g_hist_n; i++) {
        RHistEntry& e = g_hist[i];
        if (e.dev == dev && e.family == family && e.W == W && e.H == H) {
            e.stamp = ++g_hist_clock;
            const double c = fmin(2.0e9, (capturing ? 0.01 * g_capture_margin : 1.25) * e.hist + 4096.0);
            return (uint32_t)((((uint64_t)c + 4095) / 4096) * 4096);
        }
    }
    return 0;   // no history: the caller waits for R first
}
static void record_num_rendered(int dev, int family, int W, int H, uint32_t R) {
    std::lock_guard<std::mutex> lk(g_hist_mu);
    int slot = -1;
    for (int i = 0; i < g_hist_n; i++)
        if (g_hist[i].dev == dev && g_hist[i].family == family && g_hist[i].W == W && g_hist[i].H == H) slot = i;
    if (slot < 0) {
        if (g_hist_n < 32) slot = g_hist_n++;
        else { slot = 0; for (int i = 1; i < 32; i++) if (g_hist[i].stamp < g_hist[slot].stamp) slot = i; }   // least recently used
        g_hist[slot] = {dev, family, W, H, 0.0, 0};
    }
    // slowly decaying maximum: alternating near / far cameras keep the larger layout, a shrinking scene lets go of it
    g_hist[slot].hist = fmax((double)R, 0.97 * g_hist[slot].hist);
    g_hist[slot].stamp = ++g_hist_clock;
}

// pinned, mapped, portable host words: slot = {sequence number, R, prefiltered flag, sticky overflow}; TWO slots per host
// thread -- the one eager forwards wait on, and one that only forwards recorded into CUDA graphs write (a graph replaying
// on another stream must not overwrite the sequence number an eager forward of the same thread is waiting for)
static uint32_t* g_slots = nullptr;
static std::atomic<int> g_slot_next{0};
static std::mutex g_slot_mu;
static volatile uint32_t* my_host_slot(bool capture = false) {
    thread_local volatile uint32_t* mine = nullptr;
    if (!mine) {
        std::lock_guard<std::mutex> lk(g_slot_mu);
        if (!g_slots) {
            if (cudaHostAlloc((void**)&g_slots, 256 * 32, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) return nullptr;
            memset(g_slots, 0, 256 * 32);
        }
        mine = g_slots + 8 * (g_slot_next.fetch_add(1) % 256);
    }
    return mine + (capture ? 4 : 0);
}
static thread_local uint32_t g_seq = 0;
static const char* const kNoHistory =
    "%s: the stream is being captured into a CUDA graph but this (device, rasterizer, resolution) has no num_rendered "
    "history -- run at least one eager forward of the same configuration before capturing";

// Wait until tile_scan has published sequence number `seq`; rb = {R, prefiltered flag}.  Pure host-memory polling;
// the stream is queried now and then so that a faulted kernel turns into an error instead of a hang.
static int wait_num_rendered(volatile uint32_t* slot, uint32_t seq, cudaStream_t s, uint32_t rb[2]) {
    for (uint64_t spins = 1;; spins++) {
        if (slot[0] == seq) break;
        if ((spins & 0x3fff) == 0) {
            const cudaError_t e = cudaStreamQuery(s);
            if (e != cudaSuccess && e != cudaErrorNotReady) { set_error("forward kernels failed: %s", cudaGetErrorString(e)); return GSR_E_CUDA; }
            if (e == cudaSuccess && slot[0] != seq) {
                std::atomic_thread_fence(std::memory_order_seq_cst);
                if (slot[0] != seq) { set_error("tile_scan finished without publishing num_rendered"); return GSR_E_CUDA; }
            }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    rb[0] = slot[1]; rb[1] = slot[2];
    return GSR_OK;
}

// ---- per-kernel device timing (cudaEvents on the caller's stream, off by default) ----
// bench.py needs the dominant kernel's duration measured live, outside any profiler.
struct Prof {
    bool created = false;
    cudaEvent_t ev[GSR_PROF_SLOTS][2];
    bool used[GSR_PROF_SLOTS];
};
constexpr int PROF_MAX_DEV = 64;
static Prof g_prof[PROF_MAX_DEV];          // one event set per device: events are created on, and only recorded into streams of, that device
static std::atomic<bool> g_prof_on{false};
static inline Prof* prof_here() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PROF_MAX_DEV || !g_prof[dev].created) return nullptr;
    return &g_prof[dev];
}
static inline void prof_begin(int slot, cudaStream_t s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    if (Prof* p = prof_here()) { if (cudaEventRecord(p->ev[slot][0], s) != cudaSuccess) (void)cudaGetLastError(); }
}
static inline void prof_end(int slot, cudaStream_t s) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    if (Prof* p = prof_here()) {
        if (cudaEventRecord(p->ev[slot][1], s) == cudaSuccess) p->used[slot] = true;
        else (void)cudaGetLastError();
    }
}

}  // namespace gsr

using namespace gsr;

extern "C" {

int gsr_abi_version(void) { return GSR_ABI_VERSION; }
const char* gsr_last_error(void) { return g_err; }
const char* gsr_build_arch(void) { return "sm_100a"; }

int gsr_set_option(const char* name, int value) {
    if (!name) return GSR_E_INVALID;
    if (!strcmp(name, "no_cull")) { g_no_cull = value ? 1 : 0; return GSR_OK; }
    if (!strcmp(name, "dbg")) { g_dbg = value; return GSR_OK; }
    if (!strcmp(name, "no_used_bits")) { g_no_used_bits = value ? 1 : 0; return GSR_OK; }
    if (!strcmp(name, "force_capacity")) { g_force_cap = value > 0 ? value : 0; return GSR_OK; }
    if (!strcmp(name, "capture_margin")) { g_capture_margin = value >= 100 ? value : 100; return GSR_OK; }
    set_error("gsr_set_option: unknown option %s", name);
    return GSR_E_INVALID;
}

int gsr_last_num_rendered(void) { return g_true_R; }

unsigned int gsr_capture_overflow(int reset) {
    volatile uint32_t* slot = my_host_slot(true);
    if (!slot) return 0u;
    std::atomic_thread_fence(std::memory_order_acquire);
    const uint32_t v = slot[3];
    if (reset) slot[3] = 0u;
    return v;
}

int gsr_profile_enable(int on) {
    // applies to the CURRENT device (events are per device); the on/off switch is global
    int dev = 0;
    GSR_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= PROF_MAX_DEV) { set_error("gsr_profile_enable: device %d out of range", dev); return GSR_E_INVALID; }
    Prof& p = g_prof[dev];
    if (on && !p.created) {
        for (int i = 0; i < GSR_PROF_SLOTS; i++)
            for (int j = 0; j < 2; j++) GSR_CUDA_CHECK(cudaEventCreate(&p.ev[i][j]));
        p.created = true;
    }
    for (int i = 0; i < GSR_PROF_SLOTS; i++) p.used[i] = false;
    g_prof_on.store(on != 0);
    return GSR_OK;
}

int gsr_profile_read(float* ms_host) {
    Prof* p = prof_here();
    if (!ms_host || !p) { set_error("gsr_profile_read: profiling was never enabled on the current device"); return GSR_E_INVALID; }
    for (int i = 0; i < GSR_PROF_SLOTS; i++) {
        ms_host[i] = -1.0f;
        if (!p->used[i]) continue;
        GSR_CUDA_CHECK(cudaEventSynchronize(p->ev[i][1]));
        GSR_CUDA_CHECK(cudaEventElapsedTime(&ms_host[i], p->ev[i][0], p->ev[i][1]));
    }
    return GSR_OK;
}

int gsr_surfel_forward(gsr_buffer_fn geometryBuffer, gsr_buffer_fn binningBuffer, gsr_buffer_fn imageBuffer,
                       void* user, int P, int D, int M, const float* background, int width, int height,
                       const float* means3D, const float* shs, const float* colors_precomp,
                       const float* opacities, const float* scales, float scale_modifier,
                       const float* rotations, const float* transMat_precomp, const float* viewmatrix,
                       const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                       int prefiltered, float* out_color, float* out_others, int* radii, int debug,
                       void* stream_v) {
    (void)tan_fovx; (void)tan_fovy;
    cudaStream_t s = (cudaStream_t)stream_v;
    if (P < 0 || width <= 0 || height <= 0 || !out_color || !out_others || !background || !viewmatrix ||
        !projmatrix) { set_error("gsr_surfel_forward: invalid argument"); return GSR_E_INVALID; }
    if (P > 0 && (!means3D || !opacities || !radii)) { set_error("gsr_surfel_forward: means3D/opacities/radii required"); return GSR_E_INVALID; }
    if (P > 0 && ((shs == nullptr) == (colors_precomp == nullptr))) { set_error("provide exactly one of shs / colors_precomp"); return GSR_E_INVALID; }
    if (P > 0 && (((scales == nullptr) || (rotations == nullptr)) == (transMat_precomp == nullptr))) { set_error("provide exactly one of scales+rotations / transMat_precomp"); return GSR_E_INVALID; }
    if (P > 0 && shs && !cam_pos) { set_error("cam_pos required with shs"); return GSR_E_INVALID; }
    if (P > 0 && shs && (M < 1 || M > 16)) { set_error("shs: M = %d coefficients per channel, supported 1..16 (degree <= 3)", M); return GSR_E_INVALID; }
    const int W = width, H = height;
    const size_t N = (size_t)W * H;

    HostTimer ht;
    const ViewParams vc = make_view(viewmatrix, projmatrix, cam_pos, W, H, scale_modifier);
    const int ntiles = vc.gx * vc.gy;

    ImageWs iw;
    size_t ibytes = ImageWs::carve(iw, nullptr, W, H);
    char* ibase = imageBuffer(user, ibytes);
    if (!ibase) { set_error("imageBuffer callback failed (%zu bytes)", ibytes); return GSR_E_ALLOC; }
    ImageWs::carve(iw, align256(ibase), W, H);
    // tile_count, tile_offset and tile_cursor are adjacent: one memset clears all three
    GSR_CUDA_CHECK(cudaMemsetAsync(iw.tile_count, 0, (size_t)((char*)(iw.total + 8) - (char*)iw.tile_count), s));
    ht.mark("imgbuf");

    int R = 0;
    GeomWs gw;
    BinWs bw;
    memset(&bw, 0, sizeof(bw));
    memset(&gw, 0, sizeof(gw));
    if (P > 0) {
        size_t gbytes = GeomWs::carve(gw, nullptr, P);
        char* gbase = geometryBuffer(user, gbytes);
        if (!gbase) { set_error("geometryBuffer callback failed (%zu bytes)", gbytes); return GSR_E_ALLOC; }
        GeomWs::carve(gw, align256(gbase), P);
        ht.mark("geombuf");
        GSR_CUDA_CHECK(cudaMemsetAsync(gw.flags, 0, 32 * sizeof(int), s));

        prof_begin(GSR_PROF_PREPROCESS_FWD, s);
        surfel_preprocess_fwd<<<(P + 255) / 256, 256, 0, s>>>(
            P, D, M, means3D, (const float2*)scales, (const float4*)rotations, opacities, nullptr, transMat_precomp,
            true, vc, prefiltered != 0, no_cull(), radii, gw.geom, gw.cull, gw.depths, gw.masks,
            iw.tile_count, gw.rgb, gw.clamped, gw.flags);
        // SH -> RGB runs as its own warp-cooperative, coalesced kernel (sh.cu) on the visible Gaussians
        if (shs) GSR_CUDA_CHECK(launch_sh_forward(P, D, M, means3D, cam_pos, shs, radii, gw.rgb, gw.clamped, s));
        prof_end(GSR_PROF_PREPROCESS_FWD, s);
        GSR_CUDA_CHECK(cudaGetLastError());
        ht.mark("launch_pre");
    }

    // binning + render for a buffer laid out for `cap` entries (see "num_rendered without blocking the stream")
    auto lay_out = [&](uint32_t cap) -> int {
        const size_t bbytes = BinWs::carve(bw, nullptr, (int64_t)cap, P);
        char* bbase = binningBuffer(user, bbytes);
        if (!bbase) { set_error("binningBuffer callback failed (%zu bytes)", bbytes); return GSR_E_ALLOC; }
        BinWs::carve(bw, align256(bbase), (int64_t)cap, P);
        return GSR_OK;
    };
    const bool capturing = stream_capturing(s);
    auto scan = [&](uint32_t cap, volatile uint32_t* slot, uint32_t seq) -> int {
        prof_begin(GSR_PROF_SCAN, s);
        tile_scan<<<1, 1024, 0, s>>>(ntiles, iw.tile_count, iw.tile_offset, iw.tile_cursor, iw.total, gw.flags, cap, slot, seq,
                                     capturing ? 1 : 0);
        prof_end(GSR_PROF_SCAN, s);
        GSR_CUDA_CHECK(cudaGetLastError());
        return GSR_OK;
    };
    auto bin_and_render = [&](uint32_t cap) -> int {
        if (P > 0 && cap > 0) {
            prof_begin(GSR_PROF_DUPLICATE, s);
            scatter_keys<<<(P + 255) / 256, 256, 0, s>>>(P, &gw.geom->tu.w, &gw.geom->tv.w, (int)(sizeof(GeomRec) / 4), gw.cull,
                                                         gw.depths, radii, gw.masks, vc.gx, vc.gy, iw.tile_cursor, bw.keys, cap);
            prof_end(GSR_PROF_DUPLICATE, s);
            GSR_CUDA_CHECK(cudaGetLastError());
            const float* colors = colors_precomp ? colors_precomp : gw.rgb;
            prof_begin(GSR_PROF_BUILD_RECORDS, s);
            sort_build_records<<<ntiles, 256, 0, s>>>(iw.tile_offset, bw.keys, gw.geom, colors, vc.gx, W, H,
                                                      bw.planes, bw.plane_stride, g_dbg);
            prof_end(GSR_PROF_BUILD_RECORDS, s);
            GSR_CUDA_CHECK(cudaGetLastError());
        }
        prof_begin(GSR_PROF_RENDER_FWD, s);
        if (cap > 0 && P < (1 << REC_USED_SHIFT) && !g_no_used_bits)
            surfel_render_fwd<true><<<ntiles * fwd_ctas_per_tile(), 256 / fwd_ctas_per_tile(), 0, s>>>(
                iw.tile_offset, bw.planes, bw.plane_stride, W, H, vc.gx, background, iw.final_T, iw.n_contrib, out_color, out_others,
                bw.planes + 3 * bw.plane_stride);
        else
            surfel_render_fwd<false><<<ntiles * fwd_ctas_per_tile(), 256 / fwd_ctas_per_tile(), 0, s>>>(
                iw.tile_offset, bw.planes, bw.plane_stride, W, H, vc.gx, background, iw.final_T, iw.n_contrib, out_color, out_others,
                nullptr);
        prof_end(GSR_PROF_RENDER_FWD, s);
        GSR_CUDA_CHECK(cudaGetLastError());
        return GSR_OK;
    };

    uint32_t layout = 0;     // entries the binning buffer is laid out for == the value handed back to the caller
    if (P > 0) {
        int dev = 0;
        GSR_CUDA_CHECK(cudaGetDevice(&dev));
        volatile uint32_t* slot = my_host_slot(capturing);
        if (!slot) { set_error("cudaHostAlloc of the num_rendered slot failed"); return GSR_E_CUDA; }
        const uint32_t seq = ++g_seq;
        const uint32_t cap = predicted_capacity(dev, 0, W, H, capturing);
        int rc;
        if (capturing) {
            if (cap == 0) { set_error(kNoHistory, "gsr_surfel_forward"); return GSR_E_INVALID; }
            if ((rc = lay_out(cap)) < 0) return rc;
            if ((rc = scan(cap, slot, seq)) < 0) return rc;
            if ((rc = bin_and_render(cap)) < 0) return rc;
            return (int)cap;      // recorded once; nothing to wait for (see "CUDA-graph capture")
        }
        if (cap > 0) {
            if ((rc = lay_out(cap)) < 0) return rc;
            if ((rc = scan(cap, slot, seq)) < 0) return rc;
            if ((rc = bin_and_render(cap)) < 0) return rc;
            ht.mark("launch_spec");
        } else {
            if ((rc = scan(0xffffffffu, slot, seq)) < 0) return rc;
        }
        uint32_t rb[2] = {0u, 0u};
        if ((rc = wait_num_rendered(slot, seq, s, rb)) < 0) return rc;
        ht.mark("wait_R");
        if (rb[1]) { set_error("Point is filtered although prefiltered is set. This shouldn't happen!"); return GSR_E_PREFILTERED; }
        if (rb[0] > 0x7fffff00u) { set_error("num_rendered overflow (%u)", rb[0]); return GSR_E_OVERFLOW; }
        g_true_R = (int)rb[0];
        if (g_force_cap <= 0) record_num_rendered(dev, 0, W, H, rb[0]);
        layout = cap;
        if (cap == 0 || rb[0] > cap) {
            // first frame of this configuration, or more pairs than predicted: exact layout, re-enqueue behind the clamped run
            layout = rb[0];
            if ((rc = lay_out(layout)) < 0) return rc;
            if (cap > 0 && (rc = scan(0xffffffffu, nullptr, 0)) < 0) return rc;
            if ((rc = bin_and_render(layout)) < 0) return rc;
        }
    } else {
        int rc;
        g_true_R = 0;
        if ((rc = lay_out(0)) < 0) return rc;
        if ((rc = bin_and_render(0)) < 0) return rc;
    }
    R = (int)layout;
    (void)N;
    ht.mark("launch_rest");
    ht.flush("forward");
    if (debug && !capturing) GSR_CUDA_CHECK(cudaStreamSynchronize(s));
    return R;
}

int gsr_surfel_backward(int P, int D, int M, int R, const float* background, int width, int height,
                        const float* means3D, const float* shs, const float* colors_precomp,
                        const float* scales, float scale_modifier, const float* rotations,
                        const float* transMat_precomp, const float* viewmatrix, const float* projmatrix,
                        const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                        char* geom_buffer, char* binning_buffer, char* image_buffer, const float* dL_dpix,
                        const float* dL_dothers, float* dL_dmean2D, float* dL_dnormal, float* dL_dopacity,
                        float* dL_dcolor, float* dL_dmean3D, float* dL_dtransMat, float* dL_dsh,
                        float* dL_dscale, float* dL_drot, int debug, void* stream_v) {
    (void)colors_precomp;
    cudaStream_t s = (cudaStream_t)stream_v;
    if (P == 0) return GSR_OK;
    if (P < 0 || R < 0 || !geom_buffer || !binning_buffer || !image_buffer || !dL_dpix || !dL_dothers ||
        !dL_dmean2D || !dL_dnormal || !dL_dopacity || !dL_dcolor || !dL_dmean3D || !dL_dtransMat || !radii ||
        !means3D) { set_error("gsr_surfel_backward: invalid argument"); return GSR_E_INVALID; }
    if (M > 0 && shs && !dL_dsh) { set_error("dL_dsh required with shs"); return GSR_E_INVALID; }
    const int W = width, H = height;
    // S/backward.cu:614-615: W,H recomputed from float32 focal * tan * 2 and truncated (quirk Q3)
    const float focal_y = (float)H / (2.0f * tan_fovy), focal_x = (float)W / (2.0f * tan_fovx);
    const int Wb = (int)(focal_x * tan_fovx * 2), Hb = (int)(focal_y * tan_fovy * 2);
    const ViewParams vc = make_view(viewmatrix, projmatrix, campos, W, H, scale_modifier);
    const int ntiles = vc.gx * vc.gy;

    GeomWs gw; ImageWs iw; BinWs bw;
    GeomWs::carve(gw, align256(geom_buffer), P);
    ImageWs::carve(iw, align256(image_buffer), W, H);
    BinWs::carve(bw, align256(binning_buffer), R, P);

    GSR_CUDA_CHECK(cudaMemsetAsync(bw.gacc, 0, (size_t)P * GACC_STRIDE * sizeof(float), s));
    if (R > 0) {
        prof_begin(GSR_PROF_RENDER_BWD, s);
        GSR_CUDA_CHECK(launch_surfel_render_bwd(P < (1 << REC_USED_SHIFT) && !g_no_used_bits, ntiles, iw.tile_offset, bw.planes,
                                                bw.plane_stride, W, H, vc.gx, background, iw.final_T, iw.n_contrib, dL_dpix,
                                                dL_dothers, bw.gacc, (g_dbg & 2) != 0, s));
        prof_end(GSR_PROF_RENDER_BWD, s);
        GSR_CUDA_CHECK(cudaGetLastError());
    }
    const bool precomp = (scales == nullptr);
    prof_begin(GSR_PROF_PREPROCESS_BWD, s);
    surfel_preprocess_bwd<<<(P + 255) / 256, 256, 0, s>>>(
        P, D, M, means3D, (const float2*)scales, (const float4*)rotations, nullptr, precomp, vc, Wb, Hb, radii,
        gw.geom, gw.clamped, bw.gacc, dL_dmean2D, dL_dnormal, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dtransMat,
        nullptr, dL_dscale, dL_drot);
    // dL/dsh (fully written) and the view-direction term of dL/dmean3D: coalesced SH kernel (sh.cu)
    if (shs && M > 0) GSR_CUDA_CHECK(launch_sh_backward(P, D, M, means3D, campos, shs, gw.clamped, radii, dL_dcolor, dL_dsh, dL_dmean3D, s));
    prof_end(GSR_PROF_PREPROCESS_BWD, s);
    GSR_CUDA_CHECK(cudaGetLastError());
    if (debug && !stream_capturing(s)) GSR_CUDA_CHECK(cudaStreamSynchronize(s));
    return GSR_OK;
}

int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream_v) {
    (void)projmatrix;
    cudaStream_t s = (cudaStream_t)stream_v;
    if (P == 0) return GSR_OK;
    if (P < 0 || !means3D || !viewmatrix || !present) { set_error("gsr_mark_visible: invalid argument"); return GSR_E_INVALID; }
    const ViewParams vc = make_view(viewmatrix, projmatrix, nullptr, 0, 0, 1.0f);
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, s>>>(P, means3D, vc, present);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}

}  // extern "C"

// ---- EWA family (3DGS + PGSR plane): shared host orchestration ------------------------------------
namespace gsr {

struct EwaFwdArgs {
    gsr_buffer_fn geometryBuffer, binningBuffer, imageBuffer;
    void* user;
    int P, D, M;
    const float* background;
    int W, H;
    const float *means3D, *shs, *colors_precomp, *opacities, *scales;
    float scale_modifier;
    const float *rotations, *cov3D_precomp, *all_map, *viewmatrix, *projmatrix, *cam_pos;
    float tan_fovx, tan_fovy;
    int prefiltered;
    float* out_color;
    int* radii;
    int* out_observe;         // plane only (else NULL)
    float *out_all_map, *out_plane_depth;
    bool plane, geo;
    int debug;
    cudaStream_t s;
};

static int ewa_forward(const EwaFwdArgs& a, const char* who) {
    cudaStream_t s = a.s;
    const int P = a.P, W = a.W, H = a.H;
    if (P < 0 || W <= 0 || H <= 0 || !a.out_color || !a.background || !a.viewmatrix || !a.projmatrix) {
        set_error("%s: invalid argument", who); return GSR_E_INVALID;
    }
    if (P > 0 && (!a.means3D || !a.opacities || !a.radii)) { set_error("%s: means3D/opacities/radii required", who); return GSR_E_INVALID; }
    if (P > 0 && ((a.shs == nullptr) == (a.colors_precomp == nullptr))) { set_error("provide exactly one of shs / colors_precomp"); return GSR_E_INVALID; }
    if (P > 0 && (((a.scales == nullptr) || (a.rotations == nullptr)) == (a.cov3D_precomp == nullptr))) { set_error("provide exactly one of scales+rotations / cov3D_precomp"); return GSR_E_INVALID; }
    if (P > 0 && a.shs && !a.cam_pos) { set_error("cam_pos required with shs"); return GSR_E_INVALID; }
    if (P > 0 && a.shs && (a.M < 1 || a.M > 16)) { set_error("shs: M = %d coefficients per channel, supported 1..16 (degree <= 3)", a.M); return GSR_E_INVALID; }
    if (a.plane && (!a.out_observe || (a.geo && (!a.out_all_map || !a.out_plane_depth || (P > 0 && !a.all_map))))) {
        set_error("%s: out_observe / out_all_map / out_plane_depth / all_map required", who); return GSR_E_INVALID;
    }
    const size_t N = (size_t)W * H;
    const ViewParams vc = make_view(a.viewmatrix, a.projmatrix, a.cam_pos, W, H, a.scale_modifier);
    const int ntiles = vc.gx * vc.gy;
    const float focal_y = (float)H / (2.0f * a.tan_fovy), focal_x = (float)W / (2.0f * a.tan_fovx);   // G/rasterizer_impl.cu:223-224
    const int nplanes = a.geo ? EWA_PLANES_GEO : EWA_PLANES;

    ImageWs iw;
    size_t ibytes = ImageWs::carve(iw, nullptr, W, H, 1, 1);
    char* ibase = a.imageBuffer(a.user, ibytes);
    if (!ibase) { set_error("imageBuffer callback failed (%zu bytes)", ibytes); return GSR_E_ALLOC; }
    ImageWs::carve(iw, align256(ibase), W, H, 1, 1);
    GSR_CUDA_CHECK(cudaMemsetAsync(iw.tile_count, 0, (size_t)((char*)(iw.total + 8) - (char*)iw.tile_count), s));
    if (a.plane && P > 0) GSR_CUDA_CHECK(cudaMemsetAsync(a.out_observe, 0, (size_t)P * sizeof(int), s));
    if (a.plane && !a.geo) {   // torch::full(..., 0) in the reference glue (L/rasterize_points.cu:75-76)
        if (a.out_all_map) GSR_CUDA_CHECK(cudaMemsetAsync(a.out_all_map, 0, NUM_ALL_MAP * N * sizeof(float), s));
        if (a.out_plane_depth) GSR_CUDA_CHECK(cudaMemsetAsync(a.out_plane_depth, 0, N * sizeof(float), s));
    }

    int R = 0;
    EwaGeomWs gw;
    BinWs bw;
    memset(&bw, 0, sizeof(bw));
    memset(&gw, 0, sizeof(gw));
    if (P > 0) {
        size_t gbytes = EwaGeomWs::carve(gw, nullptr, P);
        char* gbase = a.geometryBuffer(a.user, gbytes);
        if (!gbase) { set_error("geometryBuffer callback failed (%zu bytes)", gbytes); return GSR_E_ALLOC; }
        EwaGeomWs::carve(gw, align256(gbase), P);
        GSR_CUDA_CHECK(cudaMemsetAsync(gw.flags, 0, 32 * sizeof(int), s));
        prof_begin(GSR_PROF_PREPROCESS_FWD, s);
        ewa_preprocess_fwd<false><<<(P + 255) / 256, 256, 0, s>>>(
            P, a.D, a.M, a.means3D, a.scales, (const float4*)a.rotations, a.opacities, nullptr, a.cov3D_precomp,
            true, vc, focal_x, focal_y, a.tan_fovx, a.tan_fovy, a.prefiltered != 0, no_cull(),
            a.radii, gw.geom, gw.cull, gw.depths, gw.masks, iw.tile_count, gw.rgb, gw.clamped, gw.flags);
        if (a.shs) GSR_CUDA_CHECK(launch_sh_forward(P, a.D, a.M, a.means3D, a.cam_pos, a.shs, a.radii, gw.rgb, gw.clamped, s));
        prof_end(GSR_PROF_PREPROCESS_FWD, s);
        GSR_CUDA_CHECK(cudaGetLastError());
    }

    // binning + render for a buffer laid out for `cap` entries (see "num_rendered without blocking the stream")
    auto lay_out = [&](uint32_t cap) -> int {
        const size_t bbytes = BinWs::carve(bw, nullptr, (int64_t)cap, P, nplanes, EWA_GACC);
        char* bbase = a.binningBuffer(a.user, bbytes);
        if (!bbase) { set_error("binningBuffer callback failed (%zu bytes)", bbytes); return GSR_E_ALLOC; }
        BinWs::carve(bw, align256(bbase), (int64_t)cap, P, nplanes, EWA_GACC);
        return GSR_OK;
    };
    const bool capturing = stream_capturing(s);
    auto scan = [&](uint32_t cap, volatile uint32_t* slot, uint32_t seq) -> int {
        prof_begin(GSR_PROF_SCAN, s);
        tile_scan<<<1, 1024, 0, s>>>(ntiles, iw.tile_count, iw.tile_offset, iw.tile_cursor, iw.total, gw.flags, cap, slot, seq,
                                     capturing ? 1 : 0);
        prof_end(GSR_PROF_SCAN, s);
        GSR_CUDA_CHECK(cudaGetLastError());
        return GSR_OK;
    };
    auto bin_and_render = [&](uint32_t cap, bool again) -> int {
        if (P > 0 && cap > 0) {
            prof_begin(GSR_PROF_DUPLICATE, s);
            scatter_keys<<<(P + 255) / 256, 256, 0, s>>>(P, &gw.geom->a.x, &gw.geom->a.y, (int)(sizeof(EwaGeom) / 4), gw.cull,
                                                         gw.depths, a.radii, gw.masks, vc.gx, vc.gy, iw.tile_cursor, bw.keys, cap);
            prof_end(GSR_PROF_DUPLICATE, s);
            GSR_CUDA_CHECK(cudaGetLastError());
            const float* colors = a.colors_precomp ? a.colors_precomp : gw.rgb;
            prof_begin(GSR_PROF_BUILD_RECORDS, s);
            if (a.geo) ewa_build_records<true><<<ntiles, 256, 0, s>>>(iw.tile_offset, bw.keys, gw.geom, colors, a.all_map, vc.gx, bw.planes, bw.plane_stride);
            else ewa_build_records<false><<<ntiles, 256, 0, s>>>(iw.tile_offset, bw.keys, gw.geom, colors, nullptr, vc.gx, bw.planes, bw.plane_stride);
            prof_end(GSR_PROF_BUILD_RECORDS, s);
            GSR_CUDA_CHECK(cudaGetLastError());
        }
        // out_observe is accumulated with atomics by the render kernel: a re-run starts from zero again
        if (again && a.plane && P > 0) GSR_CUDA_CHECK(cudaMemsetAsync(a.out_observe, 0, (size_t)P * sizeof(int), s));
        prof_begin(GSR_PROF_RENDER_FWD, s);
        // plane 1 holds the idx|flag word that receives the forward's "blended" marks (P < 2^23)
        float4* mark = (cap > 0 && P < (1 << REC_USED_SHIFT) && !g_no_used_bits) ? bw.planes + 1 * bw.plane_stride : nullptr;
        if (a.geo) ewa_render_fwd<true><<<ntiles, TILE_PIX, 0, s>>>(iw.tile_offset, bw.planes, bw.plane_stride, W, H, vc.gx, a.background, focal_x, focal_y,
                                                                    iw.final_T, iw.n_contrib, a.out_color, a.out_observe, a.out_all_map, a.out_plane_depth, mark);
        else ewa_render_fwd<false><<<ntiles, TILE_PIX, 0, s>>>(iw.tile_offset, bw.planes, bw.plane_stride, W, H, vc.gx, a.background, focal_x, focal_y,
                                                               iw.final_T, iw.n_contrib, a.out_color, a.plane ? a.out_observe : nullptr, nullptr, nullptr, mark);
        prof_end(GSR_PROF_RENDER_FWD, s);
        GSR_CUDA_CHECK(cudaGetLastError());
        return GSR_OK;
    };

    uint32_t layout = 0;
    if (P > 0) {
        int dev = 0;
        GSR_CUDA_CHECK(cudaGetDevice(&dev));
        volatile uint32_t* slot = my_host_slot(capturing);
        if (!slot) { set_error("cudaHostAlloc of the num_rendered slot failed"); return GSR_E_CUDA; }
        const uint32_t seq = ++g_seq;
        const int family = a.geo ? 3 : (a.plane ? 2 : 1);
        const uint32_t cap = predicted_capacity(dev, family, W, H, capturing);
        int rc;
        if (capturing) {
            if (cap == 0) { set_error(kNoHistory, who); return GSR_E_INVALID; }
            if ((rc = lay_out(cap)) < 0) return rc;
            if ((rc = scan(cap, slot, seq)) < 0) return rc;
            if ((rc = bin_and_render(cap, false)) < 0) return rc;
            return (int)cap;
        }
        if (cap > 0) {
            if ((rc = lay_out(cap)) < 0) return rc;
            if ((rc = scan(cap, slot, seq)) < 0) return rc;
            if ((rc = bin_and_render(cap, false)) < 0) return rc;
        } else {
            if ((rc = scan(0xffffffffu, slot, seq)) < 0) return rc;
        }
        uint32_t rb[2] = {0u, 0u};
        if ((rc = wait_num_rendered(slot, seq, s, rb)) < 0) return rc;
        if (rb[1]) { set_error("Point is filtered although prefiltered is set. This shouldn't happen!"); return GSR_E_PREFILTERED; }
        if (rb[0] > 0x7fffff00u) { set_error("num_rendered overflow (%u)", rb[0]); return GSR_E_OVERFLOW; }
        g_true_R = (int)rb[0];
        if (g_force_cap <= 0) record_num_rendered(dev, family, W, H, rb[0]);
        layout = cap;
        if (cap == 0 || rb[0] > cap) {
            layout = rb[0];
            if ((rc = lay_out(layout)) < 0) return rc;
            if (cap > 0 && (rc = scan(0xffffffffu, nullptr, 0)) < 0) return rc;
            if ((rc = bin_and_render(layout, cap > 0)) < 0) return rc;
        }
    } else {
        int rc;
        g_true_R = 0;
        if ((rc = lay_out(0)) < 0) return rc;
        if ((rc = bin_and_render(0, false)) < 0) return rc;
    }
    R = (int)layout;
    if (a.debug && !capturing) GSR_CUDA_CHECK(cudaStreamSynchronize(s));
    return R;
}

struct EwaBwdArgs {
    int P, D, M, R;
    const float* background;
    const float* all_map_pixels;
    int W, H;
    const float *means3D, *shs, *all_maps, *scales;
    float scale_modifier;
    const float *rotations, *cov3D_precomp, *viewmatrix, *projmatrix, *campos;
    float tan_fovx, tan_fovy;
    const int* radii;
    char *geom_buffer, *binning_buffer, *image_buffer;
    const float *dL_dpix, *dL_dout_all_map, *dL_dout_plane_depth;
    float *dL_dmean2D, *dL_dmean2D_abs, *dL_dconic, *dL_dopacity, *dL_dcolor, *dL_dmean3D, *dL_dcov3D, *dL_dsh,
        *dL_dscale, *dL_drot, *dL_dall_map;
    bool plane, geo;
    int debug;
    cudaStream_t s;
};

static int ewa_backward(const EwaBwdArgs& a, const char* who) {
    cudaStream_t s = a.s;
    const int P = a.P, W = a.W, H = a.H, R = a.R;
    if (P == 0) return GSR_OK;
    if (P < 0 || R < 0 || !a.geom_buffer || !a.binning_buffer || !a.image_buffer || !a.dL_dpix || !a.dL_dmean2D ||
        !a.dL_dopacity || !a.dL_dcolor || !a.dL_dmean3D || !a.dL_dcov3D || !a.radii || !a.means3D || !a.background) {
        set_error("%s: invalid argument", who); return GSR_E_INVALID;
    }
    if (a.M > 0 && a.shs && !a.dL_dsh) { set_error("dL_dsh required with shs"); return GSR_E_INVALID; }
    if (a.plane && (!a.dL_dmean2D_abs || !a.dL_dall_map)) { set_error("%s: dL_dmean2D_abs / dL_dall_map required", who); return GSR_E_INVALID; }
    if (a.geo && (!a.all_map_pixels || !a.dL_dout_all_map || !a.dL_dout_plane_depth || !a.all_maps)) {
        set_error("%s: all_map_pixels / dL_dout_all_map / dL_dout_plane_depth / all_maps required with render_geo", who); return GSR_E_INVALID;
    }
    const ViewParams vc = make_view(a.viewmatrix, a.projmatrix, a.campos, W, H, a.scale_modifier);
    const int ntiles = vc.gx * vc.gy;
    const float focal_y = (float)H / (2.0f * a.tan_fovy), focal_x = (float)W / (2.0f * a.tan_fovx);   // G/rasterizer_impl.cu:369-370
    const int nplanes = a.geo ? EWA_PLANES_GEO : EWA_PLANES;
    EwaGeomWs gw; ImageWs iw; BinWs bw;
    EwaGeomWs::carve(gw, align256(a.geom_buffer), P);
    ImageWs::carve(iw, align256(a.image_buffer), W, H, 1, 1);
    BinWs::carve(bw, align256(a.binning_buffer), R, P, nplanes, EWA_GACC);

    GSR_CUDA_CHECK(cudaMemsetAsync(bw.gacc, 0, (size_t)P * EWA_GACC * sizeof(float), s));
    if (R > 0) {
        prof_begin(GSR_PROF_RENDER_BWD, s);
        const bool used = P < (1 << REC_USED_SHIFT) && !g_no_used_bits;
        GSR_CUDA_CHECK(launch_ewa_render_bwd(a.geo ? 2 : (a.plane ? 1 : 0), used, ntiles, iw.tile_offset, bw.planes, bw.plane_stride, W, H,
                                             vc.gx, a.background, focal_x, focal_y, iw.final_T, iw.n_contrib,
                                             a.geo ? a.all_map_pixels : nullptr, a.dL_dpix, a.geo ? a.dL_dout_all_map : nullptr,
                                             a.geo ? a.dL_dout_plane_depth : nullptr, bw.gacc, (g_dbg & 2) != 0, s));
        prof_end(GSR_PROF_RENDER_BWD, s);
        GSR_CUDA_CHECK(cudaGetLastError());
    }
    prof_begin(GSR_PROF_PREPROCESS_BWD, s);
    ewa_preprocess_bwd<<<(P + 255) / 256, 256, 0, s>>>(
        P, a.D, a.M, a.means3D, a.scales, (const float4*)a.rotations, nullptr, a.cov3D_precomp, vc, focal_x, focal_y,
        a.tan_fovx, a.tan_fovy, a.radii, gw.clamped, bw.gacc, a.dL_dmean2D, a.plane ? a.dL_dmean2D_abs : nullptr,
        a.dL_dconic, a.dL_dopacity, a.dL_dcolor, a.dL_dmean3D, a.dL_dcov3D, nullptr,
        a.scales ? a.dL_dscale : nullptr, a.scales ? a.dL_drot : nullptr, a.plane ? a.dL_dall_map : nullptr);
    if (a.shs && a.M > 0) GSR_CUDA_CHECK(launch_sh_backward(P, a.D, a.M, a.means3D, a.campos, a.shs, gw.clamped, a.radii, a.dL_dcolor, a.dL_dsh, a.dL_dmean3D, s));
    prof_end(GSR_PROF_PREPROCESS_BWD, s);
    GSR_CUDA_CHECK(cudaGetLastError());
    if (a.debug && !stream_capturing(s)) GSR_CUDA_CHECK(cudaStreamSynchronize(s));
    return GSR_OK;
}
}  // namespace gsr

extern "C" {

int gsr_gaussian_forward(gsr_buffer_fn geometryBuffer, gsr_buffer_fn binningBuffer, gsr_buffer_fn imageBuffer,
                         void* user, int P, int D, int M, const float* background, int width, int height,
                         const float* means3D, const float* shs, const float* colors_precomp,
                         const float* opacities, const float* scales, float scale_modifier,
                         const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                         const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                         int prefiltered, float* out_color, int* radii, int debug, void* stream) {
    const EwaFwdArgs a = {geometryBuffer, binningBuffer, imageBuffer, user, P, D, M, background, width, height, means3D, shs,
                          colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp, nullptr, viewmatrix,
                          projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered, out_color, radii, nullptr, nullptr, nullptr,
                          false, false, debug, (cudaStream_t)stream};
    return ewa_forward(a, "gsr_gaussian_forward");
}

int gsr_gaussian_backward(int P, int D, int M, int R, const float* background, int width, int height,
                          const float* means3D, const float* shs, const float* colors_precomp,
                          const float* scales, float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                          const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                          char* geom_buffer, char* binning_buffer, char* image_buffer, const float* dL_dpix,
                          float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                          float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                          int debug, void* stream) {
    (void)colors_precomp;
    const EwaBwdArgs a = {P, D, M, R, background, nullptr, width, height, means3D, shs, nullptr, scales, scale_modifier,
                          rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii, geom_buffer,
                          binning_buffer, image_buffer, dL_dpix, nullptr, nullptr, dL_dmean2D, nullptr, dL_dconic,
                          dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, nullptr, false, false,
                          debug, (cudaStream_t)stream};
    return ewa_backward(a, "gsr_gaussian_backward");
}

int gsr_plane_forward(gsr_buffer_fn geometryBuffer, gsr_buffer_fn binningBuffer, gsr_buffer_fn imageBuffer,
                      void* user, int P, int D, int M, const float* background, int width, int height,
                      const float* means3D, const float* shs, const float* colors_precomp,
                      const float* opacities, const float* scales, float scale_modifier,
                      const float* rotations, const float* cov3D_precomp, const float* all_map,
                      const float* viewmatrix, const float* projmatrix, const float* cam_pos, float tan_fovx,
                      float tan_fovy, int prefiltered, float* out_color, int* radii, int* out_observe,
                      float* out_all_map, float* out_plane_depth, int render_geo, int debug, void* stream) {
    const EwaFwdArgs a = {geometryBuffer, binningBuffer, imageBuffer, user, P, D, M, background, width, height, means3D, shs,
                          colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp, all_map, viewmatrix,
                          projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered, out_color, radii, out_observe, out_all_map,
                          out_plane_depth, true, render_geo != 0, debug, (cudaStream_t)stream};
    return ewa_forward(a, "gsr_plane_forward");
}

int gsr_plane_backward(int P, int D, int M, int R, const float* background, const float* all_map_pixels, int width,
                       int height, const float* means3D, const float* shs, const float* colors_precomp,
                       const float* all_maps, const float* scales, float scale_modifier, const float* rotations,
                       const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                       const float* campos, float tan_fovx, float tan_fovy, const int* radii, char* geom_buffer,
                       char* binning_buffer, char* image_buffer, const float* dL_dpix, const float* dL_dout_all_map,
                       const float* dL_dout_plane_depth, float* dL_dmean2D, float* dL_dmean2D_abs, float* dL_dconic,
                       float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                       float* dL_dscale, float* dL_drot, float* dL_dall_map, int render_geo, int debug, void* stream) {
    (void)colors_precomp;
    const EwaBwdArgs a = {P, D, M, R, background, all_map_pixels, width, height, means3D, shs, all_maps, scales,
                          scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy,
                          radii, geom_buffer, binning_buffer, image_buffer, dL_dpix, dL_dout_all_map, dL_dout_plane_depth,
                          dL_dmean2D, dL_dmean2D_abs, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh,
                          dL_dscale, dL_drot, dL_dall_map, true, render_geo != 0, debug, (cudaStream_t)stream};
    return ewa_backward(a, "gsr_plane_backward");
}

int gsr_visible_filter(int P, int width, int height, const float* means3D, const float* scales, float scale_modifier,
                       const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                       const float* projmatrix, float tan_fovx, float tan_fovy, int prefiltered, int* radii, int debug,
                       void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P == 0) return GSR_OK;
    if (P < 0 || width <= 0 || height <= 0 || !means3D || !viewmatrix || !projmatrix || !radii) {
        set_error("gsr_visible_filter: invalid argument"); return GSR_E_INVALID;
    }
    if (((scales == nullptr) || (rotations == nullptr)) == (cov3D_precomp == nullptr)) {
        set_error("provide exactly one of scales+rotations / cov3D_precomp"); return GSR_E_INVALID;
    }
    const ViewParams vc = make_view(viewmatrix, projmatrix, nullptr, width, height, scale_modifier);
    const float focal_y = (float)height / (2.0f * tan_fovy), focal_x = (float)width / (2.0f * tan_fovx);   // F/rasterizer_impl.cu:358-359
    // `prefiltered` is accepted for signature parity and IGNORED: reporting a violation needs a flag word and a read-back,
    // which this asynchronous entry point does not do (the reference __trap()s inside the kernel, F/auxiliary.h:204-208)
    (void)prefiltered;
    ewa_preprocess_fwd<true><<<(P + 255) / 256, 256, 0, s>>>(
        P, 0, 0, means3D, scales, (const float4*)rotations, nullptr, nullptr, cov3D_precomp, true, vc, focal_x, focal_y,
        tan_fovx, tan_fovy, false, true, radii, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    GSR_CUDA_CHECK(cudaGetLastError());
    if (debug && !stream_capturing(s)) GSR_CUDA_CHECK(cudaStreamSynchronize(s));
    return GSR_OK;
}

int gsr_surfel_audit(int P, int R, int width, int height, char* binning_buffer, char* image_buffer, float* margins,
                     int* info, int* mismatches, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0 || R < 0 || width <= 0 || height <= 0 || !binning_buffer || !image_buffer || !margins || !info || !mismatches) {
        set_error("gsr_surfel_audit: invalid argument"); return GSR_E_INVALID;
    }
    const ViewParams vc = make_view(nullptr, nullptr, nullptr, width, height, 1.0f);
    ImageWs iw; BinWs bw;
    ImageWs::carve(iw, align256(image_buffer), width, height);
    BinWs::carve(bw, align256(binning_buffer), R, P);
    const uint32_t idx_mask = (P < (1 << REC_USED_SHIFT) && !g_no_used_bits) ? REC_INDEX_MASK_USED : ~REC_FLAG_ALWAYS;
    GSR_CUDA_CHECK(cudaMemsetAsync(mismatches, 0, sizeof(int), s));
    GSR_CUDA_CHECK(launch_surfel_audit(vc.gx * vc.gy, iw.tile_offset, bw.planes, bw.plane_stride, width, height, vc.gx, iw.final_T,
                                       iw.n_contrib, idx_mask, margins, info, mismatches, s));
    return GSR_OK;
}

int gsr_ewa_audit(int P, int R, int width, int height, int render_geo, char* binning_buffer, char* image_buffer,
                  float* margins, int* info, int* mismatches, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (P <= 0 || R < 0 || width <= 0 || height <= 0 || !binning_buffer || !image_buffer || !margins || !info || !mismatches) {
        set_error("gsr_ewa_audit: invalid argument"); return GSR_E_INVALID;
    }
    const ViewParams vc = make_view(nullptr, nullptr, nullptr, width, height, 1.0f);
    ImageWs iw; BinWs bw;
    ImageWs::carve(iw, align256(image_buffer), width, height, 1, 1);
    BinWs::carve(bw, align256(binning_buffer), R, P, render_geo ? EWA_PLANES_GEO : EWA_PLANES, EWA_GACC);
    const uint32_t idx_mask = (P < (1 << REC_USED_SHIFT) && !g_no_used_bits) ? REC_INDEX_MASK_USED : ~REC_FLAG_ALWAYS;
    GSR_CUDA_CHECK(cudaMemsetAsync(mismatches, 0, sizeof(int), s));
    GSR_CUDA_CHECK(launch_ewa_audit(vc.gx * vc.gy, iw.tile_offset, bw.planes, bw.plane_stride, width, height, vc.gx, iw.final_T,
                                    iw.n_contrib, idx_mask, margins, info, mismatches, s));
    return GSR_OK;
}

int gsr_depth_normal_forward(int H, int W, const float* depth, const float* kinv, const float* weight, float* normal,
                             void* stream) {
    if (H <= 0 || W <= 0 || !depth || !kinv || !normal) { set_error("gsr_depth_normal_forward: invalid argument"); return GSR_E_INVALID; }
    GSR_CUDA_CHECK(launch_depth_normal_fwd(H, W, depth, kinv, weight, normal, (cudaStream_t)stream));
    return GSR_OK;
}

int gsr_depth_normal_backward(int H, int W, const float* depth, const float* kinv, const float* weight,
                              const float* dL_dnormal, float* scratch, float* dL_ddepth, void* stream) {
    if (H <= 0 || W <= 0 || !depth || !kinv || !dL_dnormal || !scratch || !dL_ddepth) {
        set_error("gsr_depth_normal_backward: invalid argument"); return GSR_E_INVALID;
    }
    GSR_CUDA_CHECK(launch_depth_normal_bwd(H, W, depth, kinv, weight, dL_dnormal, scratch, dL_ddepth, (cudaStream_t)stream));
    return GSR_OK;
}

}  // extern "C"
