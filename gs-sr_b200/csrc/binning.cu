// binning.cu -- tile binning for sm_100a without a global sort.
//
// Ordering contract (identical to the reference, S/cuda_rasterizer/rasterizer_impl.cu:70-138,
// 301-315): within a tile, entries are ordered by (float bits of view depth, Gaussian index) --
// exactly what the reference's stable radix sort of (tile << 32 | depth_bits) keys over pairs
// emitted in ascending Gaussian index produces.  The reference sorts all R pairs globally
// (6 radix passes over 12-byte pairs); here the tile is known when a pair is emitted, so
//   1. the forward preprocess counts pairs per tile (atomicAdd on a tiles-sized histogram; the tile test runs on the
//      warp's -- for big rectangles the CTA's -- flattened (Gaussian, tile) list, cull.cuh),
//   2. tile_scan turns counts into per-tile offsets (one CTA; also yields R, published to the host through a
//      pinned slot so that the host never blocks the stream to learn it, see abi.cu),
//   3. scatter_keys drops each pair's 64-bit (depth_bits << 32 | index) key into its tile's
//      bucket (slot order inside a bucket is arbitrary -- the key is a total order), again over flattened lists,
//   4. sort_build_records: one CTA per tile sorts its bucket (tile_sort.cuh: depth buckets + rank counting in shared
//      memory, through global memory for tiles of more than 2048 entries, bitonic networks as the fallback) and,
//      fused, materialises the tile-local record planes the render kernels stream (six float4 planes, coalesced stores).
// A (Gaussian, tile) pair of the reference's getRect rectangle is only emitted when the splat
// can reach alpha >= 1/255 somewhere in that tile (cull.cuh); dropped pairs are list entries
// on which every pixel of the tile would `continue`.
#include "common.cuh"
#include "cull.cuh"
#include "tile_sort.cuh"

namespace gsr {

// ---- 2. exclusive scan of per-tile counts (single CTA) -------------------------------------
__global__ void __launch_bounds__(1024)
tile_scan(int ntiles, const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
          uint32_t* __restrict__ cursors, uint32_t* __restrict__ total, const int* __restrict__ flags, uint32_t cap,
          volatile uint32_t* host_slot, uint32_t seq, int sticky) {
    // cap = number of list entries the binning buffer was laid out for.  The host sizes that buffer BEFORE it knows R
    // (from earlier frames), so the lists the later kernels see are clamped to it: offsets[] saturate at cap and
    // scatter_keys drops entries at positions >= cap.  R itself goes to the host, which re-runs binning and rendering
    // with an exact buffer in the (rare) case R > cap.
    // One CTA, one block-wide scan: thread t owns the contiguous chunk of `per` tiles starting at t * per, sums it (all
    // its loads in flight at once, kept in registers for chunks of up to 8 tiles), the 1024 chunk sums are scanned with
    // shuffles, and the chunk is walked a second time to write the prefixes.  (The first version scanned 1024 tiles per round with four CTA barriers
    // and a dependent global load per round: 21 us at cfg-B's 6700 tiles, all of it latency, and the host waits for
    // this kernel's last store.)
    __shared__ uint32_t warp_sums[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per = (ntiles + 1023) / 1024;
    const int first = threadIdx.x * per, last = min(first + per, ntiles);
    constexpr int REG_CHUNK = 8;                 // chunks of up to 8 tiles (8192 tiles = 2048 x 1024 pixels) stay in registers
    uint32_t v[REG_CHUNK];
    uint32_t mine = 0;
    if (per <= REG_CHUNK) {
#pragma unroll
        for (int k = 0; k < REG_CHUNK; k++) v[k] = first + k < last ? counts[(size_t)(first + k) * TILE_CTR_STRIDE] : 0u;
#pragma unroll
        for (int k = 0; k < REG_CHUNK; k++) mine += v[k];
    } else {
        for (int i = first; i < last; i++) mine += counts[(size_t)i * TILE_CTR_STRIDE];
    }
    uint32_t x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane], sc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, sc, o);
            if (lane >= o) sc += y;
        }
        warp_sums[lane] = sc - w;  // exclusive prefix of the warp totals
    }
    __syncthreads();
    uint32_t excl = warp_sums[warp] + (x - mine);
    if (per <= REG_CHUNK) {
#pragma unroll
        for (int k = 0; k < REG_CHUNK; k++)
            if (first + k < last) {
                offsets[first + k] = min(excl, cap);
                cursors[(size_t)(first + k) * TILE_CTR_STRIDE] = excl;
                excl += v[k];
            }
    } else {
        for (int i = first; i < last; i++) {
            const uint32_t c = counts[(size_t)i * TILE_CTR_STRIDE];
            offsets[i] = min(excl, cap);
            cursors[(size_t)i * TILE_CTR_STRIDE] = excl;
            excl += c;
        }
    }
    __shared__ uint32_t carry;
    if (threadIdx.x == 1023) carry = excl;      // the last thread's running sum ends at R (empty chunks carry it through)
    __syncthreads();
    // total[0] = R, total[1] = prefiltered-violation flag; the same two words + a sequence number go to the host slot
    if (threadIdx.x == 0) {
        offsets[ntiles] = min(carry, cap); total[0] = carry; total[1] = (uint32_t)flags[0];
        if (host_slot != nullptr) {
            host_slot[1] = carry; host_slot[2] = (uint32_t)flags[0];
            // sticky (forwards recorded into a CUDA graph: nobody waits on the slot, the lists stay clamped): leave the
            // count in slot[3] when it exceeds the capacity the graph was captured with, for gsr_capture_overflow()
            if (sticky && (carry > cap || flags[0])) host_slot[3] = carry > cap ? carry : 0xffffffffu;
            __threadfence_system();
            host_slot[0] = seq;
        }
    }
}

// ---- 3. scatter (depth, index) keys into tile buckets ---------------------------------------
// One thread per Gaussian loads its rectangle and the mask of reachable tiles the preprocess recorded; the (Gaussian, tile)
// pairs of the warp's 32 Gaussians are then flattened into one list (prefix sum of the mask popcounts) and every lane
// takes item base + lane: owner by a 5-step search over the prefix, the item's tile = the k-th set bit of the owner's mask,
// one returning atomicAdd on the tile's cursor, one 8-byte key store.  Four rounds (128 items) issue their atomics before
// the first result is needed.  The kernel is bound by memory round trips, not by issue: a per-thread loop over the mask
// (the round-1 version) needed as many dependent rounds as the busiest of the warp's 32 Gaussians had tiles / 4.
__global__ void __launch_bounds__(256)
scatter_keys(int P, const float* __restrict__ centre_x, const float* __restrict__ centre_y, int centre_stride,
             const CullRec* __restrict__ cull,
             const float* __restrict__ depths, const int* __restrict__ radii, const uint32_t* __restrict__ masks, int gx, int gy,
             uint32_t* __restrict__ cursors, uint64_t* __restrict__ keys, uint32_t cap) {
    const int idx = spread_gaussian_index();
    const int lane = threadIdx.x & 31;
    const bool in_range = idx < P;
    const int li = in_range ? idx : P - 1;           // out-of-range lanes of the last warp load a valid row and ignore it
    // all five per-Gaussian loads are issued together (one memory round trip), then the gates
    const int r = __ldg(radii + li);
    const uint32_t mask = __ldg(masks + li);
    const float cx = __ldg(centre_x + (size_t)li * centre_stride), cy = __ldg(centre_y + (size_t)li * centre_stride);
    const uint32_t dbits = __float_as_uint(__ldg(depths + li));
    const bool live = in_range && r > 0 && mask != 0u;
    const bool retest = live && mask == MASK_RETEST;   // rectangle of more than 32 tiles: no mask, the tile test is repeated
    int x0 = 0, y0 = 0, x1 = 1, y1 = 1;
    if (live) get_rect(cx, cy, r, gx, gy, x0, y0, x1, y1);
    const int w = max(x1 - x0, 1);
    const uint32_t m = (live && !retest) ? mask : 0u;
    const int cnt = __popc(m);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    const int excl = incl - cnt;
    for (int base = 0; base < total; base += 128) {
        uint32_t pos[4], klo[4], khi[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int t = base + u * 32 + lane;
            int o = 0;            // owner = the last lane whose exclusive prefix is <= t
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int e = __shfl_sync(0xffffffffu, excl, o + step);
                if (e <= t) o += step;
            }
            const int k = t - __shfl_sync(0xffffffffu, excl, o);
            const uint32_t mo = __shfl_sync(0xffffffffu, m, o);
            const int xo = __shfl_sync(0xffffffffu, x0, o), yo = __shfl_sync(0xffffffffu, y0, o), wo = __shfl_sync(0xffffffffu, w, o);
            klo[u] = __shfl_sync(0xffffffffu, (uint32_t)idx, o);
            khi[u] = __shfl_sync(0xffffffffu, dbits, o);
            ok[u] = t < total;
            pos[u] = 0u;
            if (ok[u]) {
                const int bit = (int)__fns(mo, 0u, k + 1);             // k-th (0-based) reachable tile of the owner's rectangle
                const int ry = bit / wo, rx = bit - ry * wo;
                pos[u] = atomicAdd(&cursors[(size_t)((yo + ry) * gx + (xo + rx)) * TILE_CTR_STRIDE], 1u);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
            if (ok[u] && pos[u] < cap) keys[pos[u]] = ((uint64_t)khi[u] << 32) | klo[u];
    }
    // ---- rectangles without a mask: the tile test is repeated, flattened over the CTA ----
    // (a per-thread loop over a screen-filling splat's thousands of tiles was a 3 ms tail with 50 such splats among 500 k)
    __shared__ float4 s_big[5][256];          // CullRec q0, q1 | sx, sy, mode, cx | cy, x0, y0, w | key lo, key hi
    __shared__ int s_prefix[257];
    __shared__ int s_wsum[8];
    const int tid = threadIdx.x, warp = tid >> 5;
    if (!__syncthreads_or(retest)) return;
    const int area_big = retest ? w * (y1 - y0) : 0;
    if (retest) {
        const CullRec cr = cull[idx];
        s_big[0][tid] = cr.q0;
        s_big[1][tid] = cr.q1;
        s_big[2][tid] = make_float4(cr.q2.x, cr.q2.y, cr.q2.z, cx);
        s_big[3][tid] = make_float4(cy, __int_as_float(x0), __int_as_float(y0), __int_as_float(w));
        s_big[4][tid] = make_float4(__uint_as_float((uint32_t)idx), __uint_as_float(dbits), 0.f, 0.f);
    }
    int bincl = area_big;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, bincl, o);
        if (lane >= o) bincl += y;
    }
    if (lane == 31) s_wsum[warp] = bincl;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) wbase += k < warp ? s_wsum[k] : 0;
    s_prefix[tid] = wbase + bincl - area_big;
    if (tid == 255) s_prefix[256] = wbase + bincl;
    __syncthreads();
    const int btotal = s_prefix[256];
    for (int base = 0; base < btotal; base += 256) {
        const int t = base + tid;
        if (t < btotal) {
            int o = 0;                                     // the last thread whose exclusive prefix is <= t
#pragma unroll
            for (int step = 128; step > 0; step >>= 1)
                if (s_prefix[o + step] <= t) o += step;
            const int k = t - s_prefix[o];
            CullRec r;
            r.q0 = s_big[0][o];
            r.q1 = s_big[1][o];
            const float4 a = s_big[2][o], b = s_big[3][o], kk = s_big[4][o];
            r.q2 = make_float4(a.x, a.y, a.z, 0.f);
            const int rw = __float_as_int(b.w);
            const int ry = k / rw, rx = k - ry * rw;
            const int tx = __float_as_int(b.y) + rx, ty = __float_as_int(b.z) + ry;
            if (tile_may_contribute(r, a.w, b.x, tx, ty)) {
                const uint32_t pos = atomicAdd(&cursors[(size_t)(ty * gx + tx) * TILE_CTR_STRIDE], 1u);
                if (pos < cap) keys[pos] = ((uint64_t)__float_as_uint(kk.y) << 32) | __float_as_uint(kk.x);
            }
        }
    }
}

// ---- 4. per-tile sort fused with record materialisation --------------------------------------
__global__ void __launch_bounds__(256)
sort_build_records(const uint32_t* __restrict__ offsets, uint64_t* __restrict__ keys,
                   const GeomRec* __restrict__ geom,
                   const float* __restrict__ colors, int gx, int W, int H, float4* __restrict__ planes,
                   size_t pstride, int dbg) {
    __shared__ uint64_t skeys[SORT_SMEM_CAP];
    const int tile = blockIdx.x;
    const uint32_t begin = offsets[tile], end = offsets[tile + 1];
    const int n = (int)(end - begin);
    if (n == 0) return;
    __shared__ BucketSortSmem bs;
    // scratch of a large tile's sort: the tile's slice of record plane 0 (n x 16 bytes, written only after the sort)
    const uint64_t* sorted = sort_tile_bucket(keys + begin, n, skeys, bs, reinterpret_cast<uint64_t*>(planes + begin),
                                              (dbg & 1) != 0);   // dbg bit 0: bitonic only
    const float ox = (float)((tile % gx) * TILE), oy = (float)((tile / gx) * TILE);
#ifndef GSR_BUILD_UNROLL
#define GSR_BUILD_UNROLL 2       // two entries' gathers in flight per thread: 275 -> 252 us at cfg-B (4: 254)
#endif
#define GSR_PRAGMA_(x) _Pragma(#x)
#define GSR_UNROLL_(n) GSR_PRAGMA_(unroll n)
    GSR_UNROLL_(GSR_BUILD_UNROLL)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t g = (uint32_t)(sorted[i] & 0xffffffffull);
        const float4* gp = reinterpret_cast<const float4*>(geom + g);
        const float4 tu = __ldg(gp), tv = __ldg(gp + 1), tw = __ldg(gp + 2), nd = __ldg(gp + 3);
        const float cr = __ldg(colors + 3 * (size_t)g), cg = __ldg(colors + 3 * (size_t)g + 1),
                    cb = __ldg(colors + 3 * (size_t)g + 2);
        const float3 Tw = make_float3(tw.x, tw.y, tw.z);
        const float3 Tu = make_float3(fmaf(-ox, tw.x, tu.x), fmaf(-ox, tw.y, tu.y), fmaf(-ox, tw.z, tu.z));
        const float3 Tv = make_float3(fmaf(-oy, tw.x, tv.x), fmaf(-oy, tw.y, tv.y), fmaf(-oy, tw.z, tv.z));
        const float3 a = cross3(Tv, Tw), b = cross3(Tw, Tu), c = cross3(Tu, Tv);
        const float det = dot3(c, Tw);
        // nd.w = tau (>= 0: conic is an ellipse) or -(tau + 1) (always evaluate)
        const bool always = nd.w < 0.f;
        const float tau = always ? -nd.w - 1.f : nd.w;
        const uint32_t flag = always ? REC_FLAG_ALWAYS : 0u;
        const size_t o = (size_t)begin + i;
        planes[0 * pstride + o] = make_float4(a.x, a.y, a.z, tu.w - ox);
        planes[1 * pstride + o] = make_float4(b.x, b.y, b.z, tv.w - oy);
        planes[2 * pstride + o] = make_float4(c.x, c.y, c.z, tw.w);
        planes[3 * pstride + o] = make_float4(det, tau, tw.z, __uint_as_float(g | flag));
        planes[4 * pstride + o] = make_float4(nd.x, nd.y, nd.z, cr);
        // moment frame of the backward: the splat's rounded, image-clamped screen centre (tile-local)
        const float sx = moment_origin(tu.w, W), sy = moment_origin(tv.w, H);
        planes[5 * pstride + o] = make_float4(cg, cb, sx - ox, sy - oy);
    }
}

}  // namespace gsr
