// binning.cu -- tile binning for sm_100a: (tile | depth) key emission, device radix sort,
// and materialisation of the sorted per-(Gaussian, tile) record planes with per-tile ranges.
//
// Ordering contract (identical to the reference, S/cuda_rasterizer/rasterizer_impl.cu:70-138,
// 301-315): key = tile_id << 32 | float_bits(view depth); stable LSD radix sort over
// bits [0, 32 + ceil_log2(tiles)), so ties keep ascending Gaussian index.
// Difference: a (Gaussian, tile) pair of the reference's getRect rectangle is only emitted
// when the splat can reach alpha >= 1/255 somewhere in that tile (cull.cuh); the dropped
// pairs are exactly list entries on which every pixel of the tile would `continue`.
#include "common.cuh"
#include "cull.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace gsr {

size_t scan_temp_bytes(int P) {
    size_t n = 0;
    cub::DeviceScan::InclusiveSum(nullptr, n, (uint32_t*)nullptr, (uint32_t*)nullptr, P);
    return n;
}

cudaError_t inclusive_scan(char* tmp, size_t tmp_bytes, const uint32_t* in, uint32_t* out, int P,
                           cudaStream_t s) {
    return cub::DeviceScan::InclusiveSum(tmp, tmp_bytes, in, out, P, s);
}

size_t sort_temp_bytes(int64_t R) {
    size_t n = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, n, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (uint32_t*)nullptr, (uint32_t*)nullptr, (int)R);
    return n;
}

cudaError_t sort_pairs(char* tmp, size_t tmp_bytes, const uint64_t* kin, uint64_t* kout,
                       const uint32_t* vin, uint32_t* vout, int64_t R, int end_bit, cudaStream_t s) {
    return cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, (int)R, 0, end_bit, s);
}

// One thread per Gaussian; writes its (key, index) pairs at its scan offset.
__global__ void __launch_bounds__(256)
duplicate_with_keys(int P, const GeomRec* __restrict__ geom, const CullRec* __restrict__ cull,
                    const int* __restrict__ radii, const uint32_t* __restrict__ offsets, int gx, int gy,
                    uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    int r = radii[idx];
    if (r <= 0) return;
    uint32_t off = idx == 0 ? 0u : offsets[idx - 1];
    const uint32_t end = offsets[idx];
    if (off == end) return;
    const float cx = geom[idx].tu.w, cy = geom[idx].tv.w;
    const uint32_t dbits = __float_as_uint(geom[idx].nd.w);
    const CullRec cr = cull[idx];
    int x0, y0, x1, y1;
    get_rect(cx, cy, r, gx, gy, x0, y0, x1, y1);
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            if (!tile_may_contribute(cr, cx, cy, x, y)) continue;
            if (off >= end) return;  // cannot happen (same test as the count); never write past the slot range
            keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
            vals[off] = (uint32_t)idx;
            off++;
        }
}

// One thread per sorted list entry: builds the tile-local record (six float4 planes, coalesced
// stores) and marks tile range boundaries (identifyTileRanges, S/rasterizer_impl.cu:116-138).
__global__ void __launch_bounds__(256)
build_records(int R, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
              const GeomRec* __restrict__ geom, const CullRec* __restrict__ cull,
              const float* __restrict__ colors, int gx, int W, int H, float4* __restrict__ planes, size_t pstride,
              uint2* __restrict__ ranges) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const uint32_t tile = (uint32_t)(keys[i] >> 32);
    if (i == 0) ranges[tile].x = 0;
    else {
        uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
        if (prev != tile) { ranges[prev].y = i; ranges[tile].x = i; }
    }
    if (i == R - 1) ranges[tile].y = R;

    const uint32_t g = vals[i];
    const float4* gp = reinterpret_cast<const float4*>(geom + g);
    const float4 tu = __ldg(gp), tv = __ldg(gp + 1), tw = __ldg(gp + 2), nd = __ldg(gp + 3);
    const float4 c1 = __ldg(&cull[g].q1), c2 = __ldg(&cull[g].q2);
    const float ox = (float)((tile % gx) * TILE), oy = (float)((tile / gx) * TILE);
    const float3 Tw = make_float3(tw.x, tw.y, tw.z);
    const float3 Tu = make_float3(fmaf(-ox, tw.x, tu.x), fmaf(-ox, tw.y, tu.y), fmaf(-ox, tw.z, tu.z));
    const float3 Tv = make_float3(fmaf(-oy, tw.x, tv.x), fmaf(-oy, tw.y, tv.y), fmaf(-oy, tw.z, tv.z));
    const float3 a = cross3(Tv, Tw), b = cross3(Tw, Tu), c = cross3(Tu, Tv);
    const float det = dot3(c, Tw);
    const uint32_t flag = ((int)c2.z == CULL_EXACT) ? 0u : REC_FLAG_ALWAYS;
    planes[0 * pstride + i] = make_float4(a.x, a.y, a.z, tu.w - ox);
    planes[1 * pstride + i] = make_float4(b.x, b.y, b.z, tv.w - oy);
    planes[2 * pstride + i] = make_float4(c.x, c.y, c.z, tw.w);
    planes[3 * pstride + i] = make_float4(det, c1.w, tw.z, __uint_as_float(g | flag));
    const float cr = __ldg(colors + 3 * (size_t)g), cg = __ldg(colors + 3 * (size_t)g + 1),
                cb = __ldg(colors + 3 * (size_t)g + 2);
    planes[4 * pstride + i] = make_float4(nd.x, nd.y, nd.z, cr);
    // moment frame of the backward: the splat's rounded, image-clamped screen centre (tile-local)
    const float sx = moment_origin(tu.w, W), sy = moment_origin(tv.w, H);
    planes[5 * pstride + i] = make_float4(cg, cb, sx - ox, sy - oy);
}

}  // namespace gsr
