// binning.cu -- tile binning for sm_100a: (tile | depth) key emission, device
// radix sort, and materialisation of the sorted per-(Gaussian, tile) record
// stream with per-tile ranges.
//
// Ordering contract (identical to the reference, S/cuda_rasterizer/rasterizer_impl.cu:70-138,
// 301-315): key = tile_id << 32 | float_bits(view depth); stable LSD radix sort over
// bits [0, 32 + ceil_log2(tiles)), so ties keep ascending Gaussian index.
#include "common.cuh"
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

// One thread per Gaussian; writes its tiles_touched (key, index) pairs at its scan offset.
__global__ void __launch_bounds__(256)
duplicate_with_keys(int P, const GeomRec* __restrict__ geom, const int* __restrict__ radii,
                    const uint32_t* __restrict__ offsets, int gx, int gy,
                    uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    int r = radii[idx];
    if (r <= 0) return;
    uint32_t off = idx == 0 ? 0u : offsets[idx - 1];
    float cx = geom[idx].tu.w, cy = geom[idx].tv.w;
    uint32_t dbits = __float_as_uint(geom[idx].nd.w);
    int x0, y0, x1, y1;
    get_rect(cx, cy, r, gx, gy, x0, y0, x1, y1);
    for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
            keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
            vals[off] = (uint32_t)idx;
            off++;
        }
}

// One thread per sorted list entry: builds the tile-local 80-byte record and marks
// tile range boundaries (identifyTileRanges, S/rasterizer_impl.cu:116-138).
__global__ void __launch_bounds__(256)
build_records(int R, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
              const GeomRec* __restrict__ geom, const float4* __restrict__ cbox,
              const float* __restrict__ colors, int gx, SplatRec* __restrict__ recs,
              uint2* __restrict__ ranges) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    uint32_t tile = (uint32_t)(keys[i] >> 32);
    if (i == 0) ranges[tile].x = 0;
    else {
        uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
        if (prev != tile) { ranges[prev].y = i; ranges[tile].x = i; }
    }
    if (i == R - 1) ranges[tile].y = R;

    uint32_t g = vals[i];
    const float4* gp = reinterpret_cast<const float4*>(geom + g);
    float4 tu = __ldg(gp), tv = __ldg(gp + 1), tw = __ldg(gp + 2), nd = __ldg(gp + 3);
    float4 cb = __ldg(cbox + g);
    float ox = (float)((tile % gx) * TILE), oy = (float)((tile / gx) * TILE);
    SplatRec r;
    r.tu = make_float4(fmaf(-ox, tw.x, tu.x), fmaf(-ox, tw.y, tu.y), fmaf(-ox, tw.z, tu.z), tu.w - ox);
    r.tv = make_float4(fmaf(-oy, tw.x, tv.x), fmaf(-oy, tw.y, tv.y), fmaf(-oy, tw.z, tv.z), tv.w - oy);
    r.tw = tw;
    r.ng = make_float4(nd.x, nd.y, nd.z, __uint_as_float(g));
    // conservative pixel bounds, clipped to this tile (local 0..15)
    float bx0 = cb.x - ox, bx1 = cb.y - ox, by0 = cb.z - oy, by1 = cb.w - oy;
    int xmin = (int)ceilf(fmaxf(bx0, 0.f)), xmax = (int)floorf(fminf(bx1, 15.f));
    int ymin = (int)ceilf(fmaxf(by0, 0.f)), ymax = (int)floorf(fminf(by1, 15.f));
    uint32_t bounds = (xmin > xmax || ymin > ymax || !(bx0 <= bx1) || !(by0 <= by1))
                          ? BOUNDS_EMPTY : pack_bounds(xmin, xmax, ymin, ymax);
    r.cb = make_float4(__ldg(colors + 3 * (size_t)g), __ldg(colors + 3 * (size_t)g + 1),
                       __ldg(colors + 3 * (size_t)g + 2), __uint_as_float(bounds));
    float4* rp = reinterpret_cast<float4*>(recs + i);
    rp[0] = r.tu; rp[1] = r.tv; rp[2] = r.tw; rp[3] = r.ng; rp[4] = r.cb;
}

}  // namespace gsr
