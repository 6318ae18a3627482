// scan_util.cuh -- block / grid exclusive scans of small counters, shared by the mesh kernels (mcubes.cu, mesh_clusters.cu).
#pragma once
#include "common.cuh"

namespace gsr {

constexpr int SCAN_THREADS = 256;      // block size of the kernels that call block_exclusive_scan
constexpr int SCAN_CHUNK = 4096;       // counters scanned per CTA of scan_local

__device__ __forceinline__ unsigned warp_inclusive_scan(unsigned x) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    return x;
}

// exclusive scan of one value per thread over the 256-thread block; *total = block sum.  s_warp: 8 words.
__device__ __forceinline__ unsigned block_exclusive_scan(unsigned x, unsigned* s_warp, unsigned* total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += y;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    unsigned before = 0, sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; k++) {
        const unsigned t = s_warp[k];
        if (k < wid) before += t;
        sum += t;
    }
    *total = sum;
    return before + inc - x;
}

// Exclusive scan of one (TWO = false: b is ignored) or two block-count arrays of nb entries, two small launches.  scan_local: CTA c scans entries
// [4096 c, 4096 c + 4096) of a and b in place and leaves the chunk sums; scan_add: CTA c adds the sums of the
// chunks before it; the last CTA writes a[nb] / b[nb] / totals (saturated to 0xffffffff).
template <bool TWO>
__global__ void __launch_bounds__(1024) scan_local(unsigned* __restrict__ a, unsigned* __restrict__ b, unsigned nb,
                                                      unsigned long long* __restrict__ chunk_sums) {
    __shared__ unsigned s_a[32], s_b[32];
    const unsigned i0 = blockIdx.x * SCAN_CHUNK + threadIdx.x * 4;
    unsigned xa[4], xb[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        xa[k] = i0 + k < nb ? a[i0 + k] : 0u;
        xb[k] = TWO && i0 + k < nb ? b[i0 + k] : 0u;
    }
    const unsigned ta = xa[0] + xa[1] + xa[2] + xa[3], tb = xb[0] + xb[1] + xb[2] + xb[3];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned ia = ta, ib = tb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned ya = __shfl_up_sync(0xffffffffu, ia, d), yb = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) { ia += ya; ib += yb; }
    }
    if (lane == 31) { s_a[wid] = ia; s_b[wid] = ib; }
    __syncthreads();
    if (wid == 0) {
        unsigned wa = s_a[lane], wb = s_b[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned ya = __shfl_up_sync(0xffffffffu, wa, d), yb = __shfl_up_sync(0xffffffffu, wb, d);
            if (lane >= d) { wa += ya; wb += yb; }
        }
        s_a[lane] = wa;
        s_b[lane] = wb;
    }
    __syncthreads();
    unsigned ra = ia - ta + (wid ? s_a[wid - 1] : 0u), rb = ib - tb + (wid ? s_b[wid - 1] : 0u);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (i0 + k < nb) {
            a[i0 + k] = ra;
            if (TWO) b[i0 + k] = rb;
        }
        ra += xa[k];
        rb += xb[k];
    }
    if (threadIdx.x == 1023) {                      // a chunk holds at most 4096 * 5120 triangles: no 32-bit overflow
        chunk_sums[2 * blockIdx.x] = s_a[31];
        chunk_sums[2 * blockIdx.x + 1] = s_b[31];
    }
}

template <bool TWO>
__global__ void __launch_bounds__(1024) scan_add(unsigned* __restrict__ a, unsigned* __restrict__ b, unsigned nb,
                                                    const unsigned long long* __restrict__ chunk_sums,
                                                    unsigned* __restrict__ totals) {
    __shared__ unsigned long long s_a[32], s_b[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool last = blockIdx.x == gridDim.x - 1;
    const unsigned upto = blockIdx.x + (last ? 1u : 0u);      // the last CTA also needs the grand total
    unsigned long long sa = 0, sb = 0, la = 0, lb = 0;
    for (unsigned c = threadIdx.x; c < upto; c += 1024) {
        const unsigned long long ca = chunk_sums[2 * c], cb = chunk_sums[2 * c + 1];
        if (c < blockIdx.x) { sa += ca; sb += cb; } else { la = ca; lb = cb; }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, d);
        sb += __shfl_xor_sync(0xffffffffu, sb, d);
        la += __shfl_xor_sync(0xffffffffu, la, d);
        lb += __shfl_xor_sync(0xffffffffu, lb, d);
    }
    if (lane == 0) { s_a[wid] = sa; s_b[wid] = sb; }
    __syncthreads();
    unsigned long long base_a = 0, base_b = 0;
#pragma unroll
    for (int k = 0; k < 32; k++) { base_a += s_a[k]; base_b += s_b[k]; }
    __syncthreads();
    if (lane == 0) { s_a[wid] = la; s_b[wid] = lb; }
    __syncthreads();
    const unsigned i0 = blockIdx.x * SCAN_CHUNK + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (i0 + k < nb) {
            a[i0 + k] += (unsigned)base_a;          // wraps only when the total saturates, which the caller rejects
            if (TWO) b[i0 + k] += (unsigned)base_b;
        }
    if (last && threadIdx.x == 0) {
        unsigned long long ta = base_a, tb = base_b;
        for (int k = 0; k < 32; k++) { ta += s_a[k]; tb += s_b[k]; }
        a[nb] = totals[0] = ta > 0xfffffffeull ? 0xffffffffu : (unsigned)ta;
        totals[1] = tb > 0xfffffffeull ? 0xffffffffu : (unsigned)tb;
        if (TWO) b[nb] = totals[1];
    }
}

}  // namespace gsr
