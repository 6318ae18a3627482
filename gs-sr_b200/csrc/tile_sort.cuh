// tile_sort.cuh -- per-tile bitonic sort of (depth_bits << 32 | index) keys, shared by the
// record-building kernels of both rasterizer families (see binning.cu for the ordering contract).
#pragma once
#include "common.cuh"

namespace gsr {

constexpr int SORT_SMEM_CAP = 4096;            // keys sorted in shared memory (32 KB)
constexpr uint64_t KEY_INF = ~0ull;

// All-ascending bitonic network on m = pow2 >= n virtual slots; slots >= n hold +inf and are
// never materialised when sorting in global memory (a compare against them is a no-op).
template <bool kShared>
__device__ __forceinline__ void bitonic_sort(uint64_t* __restrict__ k, int n, int m) {
    for (int lsize = 1; (1 << lsize) <= m; lsize++) {
        const int size = 1 << lsize;
        for (int ls = lsize - 1; ls >= 0; ls--) {
            const int stride = 1 << ls;
            const bool first = (ls == lsize - 1);
            // comparators are handled two at a time per thread (all four loads first: ILP)
            for (int t0 = threadIdx.x; t0 < (m >> 1); t0 += 2 * blockDim.x) {
                const int t1 = t0 + blockDim.x;
                // t-th comparator of this step: lower index lo, partner hi > lo
                const int lo0 = ((t0 >> ls) << (ls + 1)) | (t0 & (stride - 1));
                const int hi0 = first ? (lo0 ^ (size - 1)) : (lo0 | stride);
                const int lo1 = ((t1 >> ls) << (ls + 1)) | (t1 & (stride - 1));
                const int hi1 = first ? (lo1 ^ (size - 1)) : (lo1 | stride);
                const bool ok0 = kShared || hi0 < n, ok1 = (t1 < (m >> 1)) && (kShared || hi1 < n);
                uint64_t a0 = 0, b0 = 0, a1 = 0, b1 = 0;
                if (ok0) { a0 = k[lo0]; b0 = k[hi0]; }
                if (ok1) { a1 = k[lo1]; b1 = k[hi1]; }
                if (ok0 && a0 > b0) { k[lo0] = b0; k[hi0] = a0; }
                if (ok1 && a1 > b1) { k[lo1] = b1; k[hi1] = a1; }
            }
            // comparators t = 32w .. 32w+31 of a step with stride <= 32 only touch the 64-slot
            // window [64w, 64w+64): a warp-level sync suffices when this step wrote and the next
            // step reads inside that window
            const int next_stride = ls > 0 ? (stride >> 1) : size;
            if (kShared && stride <= 32 && next_stride <= 32) __syncwarp();
            else __syncthreads();
        }
    }
}

// Sorts tile `tile`'s bucket (in shared memory when it fits, else in place in global memory) and
// returns a pointer to the sorted keys; n = bucket size.  All threads of the CTA must call it.
__device__ __forceinline__ const uint64_t* sort_tile_bucket(uint64_t* __restrict__ gk, int n, uint64_t* skeys) {
    int m = 1;
    while (m < n) m <<= 1;
    if (m <= SORT_SMEM_CAP) {
        for (int i = threadIdx.x; i < m; i += blockDim.x) skeys[i] = i < n ? gk[i] : KEY_INF;
        __syncthreads();
        if (n > 1) bitonic_sort<true>(skeys, n, m);
        return skeys;
    }
    bitonic_sort<false>(gk, n, m);
    return gk;
}

}  // namespace gsr
