// tile_sort.cuh -- per-tile sort of (depth_bits << 32 | index) keys, shared by the record-building kernels of both
// rasterizer families (see binning.cu for the ordering contract): a bucketed sort (linear depth buckets, refined by
// histogram equalisation when a surface piles the keys up; rank counting per key) in shared memory for tiles of up to
// 2048 entries and through global memory above, with the bitonic networks below as the fallback for coincident depths.
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

// ---- bucketed sort (n <= BUCKET_SORT_CAP) ------------------------------------------------------------------------
// The keys of a tile are (depth bits, index): depth-dominated and spread over the view frustum.  Instead of a
// 55-step bitonic network over the padded list, the keys are (1) partitioned in shared memory into B monotonic
// depth buckets (block min / max of the depth bits, a linear map, a shared-memory histogram, one scan, one
// scatter) and (2) every key finds its place by rank counting inside its bucket (~4 keys), one thread per key (keys
// are unique, so the ranks are a permutation).
// A linear map suits depths spread over the frustum (the synthetic benchmark scene); a SURFACE puts most of a tile's keys
// into a thin depth shell, i.e. into one or two linear buckets.  When a linear bucket exceeds 32 * BUCKET_RANK_ROUNDS keys
// the partition is therefore refined once (histogram equalisation): every linear bucket gets a number of fine buckets
// proportional to its key count, laid linearly over the depth range its keys actually occupy.  Only if a fine bucket
// still overflows (coincident depths) does the tile go to the bitonic network.
constexpr int BUCKET_SORT_CAP = SORT_SMEM_CAP / 2;    // two key arrays share the 32 KB of skeys[]
constexpr int BUCKET_MAX = 256;                       // linear buckets
constexpr int FINE_MAX = 512;                         // fine buckets of the equalised partition: <= n / 8 + BUCKET_MAX
constexpr int BUCKET_RANK_ROUNDS = 4;                 // a bucket may hold up to 128 keys

struct BucketSortSmem {
    uint32_t cnt[FINE_MAX];         // histogram, then exclusive offsets
    uint32_t cur[FINE_MAX];         // scatter cursors
    uint32_t lmin[BUCKET_MAX];      // equalised partition: smallest / largest depth bits present in a linear bucket,
    uint32_t lmax[BUCKET_MAX];
    uint32_t lbase[BUCKET_MAX];     //   its first fine bucket | its number of fine buckets << 16
    uint32_t dmin, dmax;
    int overflow;
};

// exclusive scan of cnt[0..m) (m <= CAP) by warp 0 into cnt[] and cur[]; sets sm.overflow when a bucket holds more
// than 32 * BUCKET_RANK_ROUNDS keys.  All threads call it (barriers inside).
template <int CAP>
__device__ __forceinline__ void bucket_scan(BucketSortSmem& sm, int m) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        constexpr int PER = CAP / 32;
        uint32_t v[PER], t = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) { const int i = lane * PER + k; v[k] = i < m ? sm.cnt[i] : 0u; t += v[k]; }
        uint32_t x = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        uint32_t e = x - t;
        bool over = false;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int i = lane * PER + k;
            over |= v[k] > 32u * BUCKET_RANK_ROUNDS;
            if (i < m) { sm.cnt[i] = e; sm.cur[i] = e; }
            e += v[k];
        }
        if (__any_sync(0xffffffffu, over) && lane == 0) sm.overflow = 1;
    }
    __syncthreads();
}

// scatter src -> dst by bucket, then rank counting inside each bucket, ONE THREAD PER KEY: thread i takes the key at
// position i of the partitioned array, re-derives its bucket, counts the smaller keys of that bucket (keys are unique, so
// the ranks are a permutation) and writes the key to its final position in the OTHER array (src is free once the
// partition is done).  A warp per bucket kept 4-9 of 32 lanes busy on ~4-key buckets and made up 74 % of the kernel's
// instructions (ncu, profiles/r02a_*).
template <typename BucketFn>
__device__ __forceinline__ void bucket_scatter_rank(uint64_t* __restrict__ src, uint64_t* __restrict__ dst, int n,
                                                    BucketSortSmem& sm, BucketFn bucket_of) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t k = src[i];
        dst[atomicAdd(&sm.cur[bucket_of(k)], 1u)] = k;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t mine = dst[i];
        const int bkt = bucket_of(mine);
        const int first = (int)sm.cnt[bkt], end = (int)sm.cur[bkt];      // cursor ended at first + count
        int rank = 0;
        for (int j = first; j < end; j++) rank += dst[j] < mine;
        src[first + rank] = mine;
    }
    __syncthreads();
}

// keys in src[0..n) (shared); dst[0..n) (shared) is scratch.  Returns true with the sorted keys back in src[0..n), or
// false (src intact) when a bucket overflows.  All threads of the 256-thread CTA must call it.
__device__ __forceinline__ bool bucket_sort_256(uint64_t* __restrict__ src, uint64_t* __restrict__ dst, int n,
                                                BucketSortSmem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // number of buckets: ~4 keys each, power of two in [8, BUCKET_MAX]
    int B = 8;
    while (B < BUCKET_MAX && B * 4 < n) B <<= 1;
    if (threadIdx.x == 0) { sm.dmin = 0xffffffffu; sm.dmax = 0u; sm.overflow = 0; }
    for (int i = threadIdx.x; i < BUCKET_MAX; i += blockDim.x) { sm.cnt[i] = 0u; }
    __syncthreads();
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t d = (uint32_t)(src[i] >> 32);
        lo = min(lo, d); hi = max(hi, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { atomicMin(&sm.dmin, lo); atomicMax(&sm.dmax, hi); }
    __syncthreads();
    const uint32_t dmin = sm.dmin;
    // bucket = floor(offset * B / span) evaluated in float32: conversion and multiplication are monotonic, which is
    // all the partition needs (the order inside a bucket comes from the exact 64-bit keys); clamped to B - 1
    const float scale = (float)B / ((float)(sm.dmax - dmin) + 1.0f);
    const int Bm1 = B - 1;
    auto bucket_of = [&](uint64_t key) -> int {
        const uint32_t off = (uint32_t)(key >> 32) - dmin;
        return min((int)(__uint2float_rz(off) * scale), Bm1);
    };
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&sm.cnt[bucket_of(src[i])], 1u);
    __syncthreads();
    bucket_scan<BUCKET_MAX>(sm, B);
    if (!sm.overflow) {
        bucket_scatter_rank(src, dst, n, sm, bucket_of);
        return true;
    }
    // ---- equalised partition: the linear buckets become ranges of fine buckets ----
    for (int i = threadIdx.x; i < B; i += blockDim.x) { sm.lmin[i] = 0xffffffffu; sm.lmax[i] = 0u; }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t k = src[i];
        const int b = bucket_of(k);
        atomicMin(&sm.lmin[b], (uint32_t)(k >> 32));
        atomicMax(&sm.lmax[b], (uint32_t)(k >> 32));
    }
    __syncthreads();
    if (warp == 0) {                                       // fine buckets per linear bucket: ~8 keys each; exclusive scan
        constexpr int PER = BUCKET_MAX / 32;
        uint32_t m[PER], t = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int i = lane * PER + k;
            uint32_t c = 0;
            if (i < B) c = (i + 1 < B ? sm.cnt[i + 1] : (uint32_t)n) - sm.cnt[i];     // cnt[] holds the exclusive offsets
            m[k] = (c + 7u) >> 3;
            t += m[k];
        }
        uint32_t x = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        uint32_t e = x - t;
        __syncwarp();                                      // every lane has read its cnt[] entries (incl. the neighbour's first)
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int i = lane * PER + k;
            if (i < B) sm.lbase[i] = e | (m[k] << 16);
            e += m[k];
        }
        if (lane == 31) sm.dmax = e;                       // total number of fine buckets (dmax is not needed any more)
        if (lane == 0) sm.overflow = 0;
    }
    __syncthreads();
    const int M = (int)sm.dmax;                            // <= n / 8 + B <= FINE_MAX
    for (int i = threadIdx.x; i < M; i += blockDim.x) sm.cnt[i] = 0u;
    __syncthreads();
    auto fine_of = [&](uint64_t key) -> int {
        const uint32_t d = (uint32_t)(key >> 32);
        const int b = bucket_of(key);
        const uint32_t lb = sm.lbase[b], l0 = sm.lmin[b];
        const int mb = (int)(lb >> 16);
        const float fs = (float)mb / ((float)(sm.lmax[b] - l0) + 1.0f);
        return (int)(lb & 0xffffu) + min((int)(__uint2float_rz(d - l0) * fs), mb - 1);
    };
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&sm.cnt[fine_of(src[i])], 1u);
    __syncthreads();
    bucket_scan<FINE_MAX>(sm, M);
    if (sm.overflow) return false;
    bucket_scatter_rank(src, dst, n, sm, fine_of);
    return true;
}

// ---- bucketed sort of a LARGE tile (n > BUCKET_SORT_CAP), keys in global memory ----------------------------------
// The same steps as bucket_sort_256 -- linear depth buckets, a counting scatter, rank counting per key inside its
// bucket, and the equalised refinement when a linear bucket overflows -- with the two key arrays in global memory (gk:
// the tile's keys, gtmp: n words of scratch) and the bucket counters in the 32 KB that hold the keys of a small tile.
// A tile of 12 000 entries took 585 us in the 105-step global-memory bitonic network this replaces (one CTA, two
// dependent global round trips per step); real scenes have such tiles even where the blend terminates early, because
// the whole list must be in order first.  Returns true with the sorted keys back in gk, or false (gk intact) when a
// bucket still exceeds BIG_BUCKET_MAX keys (coincident depths: the rank loop is quadratic in the bucket size).
constexpr int BIG_BUCKETS_MAX = 2048;                 // linear buckets
constexpr int BIG_BUCKET_MAX = 256;                   // keys per bucket the rank loop accepts
constexpr int BIG_L1 = 256;                           // equalised partition: level-1 (linear) buckets ...
constexpr int BIG_FINE_MAX = 1792;                    // ... and fine buckets: <= 1536 + BIG_L1

// exclusive scan of cnt[0..m) into cnt[] and cur[] by warp 0 (any m); *flag = 1 when an entry exceeds `limit`.
__device__ __forceinline__ void big_scan(uint32_t* cnt, uint32_t* cur, int m, uint32_t limit, uint32_t* flag) {
    if ((threadIdx.x >> 5) == 0) {
        const int lane = threadIdx.x & 31;
        const int per = (m + 31) / 32;
        const int i0 = lane * per, i1 = min(i0 + per, m);
        uint32_t t = 0;
        bool over = false;
        for (int i = i0; i < i1; i++) { const uint32_t c = cnt[i]; t += c; over |= c > limit; }
        uint32_t x = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        uint32_t e = x - t;
        for (int i = i0; i < i1; i++) { const uint32_t c = cnt[i]; cnt[i] = e; cur[i] = e; e += c; }
        if (over) *flag = 1u;
    }
    __syncthreads();
}

template <typename BucketFn>
__device__ __forceinline__ void big_scatter_rank(uint64_t* __restrict__ gk, uint64_t* __restrict__ gtmp, int n,
                                                 const uint32_t* cnt, uint32_t* cur, BucketFn bucket_of) {
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t k = gk[i];
        gtmp[atomicAdd(&cur[bucket_of(k)], 1u)] = k;
    }
    __syncthreads();                                        // (block scope: the CTA's global writes are visible to it)
#pragma unroll 2
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t mine = gtmp[i];
        const int bkt = bucket_of(mine);
        const int first = (int)cnt[bkt], end = (int)cur[bkt];
        int rank = 0;
        for (int j = first; j < end; j++) rank += gtmp[j] < mine;
        gk[first + rank] = mine;
    }
    __syncthreads();
}

// smem: 32 KB = 8192 words of scratch (the key arrays of the small-tile path)
__device__ __forceinline__ bool bucket_sort_global(uint64_t* __restrict__ gk, uint64_t* __restrict__ gtmp, int n, uint32_t* smem) {
    const int lane = threadIdx.x & 31;
    uint32_t* scnt = smem;                                  // [BIG_BUCKETS_MAX]   (equalised: [BIG_FINE_MAX])
    uint32_t* scur = scnt + BIG_BUCKETS_MAX;                // [BIG_BUCKETS_MAX]
    uint32_t* lmin = scur + BIG_BUCKETS_MAX;                // [BIG_L1] each, equalised partition only
    uint32_t* lmax = lmin + BIG_L1;
    uint32_t* lbase = lmax + BIG_L1;
    uint32_t* l1cnt = lbase + BIG_L1;
    uint32_t* sflags = l1cnt + BIG_L1;                      // [0] min depth bits [1] max [2] overflow flag [3] fine buckets
    int B = 64;                                             // ~8 keys per bucket, power of two in [64, BIG_BUCKETS_MAX]
    while (B < BIG_BUCKETS_MAX && B * 8 < n) B <<= 1;
    if (threadIdx.x == 0) { sflags[0] = 0xffffffffu; sflags[1] = 0u; sflags[2] = 0u; }
    for (int i = threadIdx.x; i < B; i += blockDim.x) scnt[i] = 0u;
    __syncthreads();
    uint32_t lo = 0xffffffffu, hi = 0u;
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t d = (uint32_t)(gk[i] >> 32);
        lo = min(lo, d); hi = max(hi, d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { atomicMin(&sflags[0], lo); atomicMax(&sflags[1], hi); }
    __syncthreads();
    const uint32_t dmin = sflags[0];
    const float span1 = (float)(sflags[1] - dmin) + 1.0f;
    const float scale = (float)B / span1;
    const int Bm1 = B - 1;
    auto bucket_of = [&](uint64_t key) -> int {            // monotonic in the depth bits, as in bucket_sort_256
        const uint32_t off = (uint32_t)(key >> 32) - dmin;
        return min((int)(__uint2float_rz(off) * scale), Bm1);
    };
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&scnt[bucket_of(gk[i])], 1u);
    __syncthreads();
    big_scan(scnt, scur, B, (uint32_t)BIG_BUCKET_MAX, &sflags[2]);
    if (!sflags[2]) {
        big_scatter_rank(gk, gtmp, n, scnt, scur, bucket_of);
        return true;
    }
    // ---- equalised partition over BIG_L1 linear level-1 buckets (see bucket_sort_256) ----
    const float scale1 = (float)BIG_L1 / span1;
    auto l1_of = [&](uint64_t key) -> int {
        const uint32_t off = (uint32_t)(key >> 32) - dmin;
        return min((int)(__uint2float_rz(off) * scale1), BIG_L1 - 1);
    };
    __syncthreads();
    for (int i = threadIdx.x; i < BIG_L1; i += blockDim.x) { l1cnt[i] = 0u; lmin[i] = 0xffffffffu; lmax[i] = 0u; }
    if (threadIdx.x == 0) sflags[2] = 0u;
    __syncthreads();
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t k = gk[i];
        const int b = l1_of(k);
        atomicAdd(&l1cnt[b], 1u);
        atomicMin(&lmin[b], (uint32_t)(k >> 32));
        atomicMax(&lmax[b], (uint32_t)(k >> 32));
    }
    __syncthreads();
    const uint32_t per_fine = max(8u, ((uint32_t)n + 1535u) / 1536u);     // keys per fine bucket: at most 1536 + BIG_L1 of them
    if ((threadIdx.x >> 5) == 0) {
        constexpr int PER = BIG_L1 / 32;
        uint32_t m[PER], t = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) { m[k] = (l1cnt[lane * PER + k] + per_fine - 1u) / per_fine; t += m[k]; }
        uint32_t x = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        uint32_t e = x - t;
#pragma unroll
        for (int k = 0; k < PER; k++) { lbase[lane * PER + k] = e | (m[k] << 16); e += m[k]; }
        if (lane == 31) sflags[3] = e;
    }
    __syncthreads();
    const int M = (int)sflags[3];
    for (int i = threadIdx.x; i < M; i += blockDim.x) scnt[i] = 0u;
    __syncthreads();
    auto fine_of = [&](uint64_t key) -> int {
        const uint32_t d = (uint32_t)(key >> 32);
        const int b = l1_of(key);
        const uint32_t lb = lbase[b], l0 = lmin[b];
        const int mb = (int)(lb >> 16);
        const float fs = (float)mb / ((float)(lmax[b] - l0) + 1.0f);
        return (int)(lb & 0xffffu) + min((int)(__uint2float_rz(d - l0) * fs), mb - 1);
    };
#pragma unroll 4
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&scnt[fine_of(gk[i])], 1u);
    __syncthreads();
    big_scan(scnt, scur, M, (uint32_t)BIG_BUCKET_MAX, &sflags[2]);
    if (sflags[2]) return false;
    big_scatter_rank(gk, gtmp, n, scnt, scur, fine_of);
    return true;
}

// Sorts tile `tile`'s bucket (in shared memory when it fits, else in place in global memory) and
// returns a pointer to the sorted keys; n = bucket size.  All threads of the CTA must call it.
// skeys: SORT_SMEM_CAP keys of shared memory; bs: scratch of the bucketed path.
// gtmp: n words of global scratch for large tiles (the callers pass the tile's own, not yet written, record plane).
__device__ __forceinline__ const uint64_t* sort_tile_bucket(uint64_t* __restrict__ gk, int n, uint64_t* skeys,
                                                            BucketSortSmem& bs, uint64_t* __restrict__ gtmp,
                                                            bool force_bitonic = false) {
    if (n > 32 && n <= BUCKET_SORT_CAP && !force_bitonic) {
        uint64_t* a = skeys;
        uint64_t* b = skeys + BUCKET_SORT_CAP;
#pragma unroll 4
        for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = gk[i];      // all of a thread's key loads in flight together
        __syncthreads();
        if (bucket_sort_256(a, b, n, bs)) return a;
        // degenerate depth distribution: fall through to the bitonic network on the keys still in a[]
        int m = 1;
        while (m < n) m <<= 1;
        for (int i = n + threadIdx.x; i < m; i += blockDim.x) a[i] = KEY_INF;
        __syncthreads();
        bitonic_sort<true>(a, n, m);
        return a;
    }
    if (n > BUCKET_SORT_CAP && !force_bitonic && gtmp != nullptr) {
        // large tile: bucketed sort through global memory; the 32 KB of skeys[] hold the bucket counters / cursors / flags
        static_assert(2 * BIG_BUCKETS_MAX + 4 * BIG_L1 + 8 <= 2 * SORT_SMEM_CAP && BIG_FINE_MAX <= BIG_BUCKETS_MAX, "scratch");
        if (bucket_sort_global(gk, gtmp, n, reinterpret_cast<uint32_t*>(skeys))) return gk;
        __syncthreads();                                    // degenerate distribution: the bitonic paths below (gk intact)
    }
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
