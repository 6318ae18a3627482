// mcubes.cu -- triangle mesh of the level set of a TSDF lattice (marching cubes), for sm_100a.
//
// This is the step GS-SR hands its fused volume to:
//   bounded path    /root/reference/gssr/utils/mesh_utils.py:178     volume.extract_triangle_mesh()  (Open3D
//                   ScalableTSDFVolume, an absent third-party dependency)
//   unbounded path  /root/reference/gssr/utils/mcube_utils.py:71-80  skimage.measure.marching_cubes(level=0) on the
//                   host, one 512^3 chunk at a time, after a device->host copy of the chunk
// Conventions follow Open3D's extractor: a corner is inside when f < level, a cell yields triangles only when all eight
// corners are observed (weight > min_weight; no weights = every corner counts, the skimage behaviour), a vertex sits at
// f0 / (f0 - f1) along its lattice edge and is shared by the cells around the edge, colours use the same weight.
// The case table (mc_table.cuh) is derived by gen_mc_table.py and consistent across cell faces: closed surfaces come
// out closed.  Parity against Open3D / skimage is unpinned; oracle/mcubes_oracle.py restates this
// extractor and the kernels reproduce it bit for bit.
//
// B200 design: everything stays on the device and every pass streams the lattice once, x fastest, coalesced.
//   mc_cases  f (+w) -> one case byte per cell (0 = no triangles), triangle count per 1024-voxel block.  A thread owns
//             four consecutive voxels: one 128-bit load (+ the next value) from each of the four lattice rows around them;
//             the rows shared with the neighbouring threads come out of L1/L2, HBM sees each line once.
//   mc_edges  case bytes of the (up to) four cells around each owned lattice edge (+x, +y, +z of a voxel), four voxels per
//             thread with 32-bit loads and byte-parallel bit tricks -> 3-bit vertex mask + the voxel's vertex rank inside
//             its block (uint16), vertex count per block.  No boundary branches: a zeroed guard in front of the case array
//             and the always-empty last cell of every row / slice absorb the reads "before" the lattice.
//   scan_local / scan_add (scan_util.cuh)   exclusive scan of both block-count arrays (4096 counts per CTA), totals to the caller.
//   mc_emit   blocks with nothing to emit leave at once; the others write their vertices (12 B, + 12 B colour) and
//             triangles (12 B) as two DENSE lists: item i is taken by thread i mod 256, which finds its voxel by binary
//             search over the ranks in shared memory (the surface touches ~1 % of the voxels; one thread per voxel would
//             leave one active lane per warp).  A triangle corner's global vertex id is block_base[owner >> 10] +
//             rank(owner) + popc(mask below the axis): no per-voxel 4-byte id array, no atomics, output order = lattice
//             order (deterministic, reproducible by the oracle).
// Algorithmic bytes: 4 (+4) read + 1 written per voxel in mc_cases, 1 + 2 in mc_edges, 3 in mc_emit, + 24 (36) per
// vertex / 12 per triangle -- HBM-bound streaming, nothing to put on tensor cores.
#include "common.cuh"
#include "mc_table.cuh"
#include "scan_util.cuh"
#include "../../include/gsr_b200.h"

namespace gsr {

constexpr int MC_THREADS = 256;
constexpr int MC_PER_THREAD = 4;                             // consecutive voxels along x per thread (one 128-bit load per row)
constexpr int MC_BLOCK_VOX = MC_THREADS * MC_PER_THREAD;     // 1024 voxels per block: vertex ranks < 3072 fit 12 bits
constexpr int MC_BLOCK_SHIFT = 10;
constexpr int MC_CASES_SUB = 2, MC_EDGES_SUB = 4, MC_EMIT_SUB = 8;   // 1024-voxel blocks per CTA of each kernel (more loads in flight,
                                                                   // fewer short-lived CTAs); the counters stay per 1024-voxel block
static_assert(MC_BLOCK_VOX == 1 << MC_BLOCK_SHIFT, "block size / shift");

struct McWorkspace {
    uint8_t* guard;        // zeros in front of `cases`: the cells "before" the lattice (reads at v - 1 - nx - nx*ny ...)
    size_t guard_bytes;
    uint8_t* cases;        // n (+ padding to a multiple of the block)
    uint16_t* ecode;       // n: rank << 3 | mask
    unsigned* tri_base;    // nb + 1
    unsigned* vert_base;   // nb + 1
    unsigned long long* chunk_sums;   // 2 per scan chunk
    unsigned* totals;      // 2
    size_t bytes;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static McWorkspace mc_layout(void* base, int nx, int ny, long long n) {
    const long long nb = (n + MC_BLOCK_VOX - 1) / MC_BLOCK_VOX;
    const long long chunks = (nb + SCAN_CHUNK - 1) / SCAN_CHUNK;
    McWorkspace w;
    size_t off = 0;
    char* b = (char*)base;
    w.guard = (uint8_t*)(b + off);
    w.guard_bytes = align256((size_t)nx * ny + nx + 1);
    off += w.guard_bytes;
    w.cases = (uint8_t*)(b + off);      off += align256((size_t)nb * MC_BLOCK_VOX);
    w.ecode = (uint16_t*)(b + off);     off += align256((size_t)nb * MC_BLOCK_VOX * 2);
    w.tri_base = (unsigned*)(b + off);  off += align256((size_t)(nb + 1) * 4);
    w.vert_base = (unsigned*)(b + off); off += align256((size_t)(nb + 1) * 4);
    w.chunk_sums = (unsigned long long*)(b + off); off += align256((size_t)chunks * 16);
    w.totals = (unsigned*)(b + off);    off += 256;
    w.bytes = off;
    return w;
}

// Division of a 32-bit index by a run-time constant without the ~25-instruction software divide (these kernels are
// issue-bound, not HBM-bound, until the per-voxel instruction count is cut): q = (t + ((n - t) >> 1)) >> sh with
// t = umulhi(n, mul), exact for every 32-bit n (Granlund & Montgomery, round-up variant).
struct FastDiv {
    unsigned d, mul, sh;
};
static FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    f.d = d;
    unsigned l = 0;
    while ((1ull << l) < d) l++;
    f.mul = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.sh = l ? l - 1 : 0;
    if (l == 0) f.mul = 0;                                  // d == 1: t = 0, q = (0 + (n >> 1)) >> 0 is wrong -> handled in fdiv
    return f;
}
__device__ __forceinline__ unsigned fdiv(unsigned n, const FastDiv& f) {
    if (f.d == 1) return n;
    const unsigned t = __umulhi(n, f.mul);
    return (t + ((n - t) >> 1)) >> f.sh;
}
struct Voxel3 {
    int ix, iy, iz;
};
__device__ __forceinline__ Voxel3 voxel_of(unsigned v, const FastDiv& dx, const FastDiv& dy) {
    const unsigned row = fdiv(v, dx), iz = fdiv(row, dy);
    Voxel3 p;
    p.ix = (int)(v - row * dx.d);
    p.iy = (int)(row - iz * dy.d);
    p.iz = (int)iz;
    return p;
}

// Bits j = 0..4: (a[idx + j] < level) for MODE 0, (a[idx + j] > level) for MODE 1; elements at or beyond n read as 0 bits.
template <int MODE>
__device__ __forceinline__ unsigned row_bits5(const float* __restrict__ a, unsigned idx, unsigned n, float level, bool vec) {
    float x[5];                                              // vec: a + idx is 16-byte aligned
    if (vec && idx + 4 < n) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(a + idx));
        x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
        x[4] = __ldg(a + idx + 4);
    } else {
#pragma unroll
        for (int j = 0; j < 5; j++) x[j] = idx + j < n ? __ldg(a + idx + j) : level;
    }
    unsigned b = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) b |= (unsigned)(MODE == 0 ? x[j] < level : x[j] > level) << j;
    return b;
}

// Case byte of every cell (0 = no triangles: outside the lattice, unobserved corner, or all corners on one side) and the
// triangle count of each 1024-voxel block.  A thread owns four consecutive voxels: four (five with the +x neighbour)
// values from each of the rows (y, z), (y+1, z), (y, z+1), (y+1, z+1).
template <bool MASKED>
__global__ void __launch_bounds__(MC_THREADS) mc_cases(const float* __restrict__ f, const float* __restrict__ w, int nx, int ny,
                                                       int nz, unsigned n, unsigned nb, FastDiv dx, FastDiv dy, float level, float min_w,
                                                       uint8_t* __restrict__ cases, unsigned* __restrict__ blk_tris) {
    __shared__ unsigned s_warp[MC_CASES_SUB][MC_THREADS / 32];
    __shared__ uint8_t s_ntri[256];
    s_ntri[threadIdx.x] = MC_NTRI[threadIdx.x];
    __syncthreads();
    const unsigned nxy = (unsigned)nx * (unsigned)ny;
    const bool vec0 = (reinterpret_cast<size_t>(f) & 15) == 0, vec = vec0 && (nx & 3) == 0;       // rows 16-byte aligned?
    const bool wvec0 = MASKED && (reinterpret_cast<size_t>(w) & 15) == 0, wvec = wvec0 && (nx & 3) == 0;
#pragma unroll
    for (int sub = 0; sub < MC_CASES_SUB; sub++) {          // independent 1024-voxel blocks: their loads overlap
        const unsigned v = ((blockIdx.x * MC_CASES_SUB + sub) * MC_THREADS + threadIdx.x) * MC_PER_THREAD;
        unsigned tris = 0;
        if (v < n) {
            unsigned r[4], o[4];
            r[0] = row_bits5<0>(f, v, n, level, vec0);
            r[1] = row_bits5<0>(f, v + nx, n, level, vec);
            r[2] = row_bits5<0>(f, v + nxy, n, level, vec);
            r[3] = row_bits5<0>(f, v + nxy + nx, n, level, vec);
            if (MASKED) {
                o[0] = row_bits5<1>(w, v, n, min_w, wvec0);
                o[1] = row_bits5<1>(w, v + nx, n, min_w, wvec);
                o[2] = row_bits5<1>(w, v + nxy, n, min_w, wvec);
                o[3] = row_bits5<1>(w, v + nxy + nx, n, min_w, wvec);
            }
            // the four rows byte-wise in one word; voxel k's cell reads bits k, k+1 of every byte:
            // ((R >> k) & 0x03030303) * 0x01041040 >> 24 moves byte j's two bits to bits 2j, 2j+1 (no two partial products
            // meet, so no carries)
            const unsigned R = r[0] | (r[1] << 8) | (r[2] << 16) | (r[3] << 24);
            const unsigned O = MASKED ? (o[0] & o[1] & o[2] & o[3]) : 0x1fu;
            const Voxel3 p = voxel_of(v, dx, dy);
            unsigned inside_lattice;                          // bit k: cell k has all eight corners in the lattice
            if (p.ix + MC_PER_THREAD < nx) {
                inside_lattice = (p.iy + 1 < ny && p.iz + 1 < nz) ? 15u : 0u;
            } else {                                          // the row ends (or wraps) inside this thread's four voxels
                inside_lattice = 0;
                int ix = p.ix, iy = p.iy, iz = p.iz;
#pragma unroll
                for (int k = 0; k < MC_PER_THREAD; k++) {
                    if (ix + 1 < nx && iy + 1 < ny && iz + 1 < nz) inside_lattice |= 1u << k;
                    if (++ix == nx) {
                        ix = 0;
                        if (++iy == ny) { iy = 0; iz++; }
                    }
                }
            }
            const unsigned ok = inside_lattice & O & (O >> 1);
            unsigned packed = 0;
#pragma unroll
            for (int k = 0; k < MC_PER_THREAD; k++) {
                unsigned c = (((R >> k) & 0x03030303u) * 0x01041040u) >> 24;
                if (!((ok >> k) & 1u)) c = 0;
                packed |= c << (8 * k);
                tris += s_ntri[c];
            }
            *reinterpret_cast<unsigned*>(cases + v) = packed;
        } else if (blockIdx.x * MC_CASES_SUB + sub < nb) {
            *reinterpret_cast<unsigned*>(cases + v) = 0u;      // padding of the last block: read by mc_edges as "no cell"
        }
        tris = __reduce_add_sync(0xffffffffu, tris);
        if ((threadIdx.x & 31) == 0) s_warp[sub][threadIdx.x >> 5] = tris;
    }
    __syncthreads();
    if (threadIdx.x < MC_CASES_SUB && blockIdx.x * MC_CASES_SUB + threadIdx.x < nb) {
        unsigned total = 0;
#pragma unroll
        for (int k = 0; k < MC_THREADS / 32; k++) total += s_warp[threadIdx.x][k];
        blk_tris[blockIdx.x * MC_CASES_SUB + threadIdx.x] = total;
    }
}

__device__ __forceinline__ unsigned load_bytes4(const uint8_t* p, bool vec) {
    if (vec) return *reinterpret_cast<const unsigned*>(p);
    return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
}
// bit 0 of byte k = (bit k0 of case byte k) xor (bit k1 of case byte k), for the four case bytes of c4 at once
__device__ __forceinline__ unsigned differs4(unsigned c4, int k0, int k1) { return ((c4 >> k0) ^ (c4 >> k1)) & 0x01010101u; }

// 3-bit vertex mask of every voxel (does its +x / +y / +z lattice edge carry a vertex: is it crossed in one of the active
// cells around it) and the voxel's vertex rank inside its block.  Reads below the lattice land in the zeroed guard, reads
// across a row / slice end land on the last cell of the previous row / slice, whose case is always 0.  All four voxels of a
// thread are handled in one 32-bit word, one byte each.
__global__ void __launch_bounds__(MC_THREADS) mc_edges(const uint8_t* __restrict__ cases, int nx, int ny, unsigned n, unsigned nb,
                                                       uint16_t* __restrict__ ecode, unsigned* __restrict__ blk_verts) {
    __shared__ unsigned s_warp[MC_EDGES_SUB][MC_THREADS / 32];
    const unsigned nxy = (unsigned)nx * (unsigned)ny;
    const bool vec = (nx & 3) == 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned m4[MC_EDGES_SUB], inc[MC_EDGES_SUB];          // m4: byte k = the 3-bit mask of voxel k
#pragma unroll
    for (int sub = 0; sub < MC_EDGES_SUB; sub++) {          // independent 1024-voxel blocks: their loads overlap
        const unsigned v = ((blockIdx.x * MC_EDGES_SUB + sub) * MC_THREADS + threadIdx.x) * MC_PER_THREAD;
        m4[sub] = 0;
        if (v < n) {
            const uint8_t* p = cases + v;
            const unsigned c = load_bytes4(p, true);
            const unsigned cy = load_bytes4(p - nx, vec), cz = load_bytes4(p - nxy, vec), cyz = load_bytes4(p - nxy - nx, vec);
            // the same four cells one step down in x: shift in the byte in front of each group
            const unsigned cx = (c << 8) | p[-1];
            const unsigned cxy = (cy << 8) | p[-1 - (int)nx];
            const unsigned cxz = (cz << 8) | *(p - 1 - nxy);
            // the +x edge is edge (0,1) of this cell, (2,3) of the cell below in y, (4,5) below in z, (6,7) below in both
            const unsigned ex = differs4(c, 0, 1) | differs4(cy, 2, 3) | differs4(cz, 4, 5) | differs4(cyz, 6, 7);
            const unsigned ey = differs4(c, 0, 2) | differs4(cx, 1, 3) | differs4(cz, 4, 6) | differs4(cxz, 5, 7);
            const unsigned ez = differs4(c, 0, 4) | differs4(cx, 1, 5) | differs4(cy, 2, 6) | differs4(cxy, 3, 7);
            m4[sub] = ex | (ey << 1) | (ez << 2);
        }
    }
#pragma unroll
    for (int sub = 0; sub < MC_EDGES_SUB; sub++) {
        inc[sub] = warp_inclusive_scan(__popc(m4[sub]));
        if (lane == 31) s_warp[sub][wid] = inc[sub];
    }
    __syncthreads();
#pragma unroll
    for (int sub = 0; sub < MC_EDGES_SUB; sub++) {
        const unsigned blk = blockIdx.x * MC_EDGES_SUB + sub;
        const unsigned v = (blk * MC_THREADS + threadIdx.x) * MC_PER_THREAD;
        unsigned before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < MC_THREADS / 32; k++) {
            const unsigned t = s_warp[sub][k];
            if (k < wid) before += t;
            total += t;
        }
        const unsigned m = m4[sub];
        const unsigned rank = before + inc[sub] - __popc(m);
        if (v < n) {
            const unsigned r1 = rank + __popc(m & 0xffu), r2 = rank + __popc(m & 0xffffu), r3 = rank + __popc(m & 0xffffffu);
            const unsigned lo = ((rank << 3) | (m & 7u)) | (((r1 << 3) | ((m >> 8) & 7u)) << 16);
            const unsigned hi = ((r2 << 3) | ((m >> 16) & 7u)) | (((r3 << 3) | (m >> 24)) << 16);
            *reinterpret_cast<uint2*>(ecode + v) = make_uint2(lo, hi);
        }
        if (threadIdx.x == 0 && blk < nb) blk_verts[blk] = total;
    }
}

// Vertices and triangles.  A block that owns neither leaves at once; in the others a thread looks at its four voxels.
template <bool COLOR>
__global__ void __launch_bounds__(MC_THREADS) mc_emit(const float* __restrict__ f, const float* __restrict__ rgb,
                                                      const uint8_t* __restrict__ cases, const uint16_t* __restrict__ ecode,
                                                      const unsigned* __restrict__ tri_base,
                                                      const unsigned* __restrict__ vert_base, int nx, int ny, unsigned n,
                                                      unsigned nb, FastDiv dx, FastDiv dy, float level, float ox, float oy, float oz, float voxel,
                                                      float* __restrict__ verts, float* __restrict__ colors,
                                                      int* __restrict__ faces) {
    __shared__ unsigned s_warp[MC_EMIT_SUB][MC_THREADS / 32];
    __shared__ unsigned s_tb[MC_EMIT_SUB + 1], s_vb[MC_EMIT_SUB + 1];
    __shared__ uint8_t s_ntri[256];
    __shared__ uint64_t s_tris[256];
    if (threadIdx.x <= MC_EMIT_SUB) {
        const unsigned i = min(blockIdx.x * MC_EMIT_SUB + threadIdx.x, nb);
        s_tb[threadIdx.x] = tri_base[i];
        s_vb[threadIdx.x] = vert_base[i];
    }
    s_ntri[threadIdx.x] = MC_NTRI[threadIdx.x];
    s_tris[threadIdx.x] = MC_TRIS[threadIdx.x];
    __syncthreads();
    if (s_tb[0] == s_tb[MC_EMIT_SUB] && s_vb[0] == s_vb[MC_EMIT_SUB]) return;
    const unsigned nxy = (unsigned)nx * (unsigned)ny;
    // phase 1: the case / edge words of all non-empty sub-blocks in flight at once, then all their triangle scans behind
    // one barrier (a serial walk over the sub-blocks was latency-bound: three dependent round trips each).  Kept in
    // shared memory: phase 2 reads other threads' words.
    __shared__ unsigned s_c4[MC_EMIT_SUB][MC_THREADS], s_pre[MC_EMIT_SUB][MC_THREADS];
    __shared__ uint2 s_e4[MC_EMIT_SUB][MC_THREADS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc[MC_EMIT_SUB];
    {
        unsigned c4s[MC_EMIT_SUB];
        uint2 e4s[MC_EMIT_SUB];
#pragma unroll
        for (int sub = 0; sub < MC_EMIT_SUB; sub++) {
            const unsigned vfirst = ((blockIdx.x * MC_EMIT_SUB + sub) * MC_THREADS + threadIdx.x) * MC_PER_THREAD;
            c4s[sub] = 0;
            e4s[sub] = make_uint2(0u, 0u);
            const bool empty = s_tb[sub] == s_tb[sub + 1] && s_vb[sub] == s_vb[sub + 1];
            if (!empty && vfirst < n) {
                c4s[sub] = *reinterpret_cast<const unsigned*>(cases + vfirst);
                e4s[sub] = *reinterpret_cast<const uint2*>(ecode + vfirst);
            }
        }
#pragma unroll
        for (int sub = 0; sub < MC_EMIT_SUB; sub++) {
            unsigned nt = 0;
#pragma unroll
            for (int k = 0; k < MC_PER_THREAD; k++) nt += s_ntri[(c4s[sub] >> (8 * k)) & 255u];
            inc[sub] = warp_inclusive_scan(nt);
            if (lane == 31) s_warp[sub][wid] = inc[sub];
            inc[sub] -= nt;                                   // exclusive within the warp
            s_c4[sub][threadIdx.x] = c4s[sub];
            s_e4[sub][threadIdx.x] = e4s[sub];
        }
    }
    __syncthreads();
#pragma unroll
    for (int sub = 0; sub < MC_EMIT_SUB; sub++) {
        unsigned before = inc[sub];
#pragma unroll
        for (int k = 0; k < MC_THREADS / 32; k++)
            if (k < wid) before += s_warp[sub][k];
        s_pre[sub][threadIdx.x] = before;                     // triangles of this sub-block before this thread's voxels
    }
    __syncthreads();

    // phase 2: the block's vertices and triangles as two dense lists -- item i goes to thread i mod 256, which finds the
    // voxel it belongs to by binary search (the surface touches ~1 % of the voxels: a thread-per-voxel walk left one
    // active lane per warp and cost 3x the instructions)
#pragma unroll 1
    for (int sub = 0; sub < MC_EMIT_SUB; sub++) {
        const unsigned vox0 = (blockIdx.x * MC_EMIT_SUB + sub) * MC_BLOCK_VOX;
        // ---- vertices: local rank r -> the voxel whose rank range holds r (ranks are non-decreasing over the voxels)
        const unsigned nvert = s_vb[sub + 1] - s_vb[sub];
        const unsigned short* codes = reinterpret_cast<const unsigned short*>(&s_e4[sub][0]);
        // voxels of this sub-block that mc_edges wrote codes for (whole threads): the ranks are monotone over these only
        const unsigned nvox = min((unsigned)MC_BLOCK_VOX, (n - vox0 + MC_PER_THREAD - 1) & ~(unsigned)(MC_PER_THREAD - 1));
        for (unsigned r = threadIdx.x; r < nvert; r += MC_THREADS) {
            unsigned lo = 0, hi = nvox - 1;                    // largest voxel with rank <= r
            while (lo < hi) {
                const unsigned mid = (lo + hi + 1) >> 1;
                if ((unsigned)(codes[mid] >> 3) <= r) lo = mid; else hi = mid - 1;
            }
            const unsigned code = codes[lo];
            const int a = (int)__fns(code & 7u, 0, (int)(r - (code >> 3)) + 1);     // the (r - rank)-th set bit of the mask
            const unsigned v = vox0 + lo;
            const Voxel3 p = voxel_of(v, dx, dy);
            const unsigned vn = v + (a == 0 ? 1u : (a == 1 ? (unsigned)nx : nxy));
            const float f0 = __fsub_rn(__ldg(f + v), level), f1 = __fsub_rn(__ldg(f + vn), level);
            const float t = __fdiv_rn(f0, __fsub_rn(f0, f1));
            const float gx = (float)p.ix, gy = (float)p.iy, gz = (float)p.iz;
            const size_t dst = 3 * (size_t)(s_vb[sub] + r);
            verts[dst + 0] = __fadd_rn(ox, __fmul_rn(a == 0 ? __fadd_rn(gx, t) : gx, voxel));
            verts[dst + 1] = __fadd_rn(oy, __fmul_rn(a == 1 ? __fadd_rn(gy, t) : gy, voxel));
            verts[dst + 2] = __fadd_rn(oz, __fmul_rn(a == 2 ? __fadd_rn(gz, t) : gz, voxel));
            if (COLOR) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    const float c0 = __ldg(rgb + 3 * (size_t)v + ch), c1 = __ldg(rgb + 3 * (size_t)vn + ch);
                    colors[dst + ch] = __fadd_rn(c0, __fmul_rn(t, __fsub_rn(c1, c0)));
                }
            }
        }
        // ---- triangles: local index i -> the thread slot whose prefix range holds i -> the voxel among its four
        const unsigned ntri = s_tb[sub + 1] - s_tb[sub];
        for (unsigned i = threadIdx.x; i < ntri; i += MC_THREADS) {
            unsigned lo = 0, hi = MC_THREADS - 1;              // largest slot with prefix <= i
            while (lo < hi) {
                const unsigned mid = (lo + hi + 1) >> 1;
                if (s_pre[sub][mid] <= i) lo = mid; else hi = mid - 1;
            }
            unsigned r = i - s_pre[sub][lo];
            const unsigned c4 = s_c4[sub][lo];
            unsigned k = 0, c = c4 & 255u;
            while (r >= s_ntri[c]) {
                r -= s_ntri[c];
                k++;
                c = (c4 >> (8 * k)) & 255u;
            }
            const unsigned v = vox0 + lo * MC_PER_THREAD + k;
            const unsigned tri = (unsigned)(s_tris[c] >> (12 * r)) & 0xfffu;
            int id[3];
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const unsigned e = (tri >> (4 * q)) & 15u;
                const unsigned a = e >> 2, j = e & 3u;
                // the edge's owner: this voxel moved along the other two axes (in increasing order) by the bits of j
                const unsigned su = a == 0 ? (unsigned)nx : 1u, sv = a == 2 ? (unsigned)nx : nxy;
                const unsigned owner = v + (j & 1u) * su + (j >> 1) * sv;
                const unsigned oc = ecode[owner];
                id[q] = (int)(vert_base[owner >> MC_BLOCK_SHIFT] + (oc >> 3) + __popc(oc & ((1u << a) - 1u)));
            }
            const size_t dst = 3 * (size_t)(s_tb[sub] + i);
            faces[dst + 0] = id[0];
            faces[dst + 1] = id[1];
            faces[dst + 2] = id[2];
        }
    }
}

}  // namespace gsr

extern "C" size_t gsr_mc_workspace_bytes(int nx, int ny, int nz) {
    if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
    return gsr::mc_layout(nullptr, nx, ny, (long long)nx * ny * nz).bytes;
}

static int mc_check_dims(const char* who, int nx, int ny, int nz, long long* n_out) {
    if (nx <= 0 || ny <= 0 || nz <= 0) {
        gsr::set_error("%s: invalid lattice %d x %d x %d", who, nx, ny, nz);
        return GSR_E_INVALID;
    }
    const long long n = (long long)nx * ny * nz;
    if (n > 0x7fffffffLL - gsr::MC_BLOCK_VOX) {
        gsr::set_error("%s: lattice too large (%lld voxels; extract it in slabs)", who, n);
        return GSR_E_OVERFLOW;
    }
    *n_out = n;
    return GSR_OK;
}

extern "C" int gsr_mc_count(int nx, int ny, int nz, const float* tsdf, const float* weight, float min_weight, float level,
                            void* workspace, long long* nverts, long long* ntris, void* stream_v) {
    using namespace gsr;
    cudaStream_t s = (cudaStream_t)stream_v;
    long long n;
    if (int rc = mc_check_dims("gsr_mc_count", nx, ny, nz, &n)) return rc;
    if (!tsdf || !workspace || !nverts || !ntris) {
        set_error("gsr_mc_count: invalid argument");
        return GSR_E_INVALID;
    }
    const McWorkspace w = mc_layout(workspace, nx, ny, n);
    const unsigned nb = (unsigned)((n + MC_BLOCK_VOX - 1) / MC_BLOCK_VOX);
    const FastDiv dx = make_fastdiv((unsigned)nx), dy = make_fastdiv((unsigned)ny);
    GSR_CUDA_CHECK(cudaMemsetAsync(w.guard, 0, w.guard_bytes, s));
    if (weight)
        mc_cases<true><<<(nb + MC_CASES_SUB - 1) / MC_CASES_SUB, MC_THREADS, 0, s>>>(tsdf, weight, nx, ny, nz, (unsigned)n, nb, dx, dy, level, min_weight,
                                                                                  w.cases, w.tri_base);
    else
        mc_cases<false><<<(nb + MC_CASES_SUB - 1) / MC_CASES_SUB, MC_THREADS, 0, s>>>(tsdf, nullptr, nx, ny, nz, (unsigned)n, nb, dx, dy, level, 0.f,
                                                                                   w.cases, w.tri_base);
    mc_edges<<<(nb + MC_EDGES_SUB - 1) / MC_EDGES_SUB, MC_THREADS, 0, s>>>(w.cases, nx, ny, (unsigned)n, nb, w.ecode, w.vert_base);
    const unsigned chunks = (nb + SCAN_CHUNK - 1) / SCAN_CHUNK;
    scan_local<true><<<chunks, 1024, 0, s>>>(w.tri_base, w.vert_base, nb, w.chunk_sums);
    scan_add<true><<<chunks, 1024, 0, s>>>(w.tri_base, w.vert_base, nb, w.chunk_sums, w.totals);
    GSR_CUDA_CHECK(cudaGetLastError());
    unsigned totals[2];
    GSR_CUDA_CHECK(cudaMemcpyAsync(totals, w.totals, sizeof(totals), cudaMemcpyDeviceToHost, s));
    GSR_CUDA_CHECK(cudaStreamSynchronize(s));
    if (totals[0] > 0x7fffffffu / 3u || totals[1] > 0x7fffffffu / 3u) {
        set_error("gsr_mc_count: mesh too large for 32-bit indices (extract it in slabs)");
        return GSR_E_OVERFLOW;
    }
    *ntris = totals[0];
    *nverts = totals[1];
    return GSR_OK;
}

extern "C" int gsr_mc_emit(int nx, int ny, int nz, const float* tsdf, const float* rgb, float level, const float* origin,
                           float voxel_size, const void* workspace, float* verts, float* colors, int* faces, void* stream_v) {
    using namespace gsr;
    cudaStream_t s = (cudaStream_t)stream_v;
    long long n;
    if (int rc = mc_check_dims("gsr_mc_emit", nx, ny, nz, &n)) return rc;
    if (!tsdf || !workspace || !origin || !(voxel_size > 0.f) || ((rgb != nullptr) != (colors != nullptr))) {
        set_error("gsr_mc_emit: invalid argument");
        return GSR_E_INVALID;
    }
    const McWorkspace w = mc_layout(const_cast<void*>(workspace), nx, ny, n);
    const unsigned nb = (unsigned)((n + MC_BLOCK_VOX - 1) / MC_BLOCK_VOX);
    const FastDiv dx = make_fastdiv((unsigned)nx), dy = make_fastdiv((unsigned)ny);
    if (rgb)
        mc_emit<true><<<(nb + MC_EMIT_SUB - 1) / MC_EMIT_SUB, MC_THREADS, 0, s>>>(tsdf, rgb, w.cases, w.ecode, w.tri_base, w.vert_base, nx, ny, (unsigned)n, nb, dx, dy, level,
                                                origin[0], origin[1], origin[2], voxel_size, verts, colors, faces);
    else
        mc_emit<false><<<(nb + MC_EMIT_SUB - 1) / MC_EMIT_SUB, MC_THREADS, 0, s>>>(tsdf, nullptr, w.cases, w.ecode, w.tri_base, w.vert_base, nx, ny, (unsigned)n, nb, dx, dy, level,
                                                 origin[0], origin[1], origin[2], voxel_size, verts, nullptr, faces);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}
