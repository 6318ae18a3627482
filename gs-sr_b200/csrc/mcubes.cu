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
//   mc_signs  f (+w) -> one bit per voxel (f < level, weight > min_weight): the only pass over the float lattices, a
//             coalesced load, a compare and a warp vote per voxel -- HBM-bound.
//   mc_cases  sign / observed bits -> one case byte per cell (0 = no triangles), triangle count per 1024-voxel block.  A
//             thread owns 32 consecutive voxels: 33 bits from each of the four lattice rows around them (funnel shifts of
//             three words), a multiply gathers a cell's eight corner bits.
//   mc_edges  case bytes of the (up to) four cells around each owned lattice edge (+x, +y, +z of a voxel), sixteen voxels
//             per thread with 128-bit loads and byte-parallel bit tricks -> 3-bit vertex mask + the voxel's vertex rank
//             inside its block (uint16), vertex count per block.  No boundary branches: a zeroed guard in front of the case array
//             and the always-empty last cell of every row / slice absorb the reads "before" the lattice.
//   scan_local / scan_add (scan_util.cuh)   exclusive scan of both block-count arrays (4096 counts per CTA), totals to the caller.
//   mc_list_blocks / mc_emit   the blocks that own something are listed; persistent 64-thread CTAs take one listed block
//             per turn (many blocks in flight per SM, the next block's words requested before the current one is
//             processed) and write its vertices (12 B, + 12 B colour) and triangles (12 B) as two DENSE lists: item i is
//             taken by thread i mod 64, which finds its voxel by binary search over the ranks in shared memory (the
//             surface touches ~1 % of the voxels; one thread per voxel would leave one active lane per warp).  A triangle
//             corner's global vertex id is block_base[owner >> 10] + rank(owner) + popc(mask below the axis): no
//             per-voxel 4-byte id array, no atomics on the outputs, output order = lattice order (deterministic,
//             reproducible by the oracle).
// Algorithmic bytes: 4 (+4) read per voxel in mc_signs, 1 written in mc_cases, 1 + 2 in mc_edges, 3 in mc_emit, + 24 (36) per
// vertex / 12 per triangle -- HBM-bound streaming, nothing to put on tensor cores.
#include <algorithm>
#include "common.cuh"
#include "mc_table.cuh"
#include "scan_util.cuh"
#include "../../include/gsr_b200.h"

namespace gsr {

constexpr int MC_THREADS = 256;
constexpr int MC_PER_THREAD = 4;                             // consecutive voxels along x per thread (one 128-bit load per row)
constexpr int MC_BLOCK_VOX = MC_THREADS * MC_PER_THREAD;     // 1024 voxels per block: vertex ranks < 3072 fit 12 bits
constexpr int MC_BLOCK_SHIFT = 10;
static_assert(MC_BLOCK_VOX == 1 << MC_BLOCK_SHIFT, "block size / shift");

struct McWorkspace {
    uint8_t* guard;        // zeros in front of `cases`: the cells "before" the lattice (reads at v - 1 - nx - nx*ny ...)
    size_t guard_bytes;
    uint8_t* cases;        // n (+ padding to a multiple of the block)
    uint16_t* ecode;       // n: rank << 3 | mask
    unsigned* inside_bits; // one bit per voxel: f < level (+ 2 zero words of padding)
    unsigned* ok_bits;     // one bit per voxel: weight > min_weight
    size_t bit_words;      // words per bit array incl. padding
    unsigned* tri_base;    // nb + 1
    unsigned* vert_base;   // nb + 1
    unsigned long long* chunk_sums;   // 2 per scan chunk
    unsigned* block_list;  // nb: the blocks that own a vertex or a triangle
    unsigned* totals;      // 0: triangles, 1: vertices, 2: listed blocks
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
    w.bit_words = (size_t)nb * (MC_BLOCK_VOX / 32) + 2;
    w.inside_bits = (unsigned*)(b + off); off += align256(w.bit_words * 4);
    w.ok_bits = (unsigned*)(b + off);     off += align256(w.bit_words * 4);
    w.tri_base = (unsigned*)(b + off);  off += align256((size_t)(nb + 1) * 4);
    w.vert_base = (unsigned*)(b + off); off += align256((size_t)(nb + 1) * 4);
    w.chunk_sums = (unsigned long long*)(b + off); off += align256((size_t)chunks * 16);
    w.block_list = (unsigned*)(b + off); off += align256((size_t)nb * 4);
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

// One bit per voxel: f < level (and weight > min_weight), the only pass that reads the float lattices -- one coalesced
// 4-byte load, one compare and 1/32 of a vote + store per voxel: HBM-bound.  Words of the last block beyond n get zeros.
template <bool MASKED>
__global__ void __launch_bounds__(MC_THREADS) mc_signs(const float* __restrict__ f, const float* __restrict__ w, unsigned n,
                                                       float level, float min_w, unsigned* __restrict__ inside_bits,
                                                       unsigned* __restrict__ ok_bits) {
    float x[MC_PER_THREAD], y[MC_PER_THREAD];
#pragma unroll
    for (int it = 0; it < MC_PER_THREAD; it++) {
        const unsigned v = (blockIdx.x * MC_PER_THREAD + it) * MC_THREADS + threadIdx.x;
        x[it] = v < n ? __ldg(f + v) : level;              // not < level: a zero bit
        if (MASKED) y[it] = v < n ? __ldg(w + v) : min_w;
    }
#pragma unroll
    for (int it = 0; it < MC_PER_THREAD; it++) {
        const unsigned v = (blockIdx.x * MC_PER_THREAD + it) * MC_THREADS + threadIdx.x;
        const unsigned bi = __ballot_sync(0xffffffffu, x[it] < level);
        if ((threadIdx.x & 31) == 0) inside_bits[v >> 5] = bi;
        if (MASKED) {
            const unsigned bo = __ballot_sync(0xffffffffu, y[it] > min_w);
            if ((threadIdx.x & 31) == 0) ok_bits[v >> 5] = bo;
        }
    }
}

// bits idx .. idx+31 of a bit stream in `lo`, bit idx+32 in `hi` (the stream is padded with two zero words)
__device__ __forceinline__ void fetch33(const unsigned* __restrict__ bits, unsigned idx, unsigned n, unsigned& lo, unsigned& hi) {
    lo = hi = 0u;
    if (idx >= n) return;
    const unsigned wi = idx >> 5, sh = idx & 31u;
    const unsigned w0 = __ldg(bits + wi), w1 = __ldg(bits + wi + 1), w2 = __ldg(bits + wi + 2);
    lo = __funnelshift_r(w0, w1, sh);
    hi = __funnelshift_r(w1, w2, sh) & 1u;
}

// Case byte of every cell (0 = no triangles: outside the lattice, unobserved corner, or all corners on one side) and the
// triangle count of each 1024-voxel block, from the sign / observed bits.  A thread owns 32 consecutive voxels -- 33 bits
// from each of the lattice rows (y, z), (y+1, z), (y, z+1), (y+1, z+1) -- and a warp owns one 1024-voxel block.
template <bool MASKED>
__global__ void __launch_bounds__(MC_THREADS) mc_cases(const unsigned* __restrict__ inside_bits, const unsigned* __restrict__ ok_bits,
                                                       int nx, int ny, int nz, unsigned n, unsigned nb, FastDiv dx, FastDiv dy,
                                                       uint8_t* __restrict__ cases, unsigned* __restrict__ blk_tris) {
    __shared__ uint8_t s_ntri[256];
    s_ntri[threadIdx.x] = MC_NTRI[threadIdx.x];
    __syncthreads();
    const unsigned blk = (blockIdx.x * MC_THREADS + threadIdx.x) >> 5;            // warp-uniform
    if (blk >= nb) return;
    const unsigned nxy = (unsigned)nx * (unsigned)ny;
    const unsigned v = (blockIdx.x * MC_THREADS + threadIdx.x) * 32u;
    unsigned lo[4], hi[4];
    fetch33(inside_bits, v, n, lo[0], hi[0]);
    fetch33(inside_bits, v + nx, n, lo[1], hi[1]);
    fetch33(inside_bits, v + nxy, n, lo[2], hi[2]);
    fetch33(inside_bits, v + nxy + nx, n, lo[3], hi[3]);
    unsigned ok = 0xffffffffu;
    if (MASKED) {
        unsigned ol[4], oh[4];
        fetch33(ok_bits, v, n, ol[0], oh[0]);
        fetch33(ok_bits, v + nx, n, ol[1], oh[1]);
        fetch33(ok_bits, v + nxy, n, ol[2], oh[2]);
        fetch33(ok_bits, v + nxy + nx, n, ol[3], oh[3]);
        const unsigned o_lo = ol[0] & ol[1] & ol[2] & ol[3], o_hi = oh[0] & oh[1] & oh[2] & oh[3];
        ok = o_lo & ((o_lo >> 1) | (o_hi << 31));            // cell k needs columns k and k + 1
    }
    // cells with all eight corners inside the lattice
    unsigned in_lattice = 0;
    if (v < n) {
        const Voxel3 p = voxel_of(v, dx, dy);
        if (p.ix + 32 <= nx) {                                // one lattice row: only its last voxel has no +x neighbour
            in_lattice = (p.iy + 1 < ny && p.iz + 1 < nz) ? (p.ix + 32 == nx ? 0x7fffffffu : 0xffffffffu) : 0u;
        } else {                                              // the row wraps inside this thread's 32 voxels (nx % 32 != 0)
            int ix = p.ix, iy = p.iy, iz = p.iz;
            for (int k = 0; k < 32; k++) {
                if (ix + 1 < nx && iy + 1 < ny && iz + 1 < nz) in_lattice |= 1u << k;
                if (++ix == nx) {
                    ix = 0;
                    if (++iy == ny) { iy = 0; iz++; }
                }
            }
        }
    }
    ok &= in_lattice;
    unsigned tris = 0;
    unsigned out[8];
#pragma unroll
    for (int g = 0; g < 8; g++) {
        // five columns of the four rows byte-wise in one word; ((R >> k) & 0x03030303) * 0x01041040 >> 24 moves byte j's
        // two bits to bits 2j, 2j+1 of cell k's case (no two partial products meet: no carries)
        unsigned R = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned r5 = g < 7 ? (lo[j] >> (4 * g)) & 31u : (lo[j] >> 28) | (hi[j] << 4);
            R |= r5 << (8 * j);
        }
        unsigned packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            unsigned c = (((R >> k) & 0x03030303u) * 0x01041040u) >> 24;
            if (!((ok >> (4 * g + k)) & 1u)) c = 0;
            packed |= c << (8 * k);
            tris += s_ntri[c];
        }
        out[g] = packed;
    }
    uint4* dst = reinterpret_cast<uint4*>(cases + v);          // v is a multiple of 32; the buffer is padded to whole blocks
    dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
    dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
    tris = __reduce_add_sync(0xffffffffu, tris);
    if ((threadIdx.x & 31) == 0) blk_tris[blk] = tris;
}

__device__ __forceinline__ unsigned load_bytes4(const uint8_t* p, bool vec) {
    if (vec) return *reinterpret_cast<const unsigned*>(p);
    return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
}
// bit 0 of byte k = (bit k0 of case byte k) xor (bit k1 of case byte k), for the four case bytes of c4 at once
__device__ __forceinline__ unsigned differs4(unsigned c4, int k0, int k1) { return ((c4 >> k0) ^ (c4 >> k1)) & 0x01010101u; }

// 16 case bytes at p as four words (little endian: byte k of word i = cell 4 i + k)
__device__ __forceinline__ void load_cases16(const uint8_t* p, unsigned out[4], bool vec16, bool vec4) {
    if (vec16) {
        const uint4 q = *reinterpret_cast<const uint4*>(p);
        out[0] = q.x; out[1] = q.y; out[2] = q.z; out[3] = q.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) out[i] = load_bytes4(p + 4 * i, vec4);
    }
}

// 3-bit vertex mask of every voxel (does its +x / +y / +z lattice edge carry a vertex: is it crossed in one of the active
// cells around it) and the voxel's vertex rank inside its 1024-voxel block.  Reads below the lattice land in the zeroed
// guard, reads across a row / slice end land on the last cell of the previous row / slice, whose case is always 0.  A thread
// owns 16 consecutive voxels (128-bit loads, four voxels per 32-bit word, one byte each); 64 threads make a block, a CTA
// holds four of them.
constexpr int MC_EDGE_VOX = 16;
__global__ void __launch_bounds__(MC_THREADS) mc_edges(const uint8_t* __restrict__ cases, int nx, int ny, unsigned n, unsigned nb,
                                                       uint16_t* __restrict__ ecode, unsigned* __restrict__ blk_verts) {
    __shared__ unsigned s_warp[MC_THREADS / 32];
    const unsigned nxy = (unsigned)nx * (unsigned)ny;
    const bool vec16 = (nx & 15) == 0, vec4 = (nx & 3) == 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned blk = blockIdx.x * (MC_THREADS * MC_EDGE_VOX / MC_BLOCK_VOX) + (threadIdx.x >> 6);
    const unsigned v = (blockIdx.x * MC_THREADS + threadIdx.x) * MC_EDGE_VOX;
    unsigned m4[4] = {0u, 0u, 0u, 0u};                     // byte k of word i = the 3-bit mask of voxel 4 i + k
    if (v < n) {
        const uint8_t* p = cases + v;
        unsigned c[4], cy[4], cz[4], cyz[4];
        load_cases16(p, c, true, true);
        load_cases16(p - nx, cy, vec16, vec4);
        load_cases16(p - nxy, cz, vec16, vec4);
        load_cases16(p - nxy - nx, cyz, vec16, vec4);
        const unsigned bx = p[-1], bxy = p[-1 - (int)nx], bxz = *(p - 1 - nxy);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            // the same cells one step down in x: shift in the byte in front of each word
            const unsigned cx = (c[i] << 8) | (i ? c[i - 1] >> 24 : bx);
            const unsigned cxy = (cy[i] << 8) | (i ? cy[i - 1] >> 24 : bxy);
            const unsigned cxz = (cz[i] << 8) | (i ? cz[i - 1] >> 24 : bxz);
            // the +x edge is edge (0,1) of this cell, (2,3) of the cell below in y, (4,5) below in z, (6,7) below in both
            const unsigned ex = differs4(c[i], 0, 1) | differs4(cy[i], 2, 3) | differs4(cz[i], 4, 5) | differs4(cyz[i], 6, 7);
            const unsigned ey = differs4(c[i], 0, 2) | differs4(cx, 1, 3) | differs4(cz[i], 4, 6) | differs4(cxz, 5, 7);
            const unsigned ez = differs4(c[i], 0, 4) | differs4(cx, 1, 5) | differs4(cy[i], 2, 6) | differs4(cxy, 3, 7);
            m4[i] = ex | (ey << 1) | (ez << 2);
        }
    }
    const unsigned cnt = __popc(m4[0]) + __popc(m4[1]) + __popc(m4[2]) + __popc(m4[3]);
    const unsigned inc = warp_inclusive_scan(cnt);
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    const unsigned first = s_warp[wid & ~1], second = s_warp[wid | 1];      // the two warps of this thread's block
    unsigned rank = inc - cnt + ((wid & 1) ? first : 0u);
    if (v < n) {
        unsigned out[8];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const unsigned m = m4[i];
            const unsigned r1 = rank + __popc(m & 0xffu), r2 = rank + __popc(m & 0xffffu), r3 = rank + __popc(m & 0xffffffu);
            out[2 * i] = ((rank << 3) | (m & 7u)) | (((r1 << 3) | ((m >> 8) & 7u)) << 16);
            out[2 * i + 1] = ((r2 << 3) | ((m >> 16) & 7u)) | (((r3 << 3) | (m >> 24)) << 16);
            rank += __popc(m);
        }
        uint4* dst = reinterpret_cast<uint4*>(ecode + v);
        dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
        dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
    }
    if ((threadIdx.x & 63) == 0 && blk < nb) blk_verts[blk] = first + second;
}

// The 1024-voxel blocks that own a vertex or a triangle, in any order (their outputs are placed by the scanned bases).
__global__ void __launch_bounds__(256) mc_list_blocks(const unsigned* __restrict__ tri_base, const unsigned* __restrict__ vert_base,
                                                      unsigned nb, unsigned* __restrict__ list, unsigned* __restrict__ count) {
    const unsigned b = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = b < nb && (tri_base[b + 1] != tri_base[b] || vert_base[b + 1] != vert_base[b]);
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (on) list[base + __popc(m & ((1u << lane) - 1u))] = b;
}

// Vertices and triangles of the listed blocks.  Small CTAs (64 threads, 16 voxels each = one block per turn) so that many
// blocks are in flight per SM: a block's work is a chain of dependent round trips (case / code words -> scan -> the corner
// owners' codes -> stores), and the next block's words are requested before the current one is processed.  Inside a block the
// vertices and the triangles are two DENSE lists: item i is taken by thread i mod 64, which finds its voxel by binary search
// over the ranks in shared memory (the surface touches ~1 % of the voxels; a thread-per-voxel walk left one active lane per
// warp and cost three times the instructions).
constexpr int MC_EMIT_THREADS = 64;
constexpr int MC_EMIT_VOX = MC_BLOCK_VOX / MC_EMIT_THREADS;      // 16
template <bool COLOR>
__global__ void __launch_bounds__(MC_EMIT_THREADS) mc_emit(const float* __restrict__ f, const float* __restrict__ rgb,
                                                           const uint8_t* __restrict__ cases, const uint16_t* __restrict__ ecode,
                                                           const unsigned* __restrict__ tri_base,
                                                           const unsigned* __restrict__ vert_base,
                                                           const unsigned* __restrict__ list, const unsigned* __restrict__ count_ptr,
                                                           int nx, int ny, unsigned n, FastDiv dx, FastDiv dy, float level,
                                                           float ox, float oy, float oz, float voxel, float* __restrict__ verts,
                                                           float* __restrict__ colors, int* __restrict__ faces) {
    __shared__ uint8_t s_ntri[256];
    __shared__ uint64_t s_tris[256];
    __shared__ __align__(16) unsigned s_c4[MC_BLOCK_VOX / 4];                // per group of four voxels: the case bytes ...
    __shared__ unsigned s_pre[MC_BLOCK_VOX / 4];                             // ... and the block's triangles before the group
    __shared__ __align__(16) unsigned short s_code[MC_BLOCK_VOX];
    __shared__ unsigned s_warp[2], s_base[4];
    const unsigned count = *count_ptr;
    if (blockIdx.x >= count) return;
#pragma unroll
    for (int k = 0; k < 256 / MC_EMIT_THREADS; k++) {
        s_ntri[threadIdx.x + k * MC_EMIT_THREADS] = MC_NTRI[threadIdx.x + k * MC_EMIT_THREADS];
        s_tris[threadIdx.x + k * MC_EMIT_THREADS] = MC_TRIS[threadIdx.x + k * MC_EMIT_THREADS];
    }
    const unsigned nxy = (unsigned)nx * (unsigned)ny;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // the words of block `turn` travel in registers from the end of the previous turn
    uint4 c16 = make_uint4(0u, 0u, 0u, 0u), e0 = c16, e1 = c16;
    unsigned base_word = 0, blk = 0;
    auto request = [&](unsigned turn) {
        blk = __ldg(list + turn);
        const unsigned v = blk * MC_BLOCK_VOX + threadIdx.x * MC_EMIT_VOX;
        c16 = e0 = e1 = make_uint4(0u, 0u, 0u, 0u);
        if (v < n) {
            c16 = *reinterpret_cast<const uint4*>(cases + v);
            e0 = *reinterpret_cast<const uint4*>(ecode + v);
            e1 = *reinterpret_cast<const uint4*>(ecode + v + 8);
        }
        if (threadIdx.x < 4) base_word = (threadIdx.x & 1) ? vert_base[blk + (threadIdx.x >> 1)] : tri_base[blk + (threadIdx.x >> 1)];
    };
    request(blockIdx.x);
    __syncthreads();                                           // the tables
    for (unsigned turn = blockIdx.x; turn < count; turn += gridDim.x) {
        const unsigned vox0 = blk * MC_BLOCK_VOX;
        // ---- phase 1: words to shared memory, triangles before each group of four voxels
        const unsigned cw[4] = {c16.x, c16.y, c16.z, c16.w};
        unsigned gcount[4], mine = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) {
            gcount[g] = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) gcount[g] += s_ntri[(cw[g] >> (8 * k)) & 255u];
            mine += gcount[g];
        }
        const unsigned inc = warp_inclusive_scan(mine);
        if (lane == 31) s_warp[wid] = inc;
        reinterpret_cast<uint4*>(s_c4)[threadIdx.x] = c16;
        reinterpret_cast<uint4*>(s_code)[2 * threadIdx.x] = e0;
        reinterpret_cast<uint4*>(s_code)[2 * threadIdx.x + 1] = e1;
        if (threadIdx.x < 4) s_base[threadIdx.x] = base_word;          // t0, v0, t1, v1
        __syncthreads();
        unsigned before = inc - mine + (wid ? s_warp[0] : 0u);
#pragma unroll
        for (int g = 0; g < 4; g++) {
            s_pre[4 * threadIdx.x + g] = before;
            before += gcount[g];
        }
        const unsigned t0 = s_base[0], v0 = s_base[1], ntri = s_base[2] - t0, nvert = s_base[3] - v0;
        __syncthreads();
        if (turn + gridDim.x < count) request(turn + gridDim.x);       // in flight during phase 2

        // ---- phase 2, vertices: local rank r -> the voxel whose rank range holds r (ranks are non-decreasing over the voxels)
        // voxels of this block that mc_edges wrote codes for (whole threads): the ranks are monotone over these only
        const unsigned nvox = min((unsigned)MC_BLOCK_VOX, (n - vox0 + MC_EMIT_VOX - 1) & ~(unsigned)(MC_EMIT_VOX - 1));
        for (unsigned r = threadIdx.x; r < nvert; r += MC_EMIT_THREADS) {
            unsigned lo = 0, hi = nvox - 1;                    // largest voxel with rank <= r
            while (lo < hi) {
                const unsigned mid = (lo + hi + 1) >> 1;
                if ((unsigned)(s_code[mid] >> 3) <= r) lo = mid; else hi = mid - 1;
            }
            const unsigned code = s_code[lo];
            const int a = (int)__fns(code & 7u, 0, (int)(r - (code >> 3)) + 1);     // the (r - rank)-th set bit of the mask
            const unsigned v = vox0 + lo;
            const Voxel3 p = voxel_of(v, dx, dy);
            const unsigned vn = v + (a == 0 ? 1u : (a == 1 ? (unsigned)nx : nxy));
            const float f0 = __fsub_rn(__ldg(f + v), level), f1 = __fsub_rn(__ldg(f + vn), level);
            const float t = __fdiv_rn(f0, __fsub_rn(f0, f1));
            const float gx = (float)p.ix, gy = (float)p.iy, gz = (float)p.iz;
            const size_t dst = 3 * (size_t)(v0 + r);
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
        // ---- phase 2, triangles: local index i -> the group of four voxels whose prefix range holds i -> the voxel
        for (unsigned i = threadIdx.x; i < ntri; i += MC_EMIT_THREADS) {
            unsigned lo = 0, hi = MC_BLOCK_VOX / 4 - 1;        // largest group with prefix <= i
            while (lo < hi) {
                const unsigned mid = (lo + hi + 1) >> 1;
                if (s_pre[mid] <= i) lo = mid; else hi = mid - 1;
            }
            unsigned r = i - s_pre[lo];
            const unsigned c4 = s_c4[lo];
            unsigned k = 0, c = c4 & 255u;
            while (r >= s_ntri[c]) {
                r -= s_ntri[c];
                k++;
                c = (c4 >> (8 * k)) & 255u;
            }
            const unsigned v = vox0 + lo * 4 + k;
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
            const size_t dst = 3 * (size_t)(t0 + i);
            faces[dst + 0] = id[0];
            faces[dst + 1] = id[1];
            faces[dst + 2] = id[2];
        }
        __syncthreads();                                       // the next turn overwrites the shared words
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
    GSR_CUDA_CHECK(cudaMemsetAsync(w.inside_bits + w.bit_words - 2, 0, 8, s));
    GSR_CUDA_CHECK(cudaMemsetAsync(w.ok_bits + w.bit_words - 2, 0, 8, s));
    const unsigned case_ctas = (nb * 32 + MC_THREADS - 1) / MC_THREADS;          // one warp per 1024-voxel block
    if (weight) {
        mc_signs<true><<<nb, MC_THREADS, 0, s>>>(tsdf, weight, (unsigned)n, level, min_weight, w.inside_bits, w.ok_bits);
        mc_cases<true><<<case_ctas, MC_THREADS, 0, s>>>(w.inside_bits, w.ok_bits, nx, ny, nz, (unsigned)n, nb, dx, dy, w.cases, w.tri_base);
    } else {
        mc_signs<false><<<nb, MC_THREADS, 0, s>>>(tsdf, nullptr, (unsigned)n, level, 0.f, w.inside_bits, nullptr);
        mc_cases<false><<<case_ctas, MC_THREADS, 0, s>>>(w.inside_bits, nullptr, nx, ny, nz, (unsigned)n, nb, dx, dy, w.cases, w.tri_base);
    }
    mc_edges<<<(nb + 3) / 4, MC_THREADS, 0, s>>>(w.cases, nx, ny, (unsigned)n, nb, w.ecode, w.vert_base);
    const unsigned chunks = (nb + SCAN_CHUNK - 1) / SCAN_CHUNK;
    scan_local<true><<<chunks, 1024, 0, s>>>(w.tri_base, w.vert_base, nb, w.chunk_sums);
    scan_add<true><<<chunks, 1024, 0, s>>>(w.tri_base, w.vert_base, nb, w.chunk_sums, w.totals);
    GSR_CUDA_CHECK(cudaMemsetAsync(w.totals + 2, 0, 4, s));
    mc_list_blocks<<<(nb + 255) / 256, 256, 0, s>>>(w.tri_base, w.vert_base, nb, w.block_list, w.totals + 2);
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
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)std::min<long long>((long long)nb, (long long)sms * 24);     // persistent: 24 small CTAs per SM
    if (rgb)
        mc_emit<true><<<grid, MC_EMIT_THREADS, 0, s>>>(tsdf, rgb, w.cases, w.ecode, w.tri_base, w.vert_base, w.block_list, w.totals + 2, nx, ny,
                                                       (unsigned)n, dx, dy, level, origin[0], origin[1], origin[2], voxel_size, verts, colors, faces);
    else
        mc_emit<false><<<grid, MC_EMIT_THREADS, 0, s>>>(tsdf, nullptr, w.cases, w.ecode, w.tri_base, w.vert_base, w.block_list, w.totals + 2, nx, ny,
                                                        (unsigned)n, dx, dy, level, origin[0], origin[1], origin[2], voxel_size, verts, nullptr, faces);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}
