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
//             triangles (12 B).  A triangle corner's global vertex id is block_base[owner >> 10] + rank(owner) +
//             popc(mask below the axis): no per-voxel 4-byte id array, no atomics, output order = lattice order
//             (deterministic, reproducible by the oracle).
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
                                                       int nz, unsigned n, unsigned nb, float level, float min_w,
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
            const unsigned row = v / (unsigned)nx;
            int ix = (int)(v - row * (unsigned)nx), iy = (int)(row % (unsigned)ny), iz = (int)(row / (unsigned)ny);
            unsigned packed = 0;
#pragma unroll
            for (int k = 0; k < MC_PER_THREAD; k++) {
                unsigned c = ((r[0] >> k) & 3u) | (((r[1] >> k) & 3u) << 2) | (((r[2] >> k) & 3u) << 4) | (((r[3] >> k) & 3u) << 6);
                bool ok = ix + 1 < nx && iy + 1 < ny && iz + 1 < nz;      // (voxels at or beyond n have iz >= nz)
                if (MASKED) ok = ok && (((o[0] & o[1] & o[2] & o[3]) >> k) & 3u) == 3u;
                if (!ok) c = 0;
                packed |= c << (8 * k);
                tris += s_ntri[c];
                if (++ix == nx) {
                    ix = 0;
                    if (++iy == ny) { iy = 0; iz++; }
                }
            }
            *reinterpret_cast<unsigned*>(cases + v) = packed;
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
// bit k0 xor bit k1 of each of the four case bytes, gathered into bits 0..3
__device__ __forceinline__ unsigned differs4(unsigned c4, int k0, int k1) {
    const unsigned d = ((c4 >> k0) ^ (c4 >> k1)) & 0x01010101u;
    return (d | (d >> 7) | (d >> 14) | (d >> 21)) & 15u;
}

// 3-bit vertex mask of every voxel (does its +x / +y / +z lattice edge carry a vertex: is it crossed in one of the active
// cells around it) and the voxel's vertex rank inside its block.  Reads below the lattice land in the zeroed guard, reads
// across a row / slice end land on the last cell of the previous row / slice, whose case is always 0.
__global__ void __launch_bounds__(MC_THREADS) mc_edges(const uint8_t* __restrict__ cases, int nx, int ny, unsigned n, unsigned nb,
                                                       uint16_t* __restrict__ ecode, unsigned* __restrict__ blk_verts) {
    __shared__ unsigned s_warp[MC_EDGES_SUB][MC_THREADS / 32];
    const unsigned nxy = (unsigned)nx * (unsigned)ny;
    const bool vec = (nx & 3) == 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned mx[MC_EDGES_SUB], my[MC_EDGES_SUB], mz[MC_EDGES_SUB], inc[MC_EDGES_SUB];
#pragma unroll
    for (int sub = 0; sub < MC_EDGES_SUB; sub++) {          // independent 1024-voxel blocks: their loads overlap
        const unsigned v = ((blockIdx.x * MC_EDGES_SUB + sub) * MC_THREADS + threadIdx.x) * MC_PER_THREAD;
        mx[sub] = my[sub] = mz[sub] = 0;
        if (v < n) {
            const uint8_t* p = cases + v;
            const unsigned c = load_bytes4(p, true);
            const unsigned cy = load_bytes4(p - nx, vec), cz = load_bytes4(p - nxy, vec), cyz = load_bytes4(p - nxy - nx, vec);
            // the same four cells one step down in x: shift in the byte in front of each group
            const unsigned cx = (c << 8) | p[-1];
            const unsigned cxy = (cy << 8) | p[-1 - (int)nx];
            const unsigned cxz = (cz << 8) | *(p - 1 - nxy);
            // the +x edge is edge (0,1) of this cell, (2,3) of the cell below in y, (4,5) below in z, (6,7) below in both
            mx[sub] = differs4(c, 0, 1) | differs4(cy, 2, 3) | differs4(cz, 4, 5) | differs4(cyz, 6, 7);
            my[sub] = differs4(c, 0, 2) | differs4(cx, 1, 3) | differs4(cz, 4, 6) | differs4(cxz, 5, 7);
            mz[sub] = differs4(c, 0, 4) | differs4(cx, 1, 5) | differs4(cy, 2, 6) | differs4(cxy, 3, 7);
        }
    }
#pragma unroll
    for (int sub = 0; sub < MC_EDGES_SUB; sub++) {
        inc[sub] = warp_inclusive_scan(__popc(mx[sub]) + __popc(my[sub]) + __popc(mz[sub]));
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
        unsigned rank = before + inc[sub] - (__popc(mx[sub]) + __popc(my[sub]) + __popc(mz[sub]));
        if (v < n) {
            unsigned short out[MC_PER_THREAD];
#pragma unroll
            for (int k = 0; k < MC_PER_THREAD; k++) {
                const unsigned m = ((mx[sub] >> k) & 1u) | (((my[sub] >> k) & 1u) << 1) | (((mz[sub] >> k) & 1u) << 2);
                out[k] = (unsigned short)((rank << 3) | m);
                rank += __popc(m);
            }
            *reinterpret_cast<uint2*>(ecode + v) = make_uint2(out[0] | ((unsigned)out[1] << 16), out[2] | ((unsigned)out[3] << 16));
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
                                                      unsigned nb, float level, float ox, float oy, float oz, float voxel,
                                                      float* __restrict__ verts, float* __restrict__ colors,
                                                      int* __restrict__ faces) {
    __shared__ unsigned s_warp[2][MC_THREADS / 32];
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
    const unsigned stride[3] = {1u, (unsigned)nx, nxy};
    int turn = 0;                                     // alternates the scan scratch between consecutive non-empty blocks
    for (int sub = 0; sub < MC_EMIT_SUB; sub++) {
    const unsigned t0 = s_tb[sub], v0 = s_vb[sub];
    if (t0 == s_tb[sub + 1] && v0 == s_vb[sub + 1]) continue;          // block-uniform
    const unsigned b = blockIdx.x * MC_EMIT_SUB + sub;
    const unsigned vfirst = (b * MC_THREADS + threadIdx.x) * MC_PER_THREAD;
    unsigned c4 = 0;
    uint2 e4 = make_uint2(0u, 0u);
    if (vfirst < n) {
        c4 = *reinterpret_cast<const unsigned*>(cases + vfirst);
        e4 = *reinterpret_cast<const uint2*>(ecode + vfirst);
    }
    unsigned nt_thread = 0;
#pragma unroll
    for (int k = 0; k < MC_PER_THREAD; k++) nt_thread += s_ntri[(c4 >> (8 * k)) & 255u];
    unsigned total;
    unsigned out = t0 + block_exclusive_scan(nt_thread, s_warp[(turn++) & 1], &total);
    if (!(e4.x | e4.y) && !nt_thread) continue;
    const unsigned row = vfirst / (unsigned)nx;
    int ix = (int)(vfirst - row * (unsigned)nx), iy = (int)(row % (unsigned)ny), iz = (int)(row / (unsigned)ny);
#pragma unroll
    for (int k = 0; k < MC_PER_THREAD; k++) {
        const unsigned v = vfirst + k;
        const unsigned c = (c4 >> (8 * k)) & 255u;
        const unsigned code = ((k < 2 ? e4.x : e4.y) >> (16 * (k & 1))) & 0xffffu;
        // ---- vertices on the voxel's own +x / +y / +z edges
        if (code & 7u) {
            const float g[3] = {(float)ix, (float)iy, (float)iz};
            const float org[3] = {ox, oy, oz};
            const float f0 = __fsub_rn(__ldg(f + v), level);
            unsigned dst = v0 + (code >> 3);
#pragma unroll
            for (int a = 0; a < 3; a++) {
                if (!((code >> a) & 1u)) continue;
                const unsigned vn = v + stride[a];
                const float f1 = __fsub_rn(__ldg(f + vn), level);
                const float t = __fdiv_rn(f0, __fsub_rn(f0, f1));
#pragma unroll
                for (int ax = 0; ax < 3; ax++) {
                    const float coord = ax == a ? __fadd_rn(g[ax], t) : g[ax];
                    verts[3 * (size_t)dst + ax] = __fadd_rn(org[ax], __fmul_rn(coord, voxel));
                }
                if (COLOR) {
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        const float c0 = __ldg(rgb + 3 * (size_t)v + ch), c1 = __ldg(rgb + 3 * (size_t)vn + ch);
                        colors[3 * (size_t)dst + ch] = __fadd_rn(c0, __fmul_rn(t, __fsub_rn(c1, c0)));
                    }
                }
                dst++;
            }
        }
        // ---- triangles of the cell
        const unsigned nt = s_ntri[c];
        if (nt) {
            const uint64_t word = s_tris[c];
            for (unsigned t = 0; t < nt; t++) {
                int id[3];
#pragma unroll
                for (int q = 0; q < 3; q++) {
                    const unsigned e = (unsigned)(word >> (12 * t + 4 * q)) & 15u;
                    const unsigned a = e >> 2, j = e & 3u;
                    // the edge's owner: this voxel moved along the other two axes (in increasing order) by the bits of j
                    const unsigned su = a == 0 ? (unsigned)nx : 1u, sv = a == 2 ? (unsigned)nx : nxy;
                    const unsigned owner = v + (j & 1u) * su + (j >> 1) * sv;
                    const unsigned oc = ecode[owner];
                    id[q] = (int)(vert_base[owner >> MC_BLOCK_SHIFT] + (oc >> 3) + __popc(oc & ((1u << a) - 1u)));
                }
                faces[3 * (size_t)out + 0] = id[0];
                faces[3 * (size_t)out + 1] = id[1];
                faces[3 * (size_t)out + 2] = id[2];
                out++;
            }
        }
        if (++ix == nx) {
            ix = 0;
            if (++iy == ny) { iy = 0; iz++; }
        }
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
    GSR_CUDA_CHECK(cudaMemsetAsync(w.guard, 0, w.guard_bytes, s));
    if (weight)
        mc_cases<true><<<(nb + MC_CASES_SUB - 1) / MC_CASES_SUB, MC_THREADS, 0, s>>>(tsdf, weight, nx, ny, nz, (unsigned)n, nb, level, min_weight,
                                                                                  w.cases, w.tri_base);
    else
        mc_cases<false><<<(nb + MC_CASES_SUB - 1) / MC_CASES_SUB, MC_THREADS, 0, s>>>(tsdf, nullptr, nx, ny, nz, (unsigned)n, nb, level, 0.f,
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
    if (rgb)
        mc_emit<true><<<(nb + MC_EMIT_SUB - 1) / MC_EMIT_SUB, MC_THREADS, 0, s>>>(tsdf, rgb, w.cases, w.ecode, w.tri_base, w.vert_base, nx, ny, (unsigned)n, nb, level,
                                                origin[0], origin[1], origin[2], voxel_size, verts, colors, faces);
    else
        mc_emit<false><<<(nb + MC_EMIT_SUB - 1) / MC_EMIT_SUB, MC_THREADS, 0, s>>>(tsdf, nullptr, w.cases, w.ecode, w.tri_base, w.vert_base, nx, ny, (unsigned)n, nb, level,
                                                 origin[0], origin[1], origin[2], voxel_size, verts, nullptr, faces);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}
