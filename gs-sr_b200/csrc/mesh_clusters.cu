// mesh_clusters.cu -- connected clusters of a triangle mesh and the order-preserving removal of triangles / unreferenced
// vertices, for sm_100a: the device side of GS-SR's post_process_mesh (/root/reference/gssr/utils/mesh_utils.py:27-49),
// which the reference runs through Open3D on the host:
//     triangle_clusters, cluster_n_triangles, cluster_area = mesh.cluster_connected_triangles()
//     mesh.remove_triangles_by_mask(cluster_n_triangles[triangle_clusters] < n_cluster)
//     mesh.remove_unreferenced_vertices();  mesh.remove_degenerate_triangles()
// Open3D is an absent third-party dependency (parity unpinned); oracle/mesh_clusters_oracle.py restates the steps with
// scipy's connected components and the kernels reproduce it exactly (integer results; cluster areas to rounding).
//
// Clusters: lock-free union-find over the VERTICES (hooking the larger root under the smaller with atomicCAS, pointer
// jumping while searching -- the ECL-CC scheme), one thread per triangle; a triangle's cluster is the component of its
// vertices, labelled by the component's smallest vertex id.  Open3D joins triangles through shared edges; the two
// notions differ only where two sheets touch in a single vertex, which a marching-cubes surface has at most at the rim of
// the observed region.
// Removal: flags -> counts per 1024-element block -> scan_util.cuh scans -> ranks; survivors keep their order.
#include "common.cuh"
#include "scan_util.cuh"
#include "../../include/gsr_b200.h"

namespace gsr {

constexpr int ML_THREADS = 256;
constexpr int ML_PER_THREAD = 4;
constexpr int ML_BLOCK = ML_THREADS * ML_PER_THREAD;

__device__ __forceinline__ unsigned cc_find(unsigned* parent, unsigned x) {
    unsigned curr = parent[x];
    if (curr != x) {
        unsigned prev = x, next;
        while (curr > (next = parent[curr])) {      // parent[i] <= i always: roots are the smallest ids
            parent[prev] = next;
            prev = curr;
            curr = next;
        }
    }
    return curr;
}

__device__ __forceinline__ void cc_unite(unsigned* parent, unsigned a, unsigned b) {
    unsigned ra = cc_find(parent, a), rb = cc_find(parent, b);
    while (ra != rb) {
        if (ra < rb) {
            const unsigned old = atomicCAS(parent + rb, rb, ra);
            if (old == rb) break;
            rb = old;
        } else {
            const unsigned old = atomicCAS(parent + ra, ra, rb);
            if (old == ra) break;
            ra = old;
        }
    }
}

__global__ void cc_init(unsigned* __restrict__ parent, unsigned nv) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nv) parent[i] = i;
}

__global__ void cc_hook(const int* __restrict__ faces, unsigned nf, unsigned nv, unsigned* parent, unsigned* bad) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nf) return;
    const unsigned a = (unsigned)faces[3 * (size_t)t], b = (unsigned)faces[3 * (size_t)t + 1], c = (unsigned)faces[3 * (size_t)t + 2];
    if (a >= nv || b >= nv || c >= nv) {
        atomicAdd(bad, 1u);
        return;
    }
    cc_unite(parent, a, b);
    cc_unite(parent, a, c);
}

__global__ void cc_flatten(unsigned* parent, unsigned nv) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    unsigned r = parent[i];
    while (r != parent[r]) r = parent[r];
    parent[i] = r;                                  // only ever lowers a pointer towards its root: safe against the other threads
}

// per-triangle root, triangles and area per root
__global__ void cc_count(const int* __restrict__ faces, const float* __restrict__ verts, unsigned nf, unsigned nv,
                         const unsigned* __restrict__ parent, int* __restrict__ tri_root, unsigned* __restrict__ root_ntris,
                         double* __restrict__ root_area) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nf) return;
    const unsigned a = (unsigned)faces[3 * (size_t)t], b = (unsigned)faces[3 * (size_t)t + 1], c = (unsigned)faces[3 * (size_t)t + 2];
    if (a >= nv || b >= nv || c >= nv) {
        tri_root[t] = -1;
        return;
    }
    const unsigned r = parent[a];
    tri_root[t] = (int)r;
    atomicAdd(root_ntris + r, 1u);
    if (verts) {
        const double ax = verts[3 * (size_t)a], ay = verts[3 * (size_t)a + 1], az = verts[3 * (size_t)a + 2];
        const double ux = verts[3 * (size_t)b] - ax, uy = verts[3 * (size_t)b + 1] - ay, uz = verts[3 * (size_t)b + 2] - az;
        const double vx = verts[3 * (size_t)c] - ax, vy = verts[3 * (size_t)c + 1] - ay, vz = verts[3 * (size_t)c + 2] - az;
        const double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
        atomicAdd(root_area + r, 0.5 * sqrt(nx * nx + ny * ny + nz * nz));
    }
}

// flags (one byte per element, non-zero = set) -> number of set flags per 1024-element block
__global__ void __launch_bounds__(ML_THREADS) flag_block_counts(const uint8_t* __restrict__ flags, unsigned n,
                                                                unsigned* __restrict__ blk) {
    __shared__ unsigned s_warp[ML_THREADS / 32];
    const unsigned i0 = (blockIdx.x * ML_THREADS + threadIdx.x) * ML_PER_THREAD;
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < ML_PER_THREAD; k++) c += (i0 + k < n && flags[i0 + k]) ? 1u : 0u;
    unsigned total;
    block_exclusive_scan(c, s_warp, &total);
    if (threadIdx.x == 0) blk[blockIdx.x] = total;
}

// kept triangles mark their vertices (before the degenerate ones are dropped, like the reference's order of calls) and
// tri_flag becomes "kept and not degenerate"
__global__ void filter_mark(const int* __restrict__ faces, unsigned nf, unsigned nv, const uint8_t* __restrict__ keep,
                            uint8_t* __restrict__ vert_flag, uint8_t* __restrict__ tri_flag) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nf) return;
    const unsigned a = (unsigned)faces[3 * (size_t)t], b = (unsigned)faces[3 * (size_t)t + 1], c = (unsigned)faces[3 * (size_t)t + 2];
    uint8_t out = 0;
    if (keep[t] && a < nv && b < nv && c < nv) {
        vert_flag[a] = 1;
        vert_flag[b] = 1;
        vert_flag[c] = 1;
        out = (a != b && b != c && a != c) ? 1 : 0;
    }
    tri_flag[t] = out;
}

// surviving vertices move to their rank; vert_new[v] = new index (undefined for dropped vertices)
__global__ void __launch_bounds__(ML_THREADS) filter_emit_verts(const float* __restrict__ verts, const float* __restrict__ colors,
                                                                const uint8_t* __restrict__ vert_flag, unsigned nv,
                                                                const unsigned* __restrict__ blk_base, int* __restrict__ vert_new,
                                                                float* __restrict__ verts_out, float* __restrict__ colors_out) {
    __shared__ unsigned s_warp[ML_THREADS / 32];
    const unsigned i0 = (blockIdx.x * ML_THREADS + threadIdx.x) * ML_PER_THREAD;
    bool on[ML_PER_THREAD];
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < ML_PER_THREAD; k++) {
        on[k] = i0 + k < nv && vert_flag[i0 + k];
        c += on[k] ? 1u : 0u;
    }
    unsigned total;
    unsigned dst = blk_base[blockIdx.x] + block_exclusive_scan(c, s_warp, &total);
#pragma unroll
    for (int k = 0; k < ML_PER_THREAD; k++) {
        if (!on[k]) continue;
        const size_t src = i0 + k;
        vert_new[src] = (int)dst;
#pragma unroll
        for (int q = 0; q < 3; q++) verts_out[3 * (size_t)dst + q] = verts[3 * src + q];
        if (colors)
#pragma unroll
            for (int q = 0; q < 3; q++) colors_out[3 * (size_t)dst + q] = colors[3 * src + q];
        dst++;
    }
}

__global__ void __launch_bounds__(ML_THREADS) filter_emit_tris(const int* __restrict__ faces, const uint8_t* __restrict__ tri_flag,
                                                               unsigned nf, const unsigned* __restrict__ blk_base,
                                                               const int* __restrict__ vert_new, int* __restrict__ faces_out) {
    __shared__ unsigned s_warp[ML_THREADS / 32];
    const unsigned i0 = (blockIdx.x * ML_THREADS + threadIdx.x) * ML_PER_THREAD;
    bool on[ML_PER_THREAD];
    unsigned c = 0;
#pragma unroll
    for (int k = 0; k < ML_PER_THREAD; k++) {
        on[k] = i0 + k < nf && tri_flag[i0 + k];
        c += on[k] ? 1u : 0u;
    }
    unsigned total;
    unsigned dst = blk_base[blockIdx.x] + block_exclusive_scan(c, s_warp, &total);
#pragma unroll
    for (int k = 0; k < ML_PER_THREAD; k++) {
        if (!on[k]) continue;
        const size_t src = i0 + k;
#pragma unroll
        for (int q = 0; q < 3; q++) faces_out[3 * (size_t)dst + q] = vert_new[faces[3 * src + q]];
        dst++;
    }
}

__global__ void keep_by_cluster_size(const int* __restrict__ tri_root, const unsigned* __restrict__ root_ntris, unsigned nf,
                                     unsigned min_triangles, uint8_t* __restrict__ keep) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nf) return;
    const int r = tri_root[t];
    keep[t] = (r >= 0 && root_ntris[r] >= min_triangles) ? 1 : 0;
}

// sizes of the clusters (the non-zero entries of root_ntris), in any order, appended behind a warp-aggregated counter
__global__ void collect_cluster_sizes(const unsigned* __restrict__ root_ntris, unsigned nv, unsigned* __restrict__ sizes,
                                      unsigned cap, unsigned* __restrict__ count) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned x = i < nv ? root_ntris[i] : 0u;
    const unsigned m = __ballot_sync(0xffffffffu, x != 0u);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned pos = base + __popc(m & ((1u << lane) - 1u));
    if (x != 0u && pos < cap) sizes[pos] = x;
}

struct FilterWorkspace {
    uint8_t* vert_flag;    // nv
    uint8_t* tri_flag;     // nf
    int* vert_new;         // nv
    unsigned* vert_base;   // vblocks + 1
    unsigned* tri_base;    // tblocks + 1
    unsigned long long* chunk_sums;
    unsigned* totals;      // 4
    size_t bytes;
};

static size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

static FilterWorkspace filter_layout(void* base, long long nv, long long nf) {
    const long long vb = (nv + ML_BLOCK - 1) / ML_BLOCK, tb = (nf + ML_BLOCK - 1) / ML_BLOCK;
    const long long chunks = ((vb > tb ? vb : tb) + SCAN_CHUNK - 1) / SCAN_CHUNK + 1;
    FilterWorkspace w;
    char* b = (char*)base;
    size_t off = 0;
    w.vert_flag = (uint8_t*)(b + off);  off += up256((size_t)nv);
    w.tri_flag = (uint8_t*)(b + off);   off += up256((size_t)nf);
    w.vert_new = (int*)(b + off);       off += up256((size_t)nv * 4);
    w.vert_base = (unsigned*)(b + off); off += up256((size_t)(vb + 1) * 4);
    w.tri_base = (unsigned*)(b + off);  off += up256((size_t)(tb + 1) * 4);
    w.chunk_sums = (unsigned long long*)(b + off); off += up256((size_t)chunks * 16);
    w.totals = (unsigned*)(b + off);    off += 256;
    w.bytes = off;
    return w;
}

static int check_sizes(const char* who, long long nv, long long nf) {
    if (nv < 0 || nf < 0 || nv > 0x7fffffffLL - ML_BLOCK || nf > 0x7fffffffLL - ML_BLOCK) {
        set_error("%s: invalid mesh size (%lld vertices, %lld triangles)", who, nv, nf);
        return GSR_E_INVALID;
    }
    return GSR_OK;
}

}  // namespace gsr

extern "C" int gsr_mesh_clusters(long long nverts, long long ntris, const float* verts, const int* faces, int* vertex_root,
                                 int* tri_root, unsigned int* root_ntris, double* root_area, void* stream_v) {
    using namespace gsr;
    cudaStream_t s = (cudaStream_t)stream_v;
    if (int rc = check_sizes("gsr_mesh_clusters", nverts, ntris)) return rc;
    if ((ntris > 0 && (!faces || !tri_root)) || (nverts > 0 && (!vertex_root || !root_ntris)) || ((verts != nullptr) != (root_area != nullptr))) {
        set_error("gsr_mesh_clusters: invalid argument");
        return GSR_E_INVALID;
    }
    if (nverts == 0 && ntris == 0) return GSR_OK;
    const unsigned nv = (unsigned)nverts, nf = (unsigned)ntris;
    unsigned* parent = reinterpret_cast<unsigned*>(vertex_root);
    if (nv) {
        cc_init<<<(nv + 255) / 256, 256, 0, s>>>(parent, nv);
        GSR_CUDA_CHECK(cudaMemsetAsync(root_ntris, 0, (size_t)nv * 4, s));
        if (root_area) GSR_CUDA_CHECK(cudaMemsetAsync(root_area, 0, (size_t)nv * 8, s));
    }
    if (nf) {
        // the "bad index" counter borrows tri_root[0] until cc_count overwrites it
        GSR_CUDA_CHECK(cudaMemsetAsync(tri_root, 0, 4, s));
        cc_hook<<<(nf + 255) / 256, 256, 0, s>>>(faces, nf, nv, parent, reinterpret_cast<unsigned*>(tri_root));
    }
    unsigned bad = 0;
    if (nf) GSR_CUDA_CHECK(cudaMemcpyAsync(&bad, tri_root, 4, cudaMemcpyDeviceToHost, s));
    if (nv) cc_flatten<<<(nv + 255) / 256, 256, 0, s>>>(parent, nv);
    if (nf) cc_count<<<(nf + 255) / 256, 256, 0, s>>>(faces, verts, nf, nv, parent, tri_root, root_ntris, root_area);
    GSR_CUDA_CHECK(cudaGetLastError());
    GSR_CUDA_CHECK(cudaStreamSynchronize(s));
    if (bad) {
        set_error("gsr_mesh_clusters: %u triangles index vertices outside [0, %lld)", bad, nverts);
        return GSR_E_INVALID;
    }
    return GSR_OK;
}

extern "C" int gsr_mesh_cluster_sizes(long long nverts, const unsigned int* root_ntris, unsigned int* sizes, long long capacity,
                                      long long* nclusters, void* stream_v) {
    using namespace gsr;
    cudaStream_t s = (cudaStream_t)stream_v;
    if (int rc = check_sizes("gsr_mesh_cluster_sizes", nverts, 0)) return rc;
    if (!nclusters || capacity < 1 || capacity > 0x7fffffffLL || !sizes || (nverts > 0 && !root_ntris)) {
        set_error("gsr_mesh_cluster_sizes: invalid argument");
        return GSR_E_INVALID;
    }
    *nclusters = 0;
    if (nverts == 0) return GSR_OK;
    // the counter lives in the last word of the caller's buffer
    unsigned* count = sizes + (capacity - 1);
    GSR_CUDA_CHECK(cudaMemsetAsync(count, 0, 4, s));
    collect_cluster_sizes<<<((unsigned)nverts + 255) / 256, 256, 0, s>>>(root_ntris, (unsigned)nverts, sizes, (unsigned)(capacity - 1), count);
    GSR_CUDA_CHECK(cudaGetLastError());
    unsigned host_count = 0;
    GSR_CUDA_CHECK(cudaMemcpyAsync(&host_count, count, 4, cudaMemcpyDeviceToHost, s));
    GSR_CUDA_CHECK(cudaStreamSynchronize(s));
    *nclusters = host_count;
    return GSR_OK;
}

extern "C" int gsr_mesh_keep_clusters(long long ntris, const int* tri_root, const unsigned int* root_ntris, unsigned int min_triangles,
                                      unsigned char* tri_keep, void* stream_v) {
    using namespace gsr;
    if (int rc = check_sizes("gsr_mesh_keep_clusters", 0, ntris)) return rc;
    if (ntris > 0 && (!tri_root || !root_ntris || !tri_keep)) {
        set_error("gsr_mesh_keep_clusters: invalid argument");
        return GSR_E_INVALID;
    }
    if (ntris == 0) return GSR_OK;
    keep_by_cluster_size<<<((unsigned)ntris + 255) / 256, 256, 0, (cudaStream_t)stream_v>>>(tri_root, root_ntris, (unsigned)ntris,
                                                                                           min_triangles, tri_keep);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}

extern "C" size_t gsr_mesh_filter_workspace_bytes(long long nverts, long long ntris) {
    if (nverts < 0 || ntris < 0) return 0;
    return gsr::filter_layout(nullptr, nverts, ntris).bytes;
}

extern "C" int gsr_mesh_filter_count(long long nverts, long long ntris, const int* faces, const unsigned char* tri_keep,
                                     void* workspace, long long* nverts_out, long long* ntris_out, void* stream_v) {
    using namespace gsr;
    cudaStream_t s = (cudaStream_t)stream_v;
    if (int rc = check_sizes("gsr_mesh_filter_count", nverts, ntris)) return rc;
    if (!workspace || !nverts_out || !ntris_out || (ntris > 0 && (!faces || !tri_keep))) {
        set_error("gsr_mesh_filter_count: invalid argument");
        return GSR_E_INVALID;
    }
    const unsigned nv = (unsigned)nverts, nf = (unsigned)ntris;
    const FilterWorkspace w = filter_layout(workspace, nverts, ntris);
    const unsigned vb = (nv + ML_BLOCK - 1) / ML_BLOCK, tb = (nf + ML_BLOCK - 1) / ML_BLOCK;
    GSR_CUDA_CHECK(cudaMemsetAsync(w.totals, 0, 16, s));
    if (nv) GSR_CUDA_CHECK(cudaMemsetAsync(w.vert_flag, 0, nv, s));
    if (nf) filter_mark<<<(nf + 255) / 256, 256, 0, s>>>(faces, nf, nv, tri_keep, w.vert_flag, w.tri_flag);
    if (vb) {
        flag_block_counts<<<vb, ML_THREADS, 0, s>>>(w.vert_flag, nv, w.vert_base);
        const unsigned chunks = (vb + SCAN_CHUNK - 1) / SCAN_CHUNK;
        scan_local<false><<<chunks, 1024, 0, s>>>(w.vert_base, nullptr, vb, w.chunk_sums);
        scan_add<false><<<chunks, 1024, 0, s>>>(w.vert_base, nullptr, vb, w.chunk_sums, w.totals);
    }
    if (tb) {
        flag_block_counts<<<tb, ML_THREADS, 0, s>>>(w.tri_flag, nf, w.tri_base);
        const unsigned chunks = (tb + SCAN_CHUNK - 1) / SCAN_CHUNK;
        scan_local<false><<<chunks, 1024, 0, s>>>(w.tri_base, nullptr, tb, w.chunk_sums);
        scan_add<false><<<chunks, 1024, 0, s>>>(w.tri_base, nullptr, tb, w.chunk_sums, w.totals + 2);
    }
    GSR_CUDA_CHECK(cudaGetLastError());
    unsigned totals[4];
    GSR_CUDA_CHECK(cudaMemcpyAsync(totals, w.totals, sizeof(totals), cudaMemcpyDeviceToHost, s));
    GSR_CUDA_CHECK(cudaStreamSynchronize(s));
    *nverts_out = totals[0];
    *ntris_out = totals[2];
    return GSR_OK;
}

extern "C" int gsr_mesh_filter_emit(long long nverts, long long ntris, const float* verts, const float* colors, const int* faces,
                                    const void* workspace, float* verts_out, float* colors_out, int* faces_out, void* stream_v) {
    using namespace gsr;
    cudaStream_t s = (cudaStream_t)stream_v;
    if (int rc = check_sizes("gsr_mesh_filter_emit", nverts, ntris)) return rc;
    if (!workspace || (nverts > 0 && !verts) || (ntris > 0 && !faces) || ((colors != nullptr) != (colors_out != nullptr))) {
        set_error("gsr_mesh_filter_emit: invalid argument");
        return GSR_E_INVALID;
    }
    const unsigned nv = (unsigned)nverts, nf = (unsigned)ntris;
    const FilterWorkspace w = filter_layout(const_cast<void*>(workspace), nverts, ntris);
    const unsigned vb = (nv + ML_BLOCK - 1) / ML_BLOCK, tb = (nf + ML_BLOCK - 1) / ML_BLOCK;
    if (vb) filter_emit_verts<<<vb, ML_THREADS, 0, s>>>(verts, colors, w.vert_flag, nv, w.vert_base, w.vert_new, verts_out, colors_out);
    if (tb) filter_emit_tris<<<tb, ML_THREADS, 0, s>>>(faces, w.tri_flag, nf, w.tri_base, w.vert_new, faces_out);
    GSR_CUDA_CHECK(cudaGetLastError());
    return GSR_OK;
}
