/* bft_graph.cuh — the coloured de Bruijn graph as a device graph, and the traversals of the reference's
 * src/snippets.c on it.
 *
 * The reference walks the graph one k-mer at a time: get_neighbors (8 trie look-ups, src/bft.c:804-1003) from inside
 * iterate_over_kmers, with visit marks kept in the trie (src/marking.c). Its traversals (BFS/DFS,
 * get_nb_connected_component, extract_simple_paths, src/snippets.c:115-958) are sequential by construction — a queue
 * or the call stack — and their results that do not depend on the iteration order are what is computed here, with
 * data-parallel algorithms:
 *
 *   vertex table    every stored k-mer gets the index the device enumeration gives it (bft_kernels.cuh,
 *                   k_extract_*), and loc2vid maps the storage location a look-up ends in back to that index.
 *   adjacency       adj[v][0..3] = predecessors, adj[v][4..7] = successors (the order of get_neighbors), as vertex
 *                   ids — 8 look-ups per vertex, done once (k_graph_adjacency).
 *   components      lock-free union-find over the edges (hook the larger root under the smaller with atomicCAS,
 *                   compress while searching), instead of one BFS/DFS per component.
 *   simple paths    a vertex with in-degree < 2 and out-degree < 2 is a chain vertex; chains are ranked with
 *                   pointer doubling (log2(longest path) rounds) and every vertex writes its own character of the
 *                   path string.
 * Neighbour look-ups here use plain set membership (see bft_b200_set_reference_exact_branching for the reference's
 * leaf-level deviation, which makes its own BFS and DFS disagree with each other).
 */
#ifndef BFT_GRAPH_CUH
#define BFT_GRAPH_CUH

#include "bft_kernels.cuh"

#define BFT_V_NONE 0xffffffffu

/* neighbour k-mers -> look-ups -> storage locations -> vertex ids: the two-phase neighbour engine of bft_kernels.cuh
 * (filter per query, survivors compacted across the warp, walked 32 at a time) with plain set membership */
template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_graph_adjacency(const bft_view_t v, const uint64_t* __restrict__ vk, size_t n,
                                                             const uint32_t* __restrict__ loc2vid, uint32_t* __restrict__ adj) {
    bft_neighbors_core<W, 1>(v, vk, n, 0, (uint8_t*)0, (uint8_t*)0, (unsigned long long*)0, adj, loc2vid);
}

/* vertex id of each queried k-mer (BFT_V_NONE when it is not stored): the handle under which callers keep their own
 * per-k-mer state — the reference keeps such state as marks inside the trie (set_flag_kmer / get_flag_kmer,
 * src/marking.c, include/bft.h:143-146) */
template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_query_vertex_ids(const bft_view_t v, const uint64_t* __restrict__ kmers, size_t n,
                                                              const uint32_t* __restrict__ loc2vid, uint32_t* __restrict__ vids) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t km[W];
        bft_load_kmer<W>(kmers, i, km);
        uint32_t loc = 0;
        const uint32_t cls = bft_lookup_loc(&v, km, W, 0, (uint32_t*)0, &loc);
        vids[i] = cls != BFT_CLS_NONE ? __ldg(loc2vid + loc) : BFT_V_NONE;
    }
}

/* is_in_subgraph (src/snippets.c:824-881) per colour class: the class holds every requested genome id */
__global__ void __launch_bounds__(BFT_TPB) k_graph_class_filter(const uint32_t* __restrict__ class_rows, int rw, size_t n_classes,
                                                                const uint32_t* __restrict__ want_row, uint8_t* __restrict__ cls_in) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_classes; c += stride) {
        int ok = 1;
        for (int w = 0; w < rw; w++) ok &= (class_rows[c * (size_t)rw + w] & want_row[w]) == want_row[w];
        cls_in[c] = (uint8_t)ok;
    }
}

/* ---- connected components -------------------------------------------------------------------------------------
 * parent[] is a forest in which every link points to a smaller vertex id, so a root is the minimum of its tree. */
__device__ __forceinline__ uint32_t bft_uf_root(uint32_t* parent, uint32_t x) {
    volatile uint32_t* p = parent;
    uint32_t cur = p[x];
    if (cur != x) {
        uint32_t prev = x, next;
        while (cur > (next = p[cur])) { /* cur is not a root: splice prev past it */
            p[prev] = next;
            prev = cur;
            cur = next;
        }
    }
    return cur;
}

__device__ __forceinline__ void bft_uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
    a = bft_uf_root(parent, a);
    b = bft_uf_root(parent, b);
    while (a != b) {
        if (a < b) { const uint32_t t = a; a = b; b = t; } /* a > b: try to hang a under b */
        const uint32_t seen = atomicCAS(parent + a, a, b);
        if (seen == a) break;
        a = seen; /* someone hooked a first: continue from its new parent */
    }
}

__global__ void __launch_bounds__(BFT_TPB) k_graph_iota(uint32_t* __restrict__ p, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = (uint32_t)i;
}

/* one thread per vertex; the adjacency is symmetric, so each undirected edge is taken from its larger end.
 * cls_in == NULL: whole graph (BFS / DFS, src/snippets.c:605-665, 743-771); otherwise only edges between two vertices
 * of the colour subgraph count (BFS_subgraph / DFS_subgraph, :667-741, 773-822). */
__global__ void __launch_bounds__(BFT_TPB) k_graph_hook(const uint32_t* __restrict__ adj, const uint32_t* __restrict__ vcls,
                                                        const uint8_t* __restrict__ cls_in, size_t n, uint32_t* parent) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (cls_in && !cls_in[vcls[i]]) continue;
        const uint4 a = __ldg((const uint4*)(adj + i * 8)), b = __ldg((const uint4*)(adj + i * 8 + 4));
        const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (u[j] == BFT_V_NONE || u[j] >= (uint32_t)i) continue;
            if (cls_in && !cls_in[vcls[u[j]]]) continue;
            bft_uf_union(parent, (uint32_t)i, u[j]);
        }
    }
}

/* labels[v] = smallest vertex id of v's component (BFT_V_NONE outside the subgraph); counts the components */
__global__ void __launch_bounds__(BFT_TPB) k_graph_labels(uint32_t* parent, const uint32_t* __restrict__ vcls,
                                                          const uint8_t* __restrict__ cls_in, size_t n, uint32_t* __restrict__ labels,
                                                          unsigned long long* __restrict__ n_components) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long local = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t l = BFT_V_NONE;
        if (!cls_in || cls_in[vcls[i]]) {
            l = bft_uf_root(parent, (uint32_t)i);
            local += l == (uint32_t)i;
        }
        if (labels) labels[i] = l;
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(n_components, local);
}

/* ---- simple paths (extract_simple_paths / extract_core_simple_paths, src/snippets.c:115-308, 346-571) ----------
 * Chain vertex: fewer than two successors and fewer than two predecessors (:145, :389) and, for core paths, at least
 * `core` genomes in its colour set (:376). */
__global__ void __launch_bounds__(BFT_TPB) k_paths_vertices(const uint32_t* __restrict__ adj, const uint32_t* __restrict__ vcls,
                                                            const uint32_t* __restrict__ class_counts, uint32_t core, size_t n,
                                                            uint8_t* __restrict__ chain, uint32_t* __restrict__ usucc) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4 a = __ldg((const uint4*)(adj + i * 8)), b = __ldg((const uint4*)(adj + i * 8 + 4));
        const int np = (a.x != BFT_V_NONE) + (a.y != BFT_V_NONE) + (a.z != BFT_V_NONE) + (a.w != BFT_V_NONE);
        const int ns = (b.x != BFT_V_NONE) + (b.y != BFT_V_NONE) + (b.z != BFT_V_NONE) + (b.w != BFT_V_NONE);
        const int ok = ns < 2 && np < 2 && (core == 0 || class_counts[vcls[i]] >= core);
        chain[i] = (uint8_t)ok;
        uint32_t s = BFT_V_NONE;
        if (ok && ns == 1) s = b.x != BFT_V_NONE ? b.x : b.y != BFT_V_NONE ? b.y : b.z != BFT_V_NONE ? b.z : b.w;
        usucc[i] = s;
    }
}

/* Link v -> u when u is v's only successor, both are chain vertices and, for core paths, they share at least `core`
 * genomes (intersection_annotations, src/snippets.c:421-427). u's only predecessor is then v, so prev[u] has one writer. */
__global__ void __launch_bounds__(BFT_TPB) k_paths_link(const uint8_t* __restrict__ chain, const uint32_t* __restrict__ usucc,
                                                        const uint32_t* __restrict__ vcls, const uint32_t* __restrict__ class_rows, int rw,
                                                        uint32_t core, size_t n, uint32_t* __restrict__ next, uint32_t* __restrict__ prev) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t u = chain[i] ? usucc[i] : BFT_V_NONE;
        if (u != BFT_V_NONE && (u == (uint32_t)i || !chain[u])) u = BFT_V_NONE;
        if (u != BFT_V_NONE && core) {
            const uint32_t* ra = class_rows + (size_t)vcls[i] * rw;
            const uint32_t* rb = class_rows + (size_t)vcls[u] * rw;
            uint32_t shared = 0;
            for (int w = 0; w < rw; w++) shared += __popc(__ldg(ra + w) & __ldg(rb + w));
            if (shared < core) u = BFT_V_NONE;
        }
        next[i] = u;
        if (u != BFT_V_NONE) prev[u] = (uint32_t)i;
    }
}

/* pointer doubling towards the head of the chain: to[v] = the vertex `dist[v]` links upstream of v (a head points to
 * itself with distance 0). low[] carries the smallest vertex id seen on the way, which names a cycle's break point. */
__global__ void __launch_bounds__(BFT_TPB) k_paths_rank_init(const uint8_t* __restrict__ chain, const uint32_t* __restrict__ prev, size_t n,
                                                             uint32_t* __restrict__ to, uint32_t* __restrict__ dist, uint32_t* __restrict__ low) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t p = chain[i] ? prev[i] : BFT_V_NONE;
        to[i] = p != BFT_V_NONE ? p : (uint32_t)i;
        dist[i] = p != BFT_V_NONE ? 1u : 0u;
        if (low) low[i] = (uint32_t)i;
    }
}

/* *pending is raised while some pointer has not arrived at a head (a vertex without a predecessor link). On a cycle
 * that never happens — and a cycle whose length is a power of two even maps every vertex back onto itself — so
 * arrival is tested on prev[], not on the pointer having stopped moving. */
__global__ void __launch_bounds__(BFT_TPB) k_paths_rank_step(const uint32_t* __restrict__ to0, const uint32_t* __restrict__ dist0,
                                                             const uint32_t* __restrict__ low0, const uint32_t* __restrict__ prev, size_t n,
                                                             uint32_t* __restrict__ to1, uint32_t* __restrict__ dist1,
                                                             uint32_t* __restrict__ low1, int* __restrict__ pending) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    int any = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t t = to0[i];
        const uint32_t tt = to0[t];
        to1[i] = tt;
        dist1[i] = dist0[i] + dist0[t];
        if (low0) low1[i] = min(low0[i], low0[t]);
        any |= prev[tt] != BFT_V_NONE; /* (a vertex outside every chain points to itself and has no predecessor link) */
    }
    if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) *pending = 1;
}

/* after ceil(log2 n) + 2 rounds a vertex whose pointer still has a predecessor sits on a cycle of chain vertices
 * (the reference opens such a cycle at whichever k-mer its iteration reaches first; here: at its smallest vertex).
 * Two kernels, so that no thread tests prev[] while another one is cutting it. */
__global__ void __launch_bounds__(BFT_TPB) k_paths_find_cuts(const uint8_t* __restrict__ chain, const uint32_t* __restrict__ to,
                                                             const uint32_t* __restrict__ low, const uint32_t* __restrict__ prev, size_t n,
                                                             uint8_t* __restrict__ cut) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        cut[i] = chain[i] && prev[i] != BFT_V_NONE && prev[to[i]] != BFT_V_NONE && low[i] == (uint32_t)i;
}

__global__ void __launch_bounds__(BFT_TPB) k_paths_apply_cuts(const uint8_t* __restrict__ cut, size_t n, uint32_t* __restrict__ next,
                                                              uint32_t* __restrict__ prev) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!cut[i]) continue;
        next[prev[i]] = BFT_V_NONE;
        prev[i] = BFT_V_NONE;
    }
}

/* the tail of each path tells its head how many vertices the path has; heads get their output size.
 * stats[0] += paths, stats[1] = max(characters of a path) — reduced per warp before the atomics */
__global__ void __launch_bounds__(BFT_TPB) k_paths_sizes(const uint8_t* __restrict__ chain, const uint32_t* __restrict__ next,
                                                         const uint32_t* __restrict__ to, const uint32_t* __restrict__ dist, size_t n, int k,
                                                         unsigned long long* __restrict__ size, unsigned long long* __restrict__ stats) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long count = 0, longest = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!chain[i] || next[i] != BFT_V_NONE) continue;
        const unsigned long long chars = (unsigned long long)k + dist[i];
        size[to[i]] = chars + 1; /* + '\n' */
        count++;
        longest = max(longest, chars);
    }
    for (int o = 16; o > 0; o >>= 1) {
        count += __shfl_down_sync(0xffffffffu, count, o);
        longest = max(longest, __shfl_down_sync(0xffffffffu, longest, o));
    }
    if ((threadIdx.x & 31) == 0 && count) {
        atomicAdd(stats, count);
        atomicMax(stats + 1, longest);
    }
}

template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_paths_write(const uint8_t* __restrict__ chain, const uint32_t* __restrict__ next,
                                                         const uint32_t* __restrict__ to, const uint32_t* __restrict__ dist,
                                                         const unsigned long long* __restrict__ offs, const uint64_t* __restrict__ vk, size_t n,
                                                         int k, char* __restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (!chain[i]) continue;
        const uint32_t d = dist[i];
        char* p = out + offs[to[i]];
        if (d == 0) {
            for (int j = 0; j < k; j++) p[j] = "ACGT"[(vk[i * W + (size_t)(j >> 5)] >> (2 * (j & 31))) & 3];
        } else {
            const int j = k - 1;
            p[(size_t)k - 1 + d] = "ACGT"[(vk[i * W + (size_t)(j >> 5)] >> (2 * (j & 31))) & 3];
        }
        if (next[i] == BFT_V_NONE) p[(size_t)k + d] = '\n';
    }
}

#endif /* BFT_GRAPH_CUH */
