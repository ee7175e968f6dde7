/* bft_kernels.cuh — sm_100a kernels of the BFT query engine.
 *
 * Nothing on this path is a dense contraction, so tensor cores / TMEM are not used; the kernels are bound by the
 * random-access rate of the memory system (one dependent 32-byte sector per step of a lookup). The design levers
 * are therefore: (1) as few dependent loads per lookup as possible (root directory, prefix-sum directories — see
 * bft_arena.h), (2) as many independent lookups in flight as the SMs hold (one lookup per thread, 2048 threads/SM),
 * (3) coalesced streaming of the query and result arrays, (4) small hot tables (root directory 2 MB, class rows)
 * resident in the 126 MB L2.
 */
#ifndef BFT_KERNELS_CUH
#define BFT_KERNELS_CUH

#include <cuda_runtime.h>
#include "bft_arena.h"
#include "bft_colour.h"

#define BFT_TPB 256
#define BFT_N_WALK_STATS 8 /* counters of k_kmer_walk_stats (bft_lookup_loc documents them) */

/* ---- multi-word helpers (W = 1, 2 or 4 words, compile-time; loops unroll and runtime word indices become selects,
 * so nothing lands in local memory) */
template <int W>
__device__ __forceinline__ void bft_load_kmer(const uint64_t* __restrict__ kmers, size_t i, uint64_t* km) {
    if (W == 1) {
        km[0] = __ldcs((const unsigned long long*)kmers + i);
    } else {
#pragma unroll
        for (int w = 0; w < W; w += 2) {
            const ulonglong2 t = __ldcs((const ulonglong2*)(kmers + i * W + w));
            km[w] = t.x;
            km[w + 1] = t.y;
        }
    }
}

/* mask of the bits of word w that belong to a value of n_bits bits */
__device__ __forceinline__ uint64_t bft_word_mask(int n_bits, int w) {
    const int b = n_bits - 64 * w;
    return b >= 64 ? ~0ULL : (b <= 0 ? 0ULL : ((1ULL << b) - 1ULL));
}

/* out = in >> sh, 0 <= sh < 64*W */
template <int W>
__device__ __forceinline__ void bft_shr(const uint64_t* in, int sh, uint64_t* out) {
    const int ws = sh >> 6, bs = sh & 63;
#pragma unroll
    for (int w = 0; w < W; w++) {
        uint64_t lo = 0, hi = 0; /* selected by value (no indexed access: the words must stay in registers, see BFT_OPAQUE) */
#pragma unroll
        for (int u = 0; u < W; u++) {
            uint64_t t = in[u];
            if (W > 1) BFT_OPAQUE(t);
            lo |= (u == w + ws) ? t : 0ULL;
            hi |= (u == w + ws + 1) ? t : 0ULL;
        }
        out[w] = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
    }
}

/* out = in << sh, 0 <= sh < 64*W (bits shifted past the top are dropped) */
template <int W>
__device__ __forceinline__ void bft_shl(const uint64_t* in, int sh, uint64_t* out) {
    const int ws = sh >> 6, bs = sh & 63;
#pragma unroll
    for (int w = 0; w < W; w++) {
        uint64_t lo = 0, hi = 0; /* hi: the word that lands at w, lo: the one below it */
#pragma unroll
        for (int u = 0; u < W; u++) {
            uint64_t t = in[u];
            if (W > 1) BFT_OPAQUE(t);
            hi |= (u + ws == w) ? t : 0ULL;
            lo |= (u + ws + 1 == w) ? t : 0ULL;
        }
        out[w] = bs ? (hi << bs) | (lo >> (64 - bs)) : hi;
    }
}

/* a >= b as W-word integers */
template <int W>
__device__ __forceinline__ bool bft_ge(const uint64_t* a, const uint64_t* b) {
    bool ge = true; /* equal so far */
#pragma unroll
    for (int w = 0; w < W; w++) { /* from the least significant word up: a higher word overrides */
        if (a[w] > b[w]) ge = true;
        else if (a[w] < b[w]) ge = false;
    }
    return ge;
}

/* The batch's hit count (the number `Nb k-mers present` of the reference driver, src/file_io.c:813): warp shuffle, one
 * shared-memory word per warp, ONE atomic per CTA. The atomic is system-scope because the counter may live in another
 * GPU's HBM (a CUDA-IPC mapping, bft_b200_peer_import): every rank's kernel then adds its share straight into the
 * owner's counter over NVLink — the only "collective" the sharded k-mer path has, fused into the query kernel. */
__device__ __forceinline__ void bft_block_count(unsigned int hits, unsigned long long* counter) {
    __shared__ unsigned int bft_warp_hits[BFT_TPB / 32];
    for (int o = 16; o > 0; o >>= 1) hits += __shfl_down_sync(0xffffffffu, hits, o);
    if ((threadIdx.x & 31) == 0) bft_warp_hits[threadIdx.x >> 5] = hits;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long total = 0;
#pragma unroll
        for (int w = 0; w < BFT_TPB / 32; w++) total += bft_warp_hits[w];
        if (total) atomicAdd_system(counter, total);
    }
}

/* class rows (decoded colour sets, read by every found k-mer): L2-resident table, loaded evict-last so that the streamed batch and
 * the one-touch buckets (evict-first) do not push it out */
__device__ __forceinline__ uint64_t bft_policy_evict_last() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t bft_ld_row(const uint32_t* p) {
    uint32_t r;
    asm("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(bft_policy_evict_last()));
    return r;
}
__device__ __forceinline__ uint2 bft_ld_row(const uint2* p) {
    uint2 r;
    asm("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(r.x), "=r"(r.y) : "l"(p), "l"(bft_policy_evict_last()));
    return r;
}
__device__ __forceinline__ uint4 bft_ld_row(const uint4* p) {
    uint4 r;
    asm("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(bft_policy_evict_last()));
    return r;
}

/* ---- a4/a5/a7/a8: k-mer lookup ---------------------------------------------------------------------------- */
template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_query_kmers(const bft_view_t v, const uint64_t* __restrict__ kmers, size_t n,
                                                         uint8_t* __restrict__ present, uint32_t* __restrict__ cls_out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t km[W];
        bft_load_kmer<W>(kmers, i, km);
        const uint32_t cls = bft_lookup_w(&v, km, W);
        if (present) present[i] = cls != BFT_CLS_NONE;
        if (cls_out) cls_out[i] = cls;
    }
}

/* a4-a10 fused for narrow colour rows (RW = 1, 2 or 4 words, i.e. up to 128 genomes): the thread that finished a
 * lookup fetches its class row (L2-resident table) and stores it with one vector store — no class-id round trip
 * through HBM and no second launch. Wider rows go through k_expand_rows. */
template <int W, int RW>
__global__ void __launch_bounds__(BFT_TPB) k_query_kmers_rows(const bft_view_t v, const uint64_t* __restrict__ kmers, size_t n,
                                                              uint8_t* __restrict__ present, uint32_t* __restrict__ cls_out,
                                                              const uint32_t* __restrict__ class_rows, uint32_t* __restrict__ rows,
                                                              unsigned long long* __restrict__ n_present) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned int hits = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t km[W];
        bft_load_kmer<W>(kmers, i, km);
        const uint32_t cls = bft_lookup_w(&v, km, W);
        hits += cls != BFT_CLS_NONE;
        if (present) present[i] = cls != BFT_CLS_NONE;
        if (cls_out) cls_out[i] = cls;
        if (RW == 4) {
            uint4 r = make_uint4(0, 0, 0, 0);
            if (cls != BFT_CLS_NONE) r = bft_ld_row((const uint4*)class_rows + cls);
            __stcs((uint4*)rows + i, r);
        } else if (RW == 2) {
            uint2 r = make_uint2(0, 0);
            if (cls != BFT_CLS_NONE) r = bft_ld_row((const uint2*)class_rows + cls);
            __stcs((uint2*)rows + i, r);
        } else {
            uint32_t r = 0;
            if (cls != BFT_CLS_NONE) r = bft_ld_row(class_rows + cls);
            __stcs(rows + i, r);
        }
    }
    if (n_present) bft_block_count(hits, n_present);
}

/* a4-a10 fused for WIDE colour rows (more than 128 genomes; 1000 colours = 32 words = 128 bytes per k-mer): each lane looks
 * one k-mer up, then the warp writes the 32 rows of its 32 consecutive k-mers together — lane t of pass p copies vector
 * (32 p + t) of the 32 * RWV-vector block, whose class id comes from the owning lane by shuffle. The block is contiguous in
 * the output (32 rows back to back), so every pass is one fully coalesced 512-byte store; the class rows come from the
 * L2-resident table. Against k_query_kmers + k_expand_rows_v4 this drops the class-id round trip through HBM (8 B per k-mer)
 * and the second launch, and — the point — lets the latency-bound walks of some warps overlap the bandwidth-bound row
 * writes of others instead of running one after the other. T = uint4 when the row width and the buffers allow, else uint32_t. */
template <int W, typename T>
__global__ void __launch_bounds__(BFT_TPB) k_query_kmers_wide(const bft_view_t v, const uint64_t* __restrict__ kmers, size_t n,
                                                              uint8_t* __restrict__ present, uint32_t* __restrict__ cls_out,
                                                              const T* __restrict__ class_rows, int rwv, T* __restrict__ rows,
                                                              unsigned long long* __restrict__ n_present) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    unsigned int hits = 0;
    const int per_warp = 32 * rwv;
    /* warp-uniform trip count: the lanes of a warp own 32 consecutive k-mers, the last warp may run past n */
    for (size_t base = (size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n; base += stride) {
        const size_t i = base + lane;
        uint32_t cls = BFT_CLS_NONE;
        if (i < n) {
            uint64_t km[W];
            bft_load_kmer<W>(kmers, i, km);
            cls = bft_lookup_w(&v, km, W);
            hits += cls != BFT_CLS_NONE;
            if (present) present[i] = cls != BFT_CLS_NONE;
            if (cls_out) cls_out[i] = cls;
        }
        T* const out = rows + base * (size_t)rwv;
        const size_t live = (n - base < 32 ? n - base : 32) * (size_t)rwv; /* vectors of this block that exist */
#pragma unroll 2
        for (int t = lane; t < per_warp; t += 32) {
            const int kk = t / rwv, w = t - kk * rwv;
            const uint32_t c = __shfl_sync(0xffffffffu, cls, kk);
            T r = T();
            /* a row table larger than L2 (1000 colours x 10^6 classes = 127 MB) gets no priority: it would only push the
             * root directory and the filters out */
            if (c != BFT_CLS_NONE) r = v.rows_keep ? bft_ld_row(class_rows + (size_t)c * rwv + w) : __ldg(class_rows + (size_t)c * rwv + w);
            if ((size_t)t < live) __stcs(out + t, r);
        }
    }
    if (n_present) bft_block_count(hits, n_present);
}

/* The same look-up on the reference's own record format: k-mers as ceil(2k/8)-byte records (a kmers_comp file,
 * BFT_kmer.kmer_comp; src/file_io.c:721-774) in, colour rows of ceil(n_genomes/8) bytes out (bit g = genome g; an absent
 * k-mer has an all-zero row). For k = 27 and 100 genomes that is 7 + 13 bytes per k-mer over PCIe instead of 8 + 17,
 * which is what bounds the host-facing call.
 *
 * A block handles tiles of BFT_TPB k-mers. BFT_TPB * anything is a multiple of 16, so with 16-byte aligned buffers a
 * full tile of records is ONE bulk asynchronous copy global -> shared (cp.async.bulk + mbarrier transaction count, the
 * TMA engine's 1-D form; SASS UBLKCP) and a full tile of rows ONE bulk copy shared -> global. Both are double
 * buffered: the records of tile t+1 arrive while the threads look tile t up, and the rows of tile t drain while
 * tile t+1 is looked up. The last, partial tile of a batch is moved by the threads themselves. */
__device__ __forceinline__ uint32_t bft_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bft_mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bft_smem_addr(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void bft_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bft_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bft_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(bft_smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bft_bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(bft_smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(bft_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bft_bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(bft_smem_addr(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

/* dynamic shared memory of k_query_records: two record tiles, two row tiles, two mbarriers */
__host__ __device__ inline size_t bft_records_smem(int nb, int rb) { return 2 * (size_t)BFT_TPB * (size_t)(nb + rb) + 16; }

/* COMPACT = false: rows of every k-mer, fixed stride (bft_b200_query_records).
 * COMPACT = true : first pass of bft_b200_query_records_compact — class id per k-mer, presence as one bit per k-mer
 *                  (a warp ballot is 32 of them) and the number of hits of every tile; k_scan_tile_counts and
 *                  k_compact_rows then lay the rows of the present k-mers out back to back, in query order. */
template <int W, bool COMPACT>
__global__ void __launch_bounds__(BFT_TPB) k_query_records(const bft_view_t v, const uint8_t* __restrict__ records, size_t n, int nb, int rb,
                                                           int rw, const uint32_t* __restrict__ class_rows, uint8_t* __restrict__ present,
                                                           uint8_t* __restrict__ rows, unsigned long long* __restrict__ n_present,
                                                           uint32_t* __restrict__ cls_out, uint32_t* __restrict__ present_bits,
                                                           uint32_t* __restrict__ tile_cnt) {
    extern __shared__ uint4 bft_tile_smem[];
    uint8_t* const sm = (uint8_t*)bft_tile_smem;
    const uint32_t in_bytes = (uint32_t)BFT_TPB * (uint32_t)nb, out_bytes = (uint32_t)BFT_TPB * (uint32_t)rb;
    uint8_t* const out_base = sm + 2 * (size_t)in_bytes;
#define BFT_IN_BUF(i_) (sm + (size_t)(i_) * in_bytes)
#define BFT_OUT_BUF(i_) (out_base + (size_t)(i_) * out_bytes)
    uint64_t* const bar = (uint64_t*)(sm + 2 * (size_t)in_bytes + 2 * (size_t)out_bytes);
    const size_t n_tiles = (n + BFT_TPB - 1) / BFT_TPB;
    const bool lead = threadIdx.x == 0;
    if (lead) {
        bft_mbar_init(bar, 1);
        bft_mbar_init(bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __shared__ unsigned int tile_hits[2];
    if (threadIdx.x < 2) tile_hits[threadIdx.x] = 0;
    __syncthreads();
    unsigned int hits = 0;
    uint32_t it = 0;
    size_t tile = blockIdx.x;
    if (lead && tile < n_tiles && n - tile * BFT_TPB >= BFT_TPB) { /* first tile of this block */
        bft_mbar_expect_tx(bar, in_bytes);
        bft_bulk_load(BFT_IN_BUF(0), records + tile * (size_t)in_bytes, in_bytes, bar);
    }
    for (; tile < n_tiles; tile += gridDim.x, it++) {
        const uint32_t b = it & 1u;
        const size_t base = tile * BFT_TPB;
        const bool full = n - base >= BFT_TPB;
        const size_t cnt = full ? BFT_TPB : n - base;
        const size_t next = tile + gridDim.x;
        if (lead && next < n_tiles && n - next * BFT_TPB >= BFT_TPB) { /* prefetch; the other record buffer was released at the last barrier of it-1 */
            bft_mbar_expect_tx(bar + (b ^ 1u), in_bytes);
            bft_bulk_load(BFT_IN_BUF(b ^ 1u), records + next * (size_t)in_bytes, in_bytes, bar + (b ^ 1u));
        }
        if (full) {
            bft_mbar_wait(bar + b, (it >> 1) & 1u); /* buffer b is filled for the (it / 2)-th time */
        } else {
            for (size_t i = threadIdx.x; i < cnt * (size_t)nb; i += blockDim.x) BFT_IN_BUF(b)[i] = records[base * (size_t)nb + i];
            __syncthreads();
        }
        uint32_t cls = BFT_CLS_NONE;
        if (threadIdx.x < cnt) {
            uint64_t km[W];
#pragma unroll
            for (int w = 0; w < W; w++) km[w] = 0;
            const uint8_t* r = BFT_IN_BUF(b) + (size_t)threadIdx.x * nb;
#pragma unroll
            for (int w = 0; w < W; w++)
                for (int j = 0; j < 8; j++)
                    if (w * 8 + j < nb) km[w] |= (uint64_t)r[w * 8 + j] << (8 * j);
#pragma unroll
            for (int w = 0; w < W; w++) km[w] &= bft_word_mask(2 * v.k, w); /* the reference ignores the pad bits of the last byte */
            cls = bft_lookup_w(&v, km, W);
        }
        if constexpr (COMPACT) {
            const bool hit = cls != BFT_CLS_NONE;
            const uint32_t ball = __ballot_sync(0xffffffffu, hit);
            if (threadIdx.x < cnt) cls_out[base + threadIdx.x] = cls;
            if ((threadIdx.x & 31) == 0) {
                if (base + threadIdx.x < n) present_bits[(base + threadIdx.x) >> 5] = ball;
                if (ball) atomicAdd(&tile_hits[b], (unsigned int)__popc(ball));
            }
            __syncthreads(); /* every warp has read its records and added its hits */
            if (lead) {
                tile_cnt[tile] = tile_hits[b];
                tile_hits[b] = 0; /* next used two tiles from now, behind another barrier */
            }
            continue;
        }
        /* the rows of tile it-2 must have left this row buffer before it is written again */
        if (lead) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < cnt) {
            const bool hit = cls != BFT_CLS_NONE;
            hits += hit;
            if (present) present[base + threadIdx.x] = hit;
            uint8_t* o = BFT_OUT_BUF(b) + (size_t)threadIdx.x * rb;
            for (int j = 0; j < rb; j += 4) {
                const uint32_t word = !hit ? 0u : (v.rows_keep ? bft_ld_row(class_rows + (size_t)cls * rw + (j >> 2)) : __ldg(class_rows + (size_t)cls * rw + (j >> 2)));
                for (int q = 0; q < 4 && j + q < rb; q++) o[j + q] = (uint8_t)(word >> (8 * q));
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy writes -> visible to the bulk copy */
        __syncthreads();
        if (full) {
            if (lead) bft_bulk_store(rows + base * (size_t)rb, BFT_OUT_BUF(b), out_bytes);
        } else {
            for (size_t i = threadIdx.x; i < cnt * (size_t)rb; i += blockDim.x) rows[base * (size_t)rb + i] = BFT_OUT_BUF(b)[i];
        }
    }
#undef BFT_IN_BUF
#undef BFT_OUT_BUF
    if (lead) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); /* shared memory must outlive the copies reading it */
    if (n_present) bft_block_count(hits, n_present);
}

/* exclusive prefix sum of the per-tile hit counts of one chunk (at most a few thousand tiles): one block */
__global__ void __launch_bounds__(1024) k_scan_tile_counts(const uint32_t* __restrict__ tile_cnt, uint32_t n_tiles, uint32_t* __restrict__ tile_off,
                                                           uint32_t* __restrict__ total) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (n_tiles + 1023u) / 1024u;
    const uint32_t lo = threadIdx.x * per, hi = min(n_tiles, lo + per);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += tile_cnt[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) { /* Hillis-Steele inclusive scan of the 1024 partial sums */
        const uint32_t t = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - sum;
    for (uint32_t i = lo; i < hi; i++) {
        tile_off[i] = run;
        run += tile_cnt[i];
    }
    if (threadIdx.x == 1023) *total = part[1023];
}

/* second pass of the compact form: the rows of a tile's present k-mers, in order, to rows[tile_off[tile] ...] */
__global__ void __launch_bounds__(BFT_TPB) k_compact_rows(const uint32_t* __restrict__ cls, size_t n, const uint32_t* __restrict__ tile_off,
                                                          const uint32_t* __restrict__ class_rows, int rw, int rb, uint8_t* __restrict__ rows) {
    extern __shared__ uint4 bft_tile_smem[];
    uint8_t* const sm = (uint8_t*)bft_tile_smem;
    __shared__ uint32_t warp_hits[BFT_TPB / 32];
    const size_t n_tiles = (n + BFT_TPB - 1) / BFT_TPB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const size_t i = tile * BFT_TPB + threadIdx.x;
        const uint32_t c = i < n ? __ldg(cls + i) : BFT_CLS_NONE;
        const bool hit = c != BFT_CLS_NONE;
        const uint32_t ball = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_hits[warp] = (uint32_t)__popc(ball);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < BFT_TPB / 32; w++) {
            const uint32_t h = warp_hits[w];
            before += w < warp ? h : 0u;
            total += h;
        }
        if (hit) {
            uint8_t* o = sm + (size_t)(before + (uint32_t)__popc(ball & ((1u << lane) - 1u))) * rb;
            for (int j = 0; j < rb; j += 4) {
                const uint32_t word = bft_ld_row(class_rows + (size_t)c * rw + (j >> 2));
                for (int q = 0; q < 4 && j + q < rb; q++) o[j + q] = (uint8_t)(word >> (8 * q));
            }
        }
        __syncthreads();
        uint8_t* dst = rows + (size_t)tile_off[tile] * rb;
        for (size_t q = threadIdx.x; q < (size_t)total * rb; q += blockDim.x) dst[q] = sm[q];
        __syncthreads();
    }
}

/* Random-access roofline probe: n independent 8-byte loads at pseudo-random offsets of a table far
 * larger than L2, one per thread per iteration — the rate the memory system sustains for dependent-free random
 * sectors. Not on the product path; bench.py runs it to put the walk's sector rate in context. */
__global__ void __launch_bounds__(BFT_TPB) k_random_gather(const uint64_t* __restrict__ table, size_t n_words, size_t n_loads,
                                                           unsigned long long* __restrict__ sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_loads; i += stride) {
        uint64_t x = (i + 0x9E3779B97F4A7C15ULL) * 0xBF58476D1CE4E5B9ULL; /* splitmix-style scramble */
        x ^= x >> 31; x *= 0x94D049BB133111EBULL; x ^= x >> 29;
        acc += __ldg(table + (x % n_words));
    }
    if (acc == 0x1234567ULL) *sink = acc; /* keep the loads alive */
}

/* Instrumented walk for the roofline accounting: sums over the batch of Nodes probed, binary-search
 * depths ceil(log2(lines+1)), hits, CCs the reference would probe and block sizes. Not on the product path; bench.py runs it once on the timed batch. */
template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_kmer_walk_stats(const bft_view_t v, const uint64_t* __restrict__ kmers, size_t n,
                                                             unsigned long long* __restrict__ acc) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long a[BFT_N_WALK_STATS] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t km[W];
#pragma unroll
        for (int w = 0; w < W; w++) km[w] = kmers[i * W + w];
        uint32_t st[BFT_N_WALK_STATS] = {0, 0, 0, 0, 0, 0, 0, 0};
        bft_lookup_ex(&v, km, W, 0, st);
#pragma unroll
        for (int j = 0; j < BFT_N_WALK_STATS; j++) a[j] += st[j];
    }
#pragma unroll
    for (int j = 0; j < BFT_N_WALK_STATS; j++) {
        for (int o = 16; o > 0; o >>= 1) a[j] += __shfl_down_sync(0xffffffffu, a[j], o);
        if ((threadIdx.x & 31) == 0) atomicAdd(acc + j, a[j]);
    }
}

/* a10 (read side of get_list_id_genomes): class id -> colour row, one thread per output word */
__global__ void __launch_bounds__(BFT_TPB) k_expand_rows(const uint32_t* __restrict__ cls, size_t n, const uint32_t* __restrict__ class_rows,
                                                         int rw, uint32_t* __restrict__ rows) {
    const size_t total = n * (size_t)rw;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const size_t i = t / (size_t)rw;
        const int w = (int)(t - i * (size_t)rw);
        const uint32_t c = cls[i];
        rows[t] = c == BFT_CLS_NONE ? 0u : __ldg(class_rows + (size_t)c * rw + w);
    }
}

/* Annotation set algebra, batched (intersection_/union_/sym_difference_annotations, src/bft.c:421-613, which fold cmp_annots,
 * src/annotation.c:2358-2552, over their arguments from left to right): group g combines the colour rows of the classes
 * cls[offs[g] .. offs[g+1]) word by word. One thread per (group, row word); the per-group genome count (get_count_id_genomes
 * of the result) is accumulated with one atomic per word. A class id of BFT_CLS_NONE (an absent k-mer) is the empty set. */
__global__ void __launch_bounds__(BFT_TPB) k_annotation_setop(const uint32_t* __restrict__ class_rows, int rw, const uint32_t* __restrict__ cls,
                                                              const uint64_t* __restrict__ offs, size_t n_groups, int op,
                                                              uint32_t* __restrict__ rows, uint32_t* __restrict__ counts) {
    const size_t total = n_groups * (size_t)rw;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const size_t g = t / (size_t)rw;
        const int w = (int)(t - g * (size_t)rw);
        const uint64_t b = offs[g], e = offs[g + 1];
        uint32_t acc = 0;
        for (uint64_t i = b; i < e; i++) {
            const uint32_t c = __ldg(cls + i);
            const uint32_t x = c == BFT_CLS_NONE ? 0u : __ldg(class_rows + (size_t)c * rw + w);
            if (i == b) acc = x;
            else if (op == 0) acc &= x;
            else if (op == 1) acc |= x;
            else acc ^= x;
        }
        if (rows) rows[t] = acc;
        if (counts) {
            if (rw == 1) counts[g] = (uint32_t)__popc(acc);
            else if (acc) atomicAdd(counts + g, (uint32_t)__popc(acc));
        }
    }
}

/* a9/a10: decode every distinct annotation once (get_id_genomes_from_annot, src/annotation.c:2086-2250) */
__global__ void __launch_bounds__(BFT_TPB) k_decode_classes(const uint32_t* __restrict__ cls_off, const uint8_t* __restrict__ cls_bytes,
                                                            size_t n_classes, const bft_pools_t pools, uint32_t* __restrict__ rows,
                                                            uint32_t* __restrict__ counts, int rw, int* __restrict__ n_bad) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_classes; c += stride) {
        uint32_t* row = rows + c * (size_t)rw;
        for (int w = 0; w < rw; w++) row[w] = 0;
        const uint32_t o = cls_off[c];
        if (bft_decode_annotation(cls_bytes + o, (int)(cls_off[c + 1] - o), &pools, row, rw)) atomicAdd(n_bad, 1);
        uint32_t cnt = 0;
        for (int w = 0; w < rw; w++) cnt += __popc(row[w]);
        counts[c] = cnt;
    }
}

/* a1: ASCII -> packed k-mers (parseKmerCount, src/fasta.c:3-53). One thread per k-mer. */
template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_encode_ascii(const char* __restrict__ ascii, size_t n, int k, uint64_t* __restrict__ kmers,
                                                          uint8_t* __restrict__ valid) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const char* s = ascii + i * (size_t)k;
        uint64_t km[W];
#pragma unroll
        for (int w = 0; w < W; w++) km[w] = 0;
        int ok = 1;
#pragma unroll
        for (int w = 0; w < W; w++) { /* word by word, so that km[] is never indexed at run time */
            uint64_t word = 0;
            const int j1 = k - 32 * w < 32 ? k - 32 * w : 32;
            for (int jj = 0; jj < j1; jj++) {
                uint64_t code = 0;
                switch (s[32 * w + jj]) {
                    case 'A': case 'a': code = 0; break;
                    case 'C': case 'c': code = 1; break;
                    case 'G': case 'g': code = 2; break;
                    case 'T': case 't': case 'U': case 'u': code = 3; break;
                    default: ok = 0; break;
                }
                word |= code << (2 * jj);
            }
            km[w] = word;
        }
        if (!ok) {
#pragma unroll
            for (int w = 0; w < W; w++) km[w] = 0;
        }
#pragma unroll
        for (int w = 0; w < W; w++) kmers[i * W + w] = km[w];
        valid[i] = (uint8_t)ok;
    }
}

/* k-mers that failed to parse were looked up as AAA...A (which may well be stored): blank their answers */
__global__ void __launch_bounds__(BFT_TPB) k_blank_invalid(const uint8_t* __restrict__ valid, size_t n, int rw, uint8_t* __restrict__ present,
                                                           uint32_t* __restrict__ cls, uint32_t* __restrict__ rows) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (valid[i]) continue;
        if (present) present[i] = 0;
        if (cls) cls[i] = BFT_CLS_NONE;
        if (rows)
            for (int w = 0; w < rw; w++) rows[i * (size_t)rw + w] = 0;
    }
}

/* Build of the stored-k-mer filter (bft_arena.h): every stored k-mer, as enumerated by k_extract_*, sets its four bits. */
template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_kf_insert(const uint64_t* __restrict__ kmers, size_t n, int k, unsigned long long* __restrict__ filter,
                                                       uint32_t n_blocks) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t km[W];
#pragma unroll
        for (int w = 0; w < W; w++) km[w] = kmers[i * W + w];
        const bft_kf_pos_t q = bft_kf_pos(km, W, k, n_blocks);
        unsigned long long* p = filter + (size_t)q.block * 4;
        atomicOr(p + 0, 1ULL << q.b0);
        atomicOr(p + 1, 1ULL << q.b1);
        atomicOr(p + 2, 1ULL << q.b2);
        atomicOr(p + 3, 1ULL << q.b3);
    }
}

/* Build of the fused root directory + filter (bft_arena.h, rootkf): every sector gets its prefix's root entry; every stored
 * k-mer sets its three bits in the sector its hash names, and is counted under its prefix (bft_b200_open sizes the table and
 * decides from the counts whether the k-mers are spread evenly enough over the prefixes for it to filter at all). */
__global__ void __launch_bounds__(BFT_TPB) k_rkf_fill_entries(const bft_entry_t* __restrict__ rootdir, unsigned long long* __restrict__ rootkf,
                                                              uint32_t n_sectors) {
    const size_t total = (size_t)BFT_ROOTDIR_SIZE * n_sectors, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const bft_entry_t e = rootdir[i / n_sectors];
        rootkf[i * 4] = (unsigned long long)e.a | ((unsigned long long)e.b << 32);
    }
}

template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_rkf_count(const uint64_t* __restrict__ kmers, size_t n, uint32_t* __restrict__ per_prefix) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        atomicAdd(per_prefix + ((uint32_t)kmers[i * W] & (BFT_ROOTDIR_SIZE - 1u)), 1u);
}

template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_rkf_insert(const uint64_t* __restrict__ kmers, size_t n, unsigned long long* __restrict__ rootkf,
                                                        uint32_t n_sectors) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t km[W];
#pragma unroll
        for (int w = 0; w < W; w++) km[w] = kmers[i * W + w];
        const bft_rkf_pos_t q = bft_rkf_pos(km, W, n_sectors);
        unsigned long long* p = rootkf + ((size_t)((uint32_t)km[0] & (BFT_ROOTDIR_SIZE - 1u)) * n_sectors + q.j) * 4;
        atomicOr(p + 1, 1ULL << q.b1);
        atomicOr(p + 2, 1ULL << q.b2);
        atomicOr(p + 3, 1ULL << q.b3);
    }
}

/* Build of the collapsed subtrees (bft_arena.h, rootdir_fast / dbuckets). k_deep_count: how many stored k-mers sit below each
 * root prefix whose suffixes live in a child Node. k_deep_insert: every such k-mer, minus its first 9 nucleotides, claims a slot
 * of its block — the bucket its hash names, or the next one with room — with one compare-and-swap on the slot's top word; the
 * words below and the class are written once the slot is owned (nothing reads the blocks before the build has completed). */
template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_deep_count(const bft_entry_t* __restrict__ rootdir, const uint64_t* __restrict__ kmers, size_t n,
                                                         uint32_t* __restrict__ per_prefix) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t p = (uint32_t)kmers[i * W] & (BFT_ROOTDIR_SIZE - 1u);
        if ((rootdir[p].b >> BFT_KIND_SHIFT) == BFT_KIND_NODE) atomicAdd(per_prefix + p, 1u);
    }
}

template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_deep_insert(const bft_entry_t* __restrict__ fastdir, const uint64_t* __restrict__ kmers,
                                                          const uint32_t* __restrict__ cls, size_t n, unsigned long long* __restrict__ dbuckets,
                                                          uint32_t* __restrict__ dslotcls, int cls_shift, unsigned int* __restrict__ failed) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t key[W];
#pragma unroll
        for (int w = 0; w < W; w++) key[w] = kmers[i * W + w];
        const bft_entry_t e = fastdir[(uint32_t)key[0] & (BFT_ROOTDIR_SIZE - 1u)];
        if ((e.b >> BFT_KIND_SHIFT) != BFT_KIND_DEEP) continue;
        bft_shift18(key, W);
        const uint32_t lb = (e.b >> BFT_LB_SHIFT) & BFT_LB_MASK, mask = (1u << lb) - 1u;
        const unsigned long long top = key[W - 1] | (cls_shift ? (unsigned long long)cls[i] << cls_shift : 0ULL);
        uint32_t b = bft_bucket_of(key, W, lb);
        bool placed = false;
        for (uint32_t probe = 0; probe <= mask && !placed; probe++, b = (b + 1u) & mask) {
            const size_t bucket = (size_t)e.a + b;
            for (int j = 0; j < BFT_BUCKET_KEYS && !placed; j++) {
                unsigned long long* slot = dbuckets + (bucket * BFT_BUCKET_KEYS + j) * W;
                if (atomicCAS(slot + (W - 1), (unsigned long long)BFT_SLOT_EMPTY, top) == (unsigned long long)BFT_SLOT_EMPTY) {
#pragma unroll
                    for (int w = 0; w < W - 1; w++) slot[w] = key[w];
                    if (!cls_shift) dslotcls[bucket * BFT_BUCKET_KEYS + j] = cls[i];
                    placed = true;
                }
            }
        }
        if (!placed) atomicAdd(failed, 1u); /* cannot happen at load <= 1/2; checked by the host */
    }
}

/* ---- enumeration: iterate_over_kmers / -extract_kmers (include/bft.h:88,164; src/extract_kmers.c:3-597) ------------
 * One warp per stored prefix. The k-mer is re-assembled from the Node's path (the prefixes above it), the prefix's own
 * 9 nucleotides and the suffix found in its buckets; the output slot of every k-mer is fixed by the exclusive counts
 * the serializer stored (pref_out), so the result is deterministic and needs no atomics. */
template <int W>
__device__ __forceinline__ void bft_emit_kmer(const uint64_t* base, const uint64_t* suffix, int shift_bits, uint64_t* out) {
    /* out = base | suffix << shift_bits (shift_bits = 18 * (depth + 1), 18..234) */
    uint64_t sh[W];
    bft_shl<W>(suffix, shift_bits, sh);
#pragma unroll
    for (int w = 0; w < W; w++) out[w] = base[w] | sh[w];
}

/* byte-swapped word: memcmp over the reference's suffix bytes (byte 0 first, src/UC.c:81-124) orders the lines of a UC
 * like the words bswap(key[0]), bswap(key[1]), ... compared lexicographically */
__device__ __forceinline__ uint64_t bft_bswap64(uint64_t x) {
    return ((uint64_t)__byte_perm((uint32_t)x, 0, 0x0123) << 32) | (uint64_t)__byte_perm((uint32_t)(x >> 32), 0, 0x0123);
}

#define BFT_EXTRACT_WARPS 4
#define BFT_EXTRACT_MAX_LINES 256 /* a prefix owns at most 255 inline lines (children_type counts are bytes, include/CC.h:358-366) */
__host__ __device__ inline size_t bft_extract_smem(int W) { return (size_t)BFT_EXTRACT_WARPS * BFT_EXTRACT_MAX_LINES * ((size_t)W * 8 + 8); }

/* One warp per stored prefix: the lines of its hashed buckets (and overflow runs) are gathered into shared memory, each
 * line's rank in the reference's stored order (ascending memcmp of the suffix bytes) is counted, and the line is written
 * at pref_out[prefix] + rank — the very order iterate_over_kmers_from_node visits them (src/extract_kmers.c:3-597). */
template <int W>
__global__ void __launch_bounds__(32 * BFT_EXTRACT_WARPS) k_extract_prefix_kmers(const bft_view_t v, size_t n_pref, uint64_t* __restrict__ kmers,
                                                                                 uint32_t* __restrict__ cls_out, uint32_t* __restrict__ loc2vid) {
    extern __shared__ __align__(16) unsigned char bft_extract_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* const skey = (uint64_t*)bft_extract_smem_raw + (size_t)warp * BFT_EXTRACT_MAX_LINES * W;
    uint32_t* const scls = (uint32_t*)((uint64_t*)bft_extract_smem_raw + (size_t)BFT_EXTRACT_WARPS * BFT_EXTRACT_MAX_LINES * W) +
                           (size_t)warp * 2 * BFT_EXTRACT_MAX_LINES;
    uint32_t* const sloc = scls + BFT_EXTRACT_MAX_LINES;
    const size_t warp_stride = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int shift = v.cls_shift;
    const uint64_t top_mask = shift ? ((1ULL << shift) - 1ULL) : ~BFT_SLOT_SPECIAL;
    for (size_t j = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n_pref; j += warp_stride) {
        const bft_entry_t e = v.pref[j];
        const uint32_t kind = e.b >> BFT_KIND_SHIFT;
        if (kind != BFT_KIND_INLINE && kind != BFT_KIND_LEAF) continue;
        const bft_path_t path = v.node_path[v.pref_node[j]];
        uint64_t base[W];
        {
            uint64_t lw[W], sh[W];
#pragma unroll
            for (int w = 0; w < W; w++) lw[w] = 0;
            lw[0] = v.pref_low18[j];
            bft_shl<W>(lw, (int)(BFT_PREFIX_BITS * path.depth), sh);
#pragma unroll
            for (int w = 0; w < W; w++) base[w] = path.acc[w] | sh[w];
        }
        const uint64_t out = v.pref_out[j];
        if (kind == BFT_KIND_LEAF) {
            if (lane == 0) {
                for (int w = 0; w < W; w++) kmers[out * W + w] = base[w];
                if (cls_out) cls_out[out] = e.a;
                if (loc2vid) loc2vid[v.loc_leaf + j] = (uint32_t)out;
            }
            continue;
        }
        const int shift_bits = BFT_PREFIX_BITS * (int)(path.depth + 1);
        const uint32_t n_slots = BFT_BUCKET_KEYS * BFT_INLINE_NBK(e);
        /* gather: every line of the block -> (byte-swapped key, class, storage location), in any order */
        uint32_t n_got = 0;
        for (uint32_t s0 = 0; s0 < n_slots; s0 += 32) {
            const uint32_t slot = s0 + lane;
            uint64_t key[W];
            uint32_t n_emit = 0, ovf_start = 0;
            size_t gslot = 0;
            if (slot < n_slots) {
                gslot = (size_t)e.a * BFT_BUCKET_KEYS + slot;
                for (int w = 0; w < W; w++) key[w] = v.buckets[gslot * W + w];
                const uint64_t top = key[W - 1];
                if (!(top & BFT_SLOT_SPECIAL)) n_emit = 1;
                else if (top != BFT_SLOT_EMPTY) { n_emit = (uint32_t)(top >> 32) & 0x7fffffffu; ovf_start = (uint32_t)top; }
            }
            uint32_t incl = n_emit;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t my = n_got + incl - n_emit;
            n_got += __shfl_sync(0xffffffffu, incl, 31);
            if (n_emit == 1 && !(key[W - 1] & BFT_SLOT_SPECIAL)) {
                if (my < BFT_EXTRACT_MAX_LINES) {
                    scls[my] = shift ? ((uint32_t)(key[W - 1] >> shift) & v.cls_mask) : v.slotcls[gslot];
                    sloc[my] = (uint32_t)gslot;
                    key[W - 1] &= top_mask;
                    for (int w = 0; w < W; w++) skey[(size_t)my * W + w] = bft_bswap64(key[w]);
                }
            } else if (n_emit) { /* overflow run of this bucket */
                for (uint32_t i = 0; i < n_emit && my + i < BFT_EXTRACT_MAX_LINES; i++) {
                    uint64_t ok[W];
                    for (int w = 0; w < W; w++) ok[w] = v.ovf[((size_t)ovf_start + i) * W + w];
                    scls[my + i] = shift ? ((uint32_t)(ok[W - 1] >> shift) & v.cls_mask) : v.ovfcls[ovf_start + i];
                    sloc[my + i] = v.loc_ovf + ovf_start + i;
                    ok[W - 1] &= top_mask;
                    for (int w = 0; w < W; w++) skey[(size_t)(my + i) * W + w] = bft_bswap64(ok[w]);
                }
            }
        }
        __syncwarp();
        if (n_got > BFT_EXTRACT_MAX_LINES) n_got = BFT_EXTRACT_MAX_LINES; /* cannot happen: the serializer refuses such blocks */
        /* rank by counting, then emit */
        for (uint32_t i = lane; i < n_got; i += 32) {
            uint64_t mine[W];
#pragma unroll
            for (int w = 0; w < W; w++) mine[w] = skey[(size_t)i * W + w];
            uint32_t rank = 0;
            for (uint32_t m = 0; m < n_got; m++) {
                bool less = false, decided = false; /* lexicographic, word 0 first */
#pragma unroll
                for (int w = 0; w < W; w++) {
                    const uint64_t x = skey[(size_t)m * W + w];
                    if (!decided && x != mine[w]) { less = x < mine[w]; decided = true; }
                }
                rank += less;
            }
            uint64_t key[W], km[W];
#pragma unroll
            for (int w = 0; w < W; w++) key[w] = bft_bswap64(mine[w]);
            bft_emit_kmer<W>(base, key, shift_bits, km);
            const uint64_t dst = out + rank;
            for (int w = 0; w < W; w++) kmers[dst * W + w] = km[w];
            if (cls_out) cls_out[dst] = scls[i];
            if (loc2vid) loc2vid[sloc[i]] = (uint32_t)dst;
        }
        __syncwarp();
    }
}

/* the Nodes' own UC lines: whole remainders below the Node's path (one thread per Node, <= 255 lines each), written in
 * the order the reference's UC stores them (uc_rank) */
template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_extract_uc_kmers(const bft_view_t v, size_t n_nodes, uint64_t* __restrict__ kmers,
                                                              uint32_t* __restrict__ cls_out, uint32_t* __restrict__ loc2vid) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t nid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; nid < n_nodes; nid += stride) {
        const bft_node_t nd = v.nodes[nid];
        if (!nd.uc_n) continue;
        const bft_path_t path = v.node_path[nid];
        const uint64_t out = ((uint64_t)path.uc_out_hi << 32) | path.uc_out_lo;
        uint64_t base[W];
#pragma unroll
        for (int w = 0; w < W; w++) base[w] = path.acc[w];
        for (uint32_t i = 0; i < nd.uc_n; i++) {
            uint64_t key[W], km[W];
            for (int w = 0; w < W; w++) key[w] = v.uckeys[((size_t)nd.uc_begin + i) * W + w];
            if (path.depth == 0) { for (int w = 0; w < W; w++) km[w] = key[w]; }
            else bft_emit_kmer<W>(base, key, BFT_PREFIX_BITS * (int)path.depth, km);
            const uint64_t dst = out + v.uc_rank[nd.uc_begin + i]; /* the line's place in the UC as the reference stores it */
            for (int w = 0; w < W; w++) kmers[dst * W + w] = km[w];
            if (cls_out) cls_out[dst] = v.uccls[nd.uc_begin + i];
            if (loc2vid) loc2vid[v.loc_uc + nd.uc_begin + i] = (uint32_t)dst;
        }
    }
}

/* ---- a13/a14: branching / neighbours ---------------------------------------------------------------------------
 * isBranchingRight / isBranchingLeft (src/branchingNode.c:16-110, 240-413) count the successors (drop nuc 0, append c)
 * and predecessors (prepend c, drop the last nuc) of a k-mer that are present in the graph; get_neighbors
 * (src/bft.c:804-886) returns them. Eight look-ups per query, most of which miss. Two phases per warp:
 *
 *   filter   one LANE per query (two queries per lane per round). The four successors of a k-mer share the middle k-2
 *            nucleotides, and so do its four predecessors, hence one block of the stored-k-mer filter each (bft_arena.h):
 *            two hashes and two 32-byte L2 loads answer all eight "certainly absent?" questions of a query.
 *   walk     the neighbours that survive (1.5 of 8 on the 100-genome pan-genome) are compacted across the warp into a
 *            shared-memory queue and walked 32 at a time with every lane busy — in the first version of this kernel each
 *            of 8 lanes per query walked its own neighbour, and a warp executed the whole walk for the 19 % of its lanes
 *            that needed it (469 warp instructions per 4 queries, issue-bound).
 * MODE 0: successor / predecessor counts, branching total, optional neighbour classes (k_query_branching).
 * MODE 1: vertex ids of the neighbours through loc2vid (k_graph_adjacency, bft_graph.cuh). */
#define BFT_NBR_Q 2 /* queries per lane per round */

template <int W>
__device__ __forceinline__ void bft_neighbor_kmer(const uint64_t* x, int k, int sub, uint64_t* y) {
    /* sub 0-3: successor with last nucleotide sub; sub 4-7: predecessor with first nucleotide sub - 4 */
    const uint32_t c = sub & 3;
    if (sub < 4) {
        bft_shr<W>(x, 2, y);
        const int top = 2 * (k - 1);
#pragma unroll
        for (int w = 0; w < W; w++) {
            uint64_t add = (uint64_t)c << (top & 63);
            if (W > 1) BFT_OPAQUE(add);
            y[w] |= ((top >> 6) == w) ? add : 0ULL;
        }
    } else {
        bft_shl<W>(x, 2, y);
        y[0] |= c;
#pragma unroll
        for (int w = 0; w < W; w++) y[w] &= bft_word_mask(2 * k, w);
    }
}

/* the 8-bit survivor mask of one query: bit sub set = neighbour sub may be stored */
template <int W>
__device__ __forceinline__ uint32_t bft_neighbor_filter(const bft_view_t& v, const uint64_t* x, bool filter_succ, bool filter_pred) {
    const int k = v.k;
    uint32_t mask = 0;
    uint64_t mid[W], blk[4];
    /* successors: middle = nucleotides 2..k-1 of x; ends = (nucleotide 1 of x, c) */
    if (filter_succ) {
        bft_shr<W>(x, 4, mid);
        const uint64_t h = bft_kf_mix_mid(mid, W);
        const uint32_t first = (uint32_t)(x[0] >> 2) & 3u;
        bft_kf_load(&v, bft_kf_finish(h, first, v.kf_blocks).block, blk);
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) mask |= (uint32_t)bft_kf_bits(blk, bft_kf_finish(h, first | (c << 2), v.kf_blocks)) << c;
    } else mask |= 0x0fu;
    /* predecessors: middle = nucleotides 0..k-3 of x; ends = (c, nucleotide k-2 of x) */
    if (filter_pred) {
        uint32_t last = 0;
        const int lpos = 2 * (k - 2);
#pragma unroll
        for (int w = 0; w < W; w++) {
            mid[w] = x[w] & bft_word_mask(lpos, w);
            uint64_t cand = (x[w] >> (lpos & 63)) & 3ULL;
            if (W > 1) BFT_OPAQUE(cand);
            last |= ((lpos >> 6) == w) ? (uint32_t)cand : 0u;
        }
        const uint64_t h = bft_kf_mix_mid(mid, W);
        bft_kf_load(&v, bft_kf_finish(h, last << 2, v.kf_blocks).block, blk);
#pragma unroll
        for (uint32_t c = 0; c < 4; c++) mask |= (uint32_t)bft_kf_bits(blk, bft_kf_finish(h, c | (last << 2), v.kf_blocks)) << (4 + c);
    } else mask |= 0xf0u;
    return mask;
}

template <int W, int MODE>
__device__ __forceinline__ void bft_neighbors_core(const bft_view_t& v, const uint64_t* __restrict__ kmers, size_t n, int ref_quirks,
                                                   uint8_t* __restrict__ succ, uint8_t* __restrict__ pred, unsigned long long* __restrict__ n_branching,
                                                   uint32_t* __restrict__ nbr_out, const uint32_t* __restrict__ loc2vid) {
    constexpr int QW = 32 * BFT_NBR_Q; /* queries per warp per round */
    __shared__ uint64_t s_x[BFT_TPB / 32][QW * W];
    __shared__ uint16_t s_queue[BFT_TPB / 32][QW * 8];
    __shared__ uint32_t s_cnt[BFT_TPB / 32][QW]; /* successors found | predecessors found << 16 */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* const xs = s_x[warp];
    uint16_t* const queue = s_queue[warp];
    uint32_t* const cnt = s_cnt[warp];
    const size_t n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t n_rounds = (n + QW - 1) / QW;
    /* the successor look-ups may have to reproduce the reference's leaf-level quirk (bft_node_probe_ex): they can be
     * pre-filtered only if the quirk cannot trigger in this trie */
    const bool have_filter = v.kf_blocks != 0;
    const bool quirk = MODE == 0 && ref_quirks;
    const bool filter_succ = have_filter && (!quirk || v.kf_quirk_safe), filter_pred = have_filter;
    unsigned long long local = 0;
    for (size_t r = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rounds; r += n_warps) {
        const size_t base = r * QW;
        uint32_t masks[BFT_NBR_Q], total_mine = 0;
#pragma unroll
        for (int j = 0; j < BFT_NBR_Q; j++) {
            const int slot = j * 32 + lane;
            const size_t q = base + slot;
            masks[j] = 0;
            if (q < n) {
                uint64_t x[W];
                bft_load_kmer<W>(kmers, q, x);
#pragma unroll
                for (int w = 0; w < W; w++) xs[slot * W + w] = x[w];
                masks[j] = bft_neighbor_filter<W>(v, x, filter_succ, filter_pred);
                if (nbr_out) { /* absent unless the walk says otherwise */
                    uint4* o = (uint4*)(nbr_out + q * 8);
                    o[0] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                    o[1] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                }
            }
            cnt[slot] = 0;
            total_mine += __popc(masks[j]);
        }
        /* queue slots: exclusive scan of the survivor counts over the warp */
        uint32_t incl = total_mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t at = incl - total_mine;
#pragma unroll
        for (int j = 0; j < BFT_NBR_Q; j++) {
            uint32_t m = masks[j];
            while (m) {
                const int sub = __ffs(m) - 1;
                m &= m - 1;
                queue[at++] = (uint16_t)(((j * 32 + lane) << 3) | sub);
            }
        }
        __syncwarp();
        /* walk: 32 surviving neighbours at a time */
        for (uint32_t i = lane; i < total; i += 32) {
            const uint32_t e = queue[i];
            const int slot = (int)(e >> 3), sub = (int)(e & 7u);
            uint64_t x[W], y[W];
#pragma unroll
            for (int w = 0; w < W; w++) x[w] = xs[slot * W + w];
            bft_neighbor_kmer<W>(x, v.k, sub, y);
            uint32_t loc = 0;
            const uint32_t cls = bft_lookup_loc(&v, y, W, ((quirk && sub < 4) ? BFT_LK_SUCC_QUIRK : 0) | BFT_LK_NO_FILTER |
                                                ((MODE == 0 && !nbr_out) ? BFT_LK_PRESENCE : 0), (uint32_t*)0, MODE == 1 ? &loc : (uint32_t*)0);
            if (cls != BFT_CLS_NONE) {
                const int pos = sub < 4 ? 4 + sub : sub - 4; /* get_neighbors order: 0-3 predecessors, 4-7 successors */
                if (MODE == 0) {
                    atomicAdd(&cnt[slot], sub < 4 ? 1u : 0x10000u);
                    if (nbr_out) nbr_out[(base + slot) * 8 + pos] = cls;
                } else {
                    nbr_out[(base + slot) * 8 + pos] = __ldg(loc2vid + loc);
                }
            }
        }
        __syncwarp();
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < BFT_NBR_Q; j++) {
                const int slot = j * 32 + lane;
                const size_t q = base + slot;
                if (q < n) {
                    const uint32_t c = cnt[slot];
                    const uint32_t ns = c & 0xffffu, np = c >> 16;
                    if (succ) succ[q] = (uint8_t)ns;
                    if (pred) pred[q] = (uint8_t)np;
                    local += (ns > 1) || (np > 1); /* isBranchingRight > 1, else isBranchingLeft > 1 (src/file_io.c:971-976) */
                }
            }
        }
        __syncwarp();
    }
    if (MODE == 0 && n_branching) {
        for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
        if (lane == 0 && local) atomicAdd(n_branching, local);
    }
}

template <int W>
__global__ void __launch_bounds__(BFT_TPB) k_query_branching(const bft_view_t v, const uint64_t* __restrict__ kmers, size_t n,
                                                             uint8_t* __restrict__ succ, uint8_t* __restrict__ pred,
                                                             unsigned long long* __restrict__ n_branching,
                                                             uint32_t* __restrict__ nbr_cls, int ref_quirks) {
    bft_neighbors_core<W, 0>(v, kmers, n, ref_quirks, succ, pred, n_branching, nbr_cls, (const uint32_t*)0);
}

/* ---- a1/a2/a12: sequences ----------------------------------------------------------------------------------
 * One warp per sequence. The warp streams the sequence in tiles: 32 characters per step are classified and
 * 2-bit-encoded by the 32 lanes and packed with warp ballots into shared-memory bit planes (codes + character-class
 * masks). Every k-mer window is then a funnel shift out of the packed plane; the reverse complement is a
 * 2-bit-group reversal of the complemented word (reverse_complement, src/fasta.c:387-440) and the canonical pick
 * compares the nucleotide-lexicographic keys (strcmp rule, src/bft.c:1287-1293). Windows sharing a colour class are
 * merged with __match_any_sync before the per-genome counters in shared memory are bumped. */
#define BFT_SEQ_TILE 512                       /* window start positions per tile */
#define BFT_SEQ_SPAN (BFT_SEQ_TILE + 128)      /* characters held per tile (k <= 126 overlap, rounded to 64) */
#define BFT_SEQ_WARPS 4

/* shared memory one warp of k_query_sequences needs (codes + 4 masks + per-genome counters; up to 32 genomes are counted in registers) */
__host__ __device__ inline size_t bft_seq_smem_per_warp(int n_genomes) {
    const size_t n_code_words = BFT_SEQ_SPAN / 32 + 2, n_mask_words = BFT_SEQ_SPAN / 64 + 2;
    const size_t b = n_code_words * 8 + 4 * n_mask_words * 8 + (n_genomes > 32 ? (size_t)((n_genomes + 31) & ~31) * 4 : 0);
    return (b + 15) & ~(size_t)15;
}

/* character classes (see bft_b200.h, bft_b200_query_sequences) */
#define BFT_CH_ACGT 0   /* A C G T U, either case */
#define BFT_CH_IUPAC 1  /* IUPAC letter accepted by reverse_complement and matched by is_substring_IUPAC */
#define BFT_CH_DOTDASH 2 /* '.' '-' : matched by is_substring_IUPAC (src/fasta.c:361) but rejected by reverse_complement */
#define BFT_CH_OTHER 3  /* anything else */

__device__ __forceinline__ void bft_classify_char(unsigned char ch, uint32_t& code, uint32_t& cls, uint32_t& nonplain) {
    code = 0; cls = BFT_CH_OTHER; nonplain = 0;
    const unsigned char up = ch & 0xdf;
    const bool letter = (up >= 'A' && up <= 'Z') && (ch & 0x40);
    if (letter) {
        const bool lower = (ch & 0x20) != 0;
        switch (up) {
            case 'A': code = 0; cls = BFT_CH_ACGT; nonplain = lower; break;
            case 'C': code = 1; cls = BFT_CH_ACGT; nonplain = lower; break;
            case 'G': code = 2; cls = BFT_CH_ACGT; nonplain = lower; break;
            case 'T': code = 3; cls = BFT_CH_ACGT; nonplain = lower; break;
            case 'U': code = 3; cls = BFT_CH_ACGT; nonplain = 1; break;
            case 'R': case 'Y': case 'S': case 'W': case 'K': case 'M': case 'B': case 'D': case 'H': case 'V': case 'N':
                cls = BFT_CH_IUPAC; break;
            default: break;
        }
    } else if (ch == '.' || ch == '-') {
        cls = BFT_CH_DOTDASH;
    }
}

__device__ __forceinline__ uint64_t bft_spread32(uint32_t x) { /* bit i -> bit 2i */
    uint64_t v = x;
    v = (v | (v << 16)) & 0x0000ffff0000ffffULL;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffULL;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0fULL;
    v = (v | (v << 2)) & 0x3333333333333333ULL;
    v = (v | (v << 1)) & 0x5555555555555555ULL;
    return v;
}

__device__ __forceinline__ uint64_t bft_rev2_64(uint64_t x) { /* reverse the 32 2-bit groups of a word */
    x = __brevll(x);
    return ((x & 0xaaaaaaaaaaaaaaaaULL) >> 1) | ((x & 0x5555555555555555ULL) << 1);
}

/* bits [pos, pos+len) of a little-endian bit plane of uint64 words (len <= 64) */
__device__ __forceinline__ uint64_t bft_extract_bits(const uint64_t* plane, int pos, int len) {
    const int w = pos >> 6, sh = pos & 63;
    uint64_t x = plane[w] >> sh;
    if (sh && sh + len > 64) x |= plane[w + 1] << (64 - sh);
    return len < 64 ? x & ((1ULL << len) - 1ULL) : x;
}

/* any bit set in [pos, pos+len) of a bit plane, len <= 128 */
__device__ __forceinline__ bool bft_any_bits(const uint64_t* plane, int pos, int len) {
    uint64_t x = bft_extract_bits(plane, pos, len < 64 ? len : 64);
    if (len > 64) x |= bft_extract_bits(plane, pos + 64, len - 64);
    return x != 0;
}

__device__ __forceinline__ char bft_rc_char(char c) { /* reverse_complement on ACGTU (src/fasta.c:396-405) */
    switch (c) {
        case 'a': return 't'; case 'A': return 'T';
        case 'c': return 'g'; case 'C': return 'G';
        case 'g': return 'c'; case 'G': return 'C';
        case 't': case 'u': return 'a';
        default: return 'A'; /* 'T', 'U' */
    }
}

/* 32 x 32 bit-matrix transpose across a warp: lane r holds row r, afterwards lane c holds column c (bit r of the result = bit c of
 * lane r's input). Five butterfly steps, one shuffle each. */
__device__ __forceinline__ uint32_t bft_warp_transpose32(uint32_t x, const int lane) {
#define BFT_TR_STEP(J_, M_)                                                                  \
    {                                                                                        \
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, J_);                              \
        x = (lane & J_) ? ((x & ~M_) | ((y & ~M_) >> J_)) : ((x & M_) | ((y & M_) << J_));   \
    }
    BFT_TR_STEP(16, 0x0000ffffu)
    BFT_TR_STEP(8, 0x00ff00ffu)
    BFT_TR_STEP(4, 0x0f0f0f0fu)
    BFT_TR_STEP(2, 0x33333333u)
    BFT_TR_STEP(1, 0x55555555u)
#undef BFT_TR_STEP
    return x;
}

/* G32: at most 32 genomes (one row word), counted in registers; otherwise the class-merge + shared-memory counters */
template <int W, bool G32>
__global__ void __launch_bounds__(32 * BFT_SEQ_WARPS) k_query_sequences(const bft_view_t v, const char* __restrict__ chars,
                                                                        const uint64_t* __restrict__ offs, size_t n_seq, double threshold,
                                                                        int canonical, const uint32_t* __restrict__ class_rows, int rw,
                                                                        int n_genomes, uint32_t* __restrict__ rows,
                                                                        uint8_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = v.k;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    /* per-warp shared layout */
    const int n_code_words = BFT_SEQ_SPAN / 32 + 2; /* 2 bits per base, +slack for the funnel */
    const int n_mask_words = BFT_SEQ_SPAN / 64 + 2;
    unsigned char* base = smem_raw + (size_t)warp * bft_seq_smem_per_warp(n_genomes);
    uint64_t* codes = (uint64_t*)base;
    uint64_t* m_iupac = codes + n_code_words;   /* class IUPAC or DOTDASH */
    uint64_t* m_rcbad = m_iupac + n_mask_words; /* class DOTDASH or OTHER */
    uint64_t* m_other = m_rcbad + n_mask_words; /* class OTHER */
    uint64_t* m_nonpl = m_other + n_mask_words; /* lowercase acgt / U / u */
    uint32_t* counts = (uint32_t*)(m_nonpl + n_mask_words); /* !G32 only */
    const int n_counts = G32 ? 0 : ((n_genomes + 31) & ~31);

    const size_t warp_stride = (size_t)gridDim.x * BFT_SEQ_WARPS;
    for (size_t s = (size_t)blockIdx.x * BFT_SEQ_WARPS + warp; s < n_seq; s += warp_stride) {
        const uint64_t o0 = offs[s], o1 = offs[s + 1];
        const long long len = (long long)(o1 - o0);
        const long long n_win = len - k + 1;
        for (int g = lane; g < n_counts; g += 32) counts[g] = 0;
        uint32_t cnt_reg = 0; /* G32: lane g counts genome g in a register */
        int bad = 0;
        /* The stored-k-mer filter pays for look-ups that miss. A read of an organism in the graph hits on nearly every
         * window, and this kernel is issue-bound, so the filter's hash and L2 load would be pure overhead there: the first
         * 32 windows of a sequence are looked up without it, and it is switched on for the rest only if most of them missed. */
        int lk_flags = BFT_LK_NO_FILTER;
        /* threshold: count >= ceil(n_win * threshold) (src/bft.c:1279, 1327-1339). The reference stops scanning a sequence as
         * soon as no genome can reach it any more — `found + windows left < need` (src/bft.c:1319) — and so does the warp, at
         * its own granularity of 32 windows: the answer is the all-zero row either way (a genome's count never exceeds the
         * number of windows found), and the warp stops no earlier than the reference does. */
        /* found + left < need  <=>  windows missed so far > n_win - need; 32-bit, saturating (a budget that does not fit never trips) */
        unsigned int miss_budget = 0xffffffffu, missed = 0;
        if (n_win > 0) {
            const long long slack = n_win - (long long)ceil((double)n_win * threshold);
            if (slack < 0xffffffffLL) miss_budget = (unsigned int)slack;
        }
        __syncwarp();
        for (long long t0 = 0; t0 < n_win && missed <= miss_budget; t0 += BFT_SEQ_TILE) {
            const int n_here = (int)(n_win - t0 < BFT_SEQ_TILE ? n_win - t0 : BFT_SEQ_TILE); /* windows in this tile */
            const int n_chars = n_here + k - 1;
            /* stage + encode: 32 characters per step. Upper-case A/C/G/T — all there is in ordinary reads — take a branch-free
             * path: code = ((ch >> 1) ^ (ch >> 2)) & 3; the 2-bit codes of the 32 lanes become one 64-bit word of the plane with
             * two warp-wide OR reductions (REDUX; lanes 0-15 fill the low half, 16-31 the high half). Anything else is classified
             * in full: the first such character of a tile clears the four mask planes, blocks with such characters write their
             * bits, and a tile without any skips every mask test below (tile_special == 0). */
            const char* const cp = chars + o0 + (uint64_t)t0;
            const int n_blocks = (n_chars + 31) >> 5;
            uint32_t tile_special = 0;
            for (int blk = 0; blk < n_blocks; blk++) {
                const int ci = (blk << 5) + lane;
                uint32_t code = 0, cls = BFT_CH_ACGT, nonplain = 0;
                if (ci < n_chars) {
                    const unsigned char ch = (unsigned char)cp[ci];
                    const uint32_t d = (uint32_t)ch - 'A';
                    if (d < 26u && ((0x00080045u >> d) & 1u)) code = ((ch >> 1) ^ (ch >> 2)) & 3u; /* A, C, G, T */
                    else bft_classify_char(ch, code, cls, nonplain);
                }
                const uint32_t sh = (uint32_t)(lane & 15) << 1;
                const uint32_t lo = __reduce_or_sync(0xffffffffu, lane < 16 ? code << sh : 0u);
                const uint32_t hi = __reduce_or_sync(0xffffffffu, lane < 16 ? 0u : code << sh);
                const uint32_t special = __ballot_sync(0xffffffffu, cls != BFT_CH_ACGT || nonplain);
                if (special) {
                    if (!tile_special) { /* the four planes are contiguous */
                        for (int i = lane; i < 8 * n_mask_words; i += 32) ((uint32_t*)m_iupac)[i] = 0u;
                        tile_special = 1;
                        __syncwarp();
                    }
                    const uint32_t bi = __ballot_sync(0xffffffffu, cls == BFT_CH_IUPAC || cls == BFT_CH_DOTDASH);
                    const uint32_t br = __ballot_sync(0xffffffffu, cls == BFT_CH_DOTDASH || cls == BFT_CH_OTHER);
                    const uint32_t bo = __ballot_sync(0xffffffffu, cls == BFT_CH_OTHER);
                    const uint32_t bn = __ballot_sync(0xffffffffu, nonplain);
                    if (lane == 1) {
                        ((uint32_t*)m_iupac)[blk] = bi;
                        ((uint32_t*)m_rcbad)[blk] = br;
                        ((uint32_t*)m_other)[blk] = bo;
                        ((uint32_t*)m_nonpl)[blk] = bn;
                    }
                }
                if (lane == 0) codes[blk] = (uint64_t)lo | ((uint64_t)hi << 32);
            }
            if (lane < 2 && n_blocks + lane < n_code_words) codes[n_blocks + lane] = 0; /* nothing reads past the data; kept defined */
            __syncwarp();
            /* windows: lane handles j = lane, lane+32, ... (whole warp iterates together for the collectives) */
            for (int j0 = 0; j0 < n_here; j0 += 32) {
                const int j = j0 + lane;
                uint32_t cls = BFT_CLS_NONE;
                if (j < n_here) {
                    bool wi = false, wr = false, wo = false;
                    if (tile_special) {
                        wi = bft_any_bits(m_iupac, j, k);
                        wr = bft_any_bits(m_rcbad, j, k);
                        wo = bft_any_bits(m_other, j, k);
                    }
                    int skip = 0;
                    if (canonical) {
                        if (wr) bad = 1;
                        if (wr || wi) skip = 1;
                    } else {
                        if (wi) skip = 1;
                        else if (wo) { bad = 1; skip = 1; }
                    }
                    if (!skip) {
                        uint64_t x[W];
#pragma unroll
                        for (int w = 0; w < W; w++) {
                            const int len = 2 * k - 64 * w;
                            x[w] = len <= 0 ? 0ULL : bft_extract_bits(codes, 2 * j + 64 * w, len < 64 ? len : 64);
                        }
                        if (canonical) {
                            /* key(fwd) = R(x) (2-bit-group reversal), key(rc) = ~x, rc = ~R(x), all within 2k bits.
                             * strcmp(fwd, rc) >= 0 -> rc (src/bft.c:1287-1293) */
                            uint64_t r[W], rx[W], nx[W];
#pragma unroll
                            for (int w = 0; w < W; w++) r[w] = bft_rev2_64(x[W - 1 - w]);
                            bft_shr<W>(r, 64 * W - 2 * k, rx);
#pragma unroll
                            for (int w = 0; w < W; w++) nx[w] = ~x[w] & bft_word_mask(2 * k, w);
                            int use_rc = bft_ge<W>(rx, nx);
                            if (tile_special && bft_any_bits(m_nonpl, j, k)) { /* mixed case / U: ASCII order decides */
                                use_rc = 1;
                                for (int i = 0; i < k; i++) {
                                    const char a = cp[j + i], b = bft_rc_char(cp[j + k - 1 - i]);
                                    if (a != b) { use_rc = (unsigned char)a > (unsigned char)b; break; }
                                }
                            }
                            if (use_rc) {
#pragma unroll
                                for (int w = 0; w < W; w++) x[w] = ~rx[w] & bft_word_mask(2 * k, w);
                            }
                        }
                        cls = bft_lookup_ex(&v, x, W, lk_flags, (uint32_t*)0);
                    }
                }
                if (t0 == 0 && j0 == 0) {
                    const uint32_t looked = __ballot_sync(0xffffffffu, j < n_here);
                    const uint32_t found = __ballot_sync(0xffffffffu, cls != BFT_CLS_NONE);
                    if (2 * __popc(found) < __popc(looked)) lk_flags = 0;
                }
                const uint32_t active = __ballot_sync(0xffffffffu, cls != BFT_CLS_NONE);
                {
                    const unsigned int m = (unsigned int)min(32, n_here - j0) - (unsigned int)__popc(active);
                    missed = missed + m < missed ? 0xffffffffu : missed + m;
                    if (missed > miss_budget) break;
                }
                if (G32) {
                    /* up to 32 genomes: every lane fetches the row of its own window (a table this small sits in L1), the warp
                     * transposes the 32 rows as a bit matrix, and lane g adds the set bits of column g — the windows of this step
                     * that carry genome g — to a register. No shared memory, and the cost does not depend on how many distinct
                     * colour classes the 32 windows have. */
                    if (active) {
                        const uint32_t row = cls != BFT_CLS_NONE ? bft_ld_row(class_rows + cls) : 0u;
                        cnt_reg += (uint32_t)__popc(bft_warp_transpose32(row, lane));
                    }
                    continue;
                }
                /* merge windows with the same colour class, then bump the per-genome counters warp-wide */
                if (active) {
                    uint32_t grp = 0;
                    if (cls != BFT_CLS_NONE) grp = __match_any_sync(active, cls);
                    uint32_t leaders = __ballot_sync(0xffffffffu, cls != BFT_CLS_NONE && (__ffs(grp) - 1) == lane);
                    while (leaders) {
                        const int src = __ffs(leaders) - 1;
                        leaders &= leaders - 1;
                        const uint32_t c = __shfl_sync(0xffffffffu, cls, src);
                        const uint32_t add = __popc(__shfl_sync(0xffffffffu, grp, src));
                        const uint32_t* row = class_rows + (size_t)c * rw;
                        for (int w = 0; w < rw; w++) {
                            const uint32_t bits = bft_ld_row(row + w);
                            if ((bits >> lane) & 1u) counts[w * 32 + lane] += add;
                        }
                    }
                }
            }
            __syncwarp();
        }
        bad = __any_sync(0xffffffffu, bad);
        long long need = 0;
        if (n_win > 0) need = (long long)ceil((double)n_win * threshold);
        const bool unreachable = missed > miss_budget;
        for (int w = 0; w < rw; w++) {
            const uint32_t cnt = G32 ? cnt_reg : counts[w * 32 + lane];
            const uint32_t bits = __ballot_sync(0xffffffffu, !unreachable && n_win > 0 && cnt > 0 && (long long)cnt >= need);
            if (lane == 0) rows[s * (size_t)rw + w] = bits;
        }
        if (lane == 0 && status) status[s] = bad ? 2 : (n_win <= 0 ? 1 : 0);
        __syncwarp();
    }
}

#endif /* BFT_KERNELS_CUH */
