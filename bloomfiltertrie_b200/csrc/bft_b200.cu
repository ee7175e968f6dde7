/* bft_b200.cu — the C-ABI (include/bft_b200.h) over the sm_100a kernels (bft_kernels.cuh).
 * Host logic only: context life cycle, arena upload, launch configuration, chunked host<->device pipelines.
 * There is deliberately no CPU query path here: every query entry point launches kernels or fails. */
#include "../../include/bft_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "bft_flatten.h"
#include "bft_io.h"
#include "bft_kernels.cuh"
#include "bft_graph.cuh"
#include <cub/device/device_scan.cuh>

static __thread char g_err[512];

static int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return set_err(BFT_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

#define BFT_CHUNK_KMERS ((size_t)1 << 22)
#define BFT_CHUNK_SEQ_CHARS ((size_t)1 << 25) /* 32 MB of characters per chunk: small enough for the two slots to overlap copies and kernels on a 1 M-read batch */
#define BFT_CHUNK_SEQS ((size_t)1 << 18)

#define BFT_N_SLOTS 4

typedef struct {
    void* d_in;        size_t cap_in;
    uint64_t* d_offs;  size_t cap_offs;
    uint64_t* d_kmers; size_t cap_kmers; /* packed k-mers when the input was ASCII */
    uint8_t* d_u8a;    size_t cap_u8a;
    uint8_t* d_u8b;    size_t cap_u8b;
    uint32_t* d_cls;   size_t cap_cls;
    uint32_t* d_rows;  size_t cap_rows;
    uint32_t* d_tile;  size_t cap_tile;  /* compact form: per-tile hit counts, their prefix sums, the chunk total */
    uint32_t* h_total;                   /* pinned: where the chunk total lands */
} slot_t;

struct bft_b200_ctx {
    int device, sm_count;
    cudaStream_t streams[BFT_N_SLOTS]; /* 0 and 1 serve every host pipeline; 2 and 3 only the compact form's deeper one */
    int k, W, G, rw;
    char** names;
    bft_b200_stats stats;
    /* device arena */
    void* d_arena[19];
    size_t n_pref, n_nodes;
    bft_view_t dview;
    bft_pools_t dpools;
    void* d_pool[4];
    void* d_hot;              /* one allocation: rootdir | class_rows | kfilter — the tables every lookup touches, kept in L2 */
    void* d_rootkf;           /* fused root directory + filter (bft_arena.h), the plain look-ups' first stop; NULL: none */
    void* d_deep[3];          /* collapsed subtrees (bft_arena.h): rootdir_fast, dbuckets, dslotcls; NULL: none */
    size_t hot_bytes;
    uint64_t n_loc;           /* storage locations (bft_view_t), counted in 64 bits */
    cudaMemPool_t pool;       /* private pool of the traversal scratch (the device's default pool is left alone) */
    uint32_t* d_class_rows;   /* inside d_hot */
    uint32_t* d_class_counts;
    uint32_t* h_class_rows;
    uint32_t* h_class_counts;
    size_t n_classes;
    unsigned long long* d_counter;
    slot_t slot[BFT_N_SLOTS];
    uint64_t launches;
    size_t seq_smem;
    int ref_quirks;
    /* device graph (bft_graph.cuh), built on first use */
    struct {
        int ready;
        size_t n;            /* vertices = stored k-mers */
        uint64_t* d_vk;      /* n * W: the k-mer of each vertex, in enumeration order */
        uint32_t* d_vcls;    /* n: its colour class */
        uint32_t* d_adj;     /* n * 8: predecessors 0-3, successors 4-7 (vertex ids or BFT_V_NONE) */
        uint32_t* d_loc2vid; /* storage location (bft_view_t) -> vertex id */
        size_t bytes;
    } graph;
};

extern "C" const char* bft_b200_last_error(void) { return g_err; }

static int drain_ret(bft_b200_ctx* c, int rc) {
    char keep[sizeof g_err];
    memcpy(keep, g_err, sizeof keep); /* the message of the failure that brought us here */
    for (int s = 0; s < BFT_N_SLOTS; s++)
        if (c->streams[s]) cudaStreamSynchronize(c->streams[s]);
    (void)cudaGetLastError();
    memcpy(g_err, keep, sizeof keep);
    return rc;
}

/* instantiate a launch for the arena's key width (W = 1, 2 or 4 words) */
#define BFT_BY_W(w_, LAUNCH)                 \
    do {                                     \
        if ((w_) == 1) { LAUNCH(1); }        \
        else if ((w_) == 2) { LAUNCH(2); }   \
        else { LAUNCH(4); }                  \
    } while (0)

static int ensure(void** p, size_t* cap, size_t need) {
    if (need <= *cap && *p) return 0;
    if (*p) cudaFree(*p);
    *p = NULL;
    *cap = 0;
    size_t n = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(p, n);
    if (e != cudaSuccess) return set_err(BFT_B200_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
    *cap = n;
    return 0;
}
#define ENSURE(ptr, cap, need)                                         \
    do {                                                               \
        int r_ = ensure((void**)&(ptr), &(cap), (need));               \
        if (r_) return r_;                                             \
    } while (0)

/* Error exits of the chunked host pipelines: copies of earlier chunks may still be reading the caller's input or writing
 * its output buffers, so every stream is drained before the error is returned (the caller is free to release its buffers
 * the moment the call comes back). CKP / ENSUREP are CK / ENSURE with that drain. */
static int drain_ret(bft_b200_ctx* c, int rc);
#define CKP(call)                                                                                                         \
    do {                                                                                                                 \
        cudaError_t e_ = (call);                                                                                         \
        if (e_ != cudaSuccess)                                                                                           \
            return drain_ret(c, set_err(BFT_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)
#define ENSUREP(ptr, cap, need)                                        \
    do {                                                               \
        int r_ = ensure((void**)&(ptr), &(cap), (need));               \
        if (r_) return drain_ret(c, r_);                               \
    } while (0)

static int upload(void** dst, const void* src, size_t bytes) {
    size_t n = bytes ? bytes : 16;
    cudaError_t e = cudaMalloc(dst, n + 32); /* slack for vector loads at the tail */
    if (e != cudaSuccess) return set_err(BFT_B200_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
    e = cudaMemset(*dst, 0, n + 32);
    if (e == cudaSuccess && bytes) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return set_err(BFT_B200_ERR_CUDA, "arena upload failed: %s", cudaGetErrorString(e));
    return 0;
}

static int grid_for(const bft_b200_ctx* c, size_t items, int per_block) {
    size_t blocks = (items + (size_t)per_block - 1) / (size_t)per_block;
    size_t cap = (size_t)c->sm_count * 8; /* persistent, grid-stride: whole multiples of the SM count */
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

extern "C" void bft_b200_close(bft_b200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 19; i++) if (c->d_arena[i]) cudaFree(c->d_arena[i]);
    for (int i = 0; i < 4; i++) if (c->d_pool[i]) cudaFree(c->d_pool[i]);
    if (c->d_hot) cudaFree(c->d_hot);
    if (c->d_rootkf) cudaFree(c->d_rootkf);
    for (int i = 0; i < 3; i++) if (c->d_deep[i]) cudaFree(c->d_deep[i]);
    if (c->d_class_counts) cudaFree(c->d_class_counts);
    if (c->d_counter) cudaFree(c->d_counter);
    bft_b200_graph_release(c);
    if (c->pool) cudaMemPoolDestroy(c->pool);
    for (int s = 0; s < BFT_N_SLOTS; s++) {
        slot_t* sl = &c->slot[s];
        if (sl->d_in) cudaFree(sl->d_in);
        if (sl->d_offs) cudaFree(sl->d_offs);
        if (sl->d_kmers) cudaFree(sl->d_kmers);
        if (sl->d_u8a) cudaFree(sl->d_u8a);
        if (sl->d_u8b) cudaFree(sl->d_u8b);
        if (sl->d_cls) cudaFree(sl->d_cls);
        if (sl->d_rows) cudaFree(sl->d_rows);
        if (sl->d_tile) cudaFree(sl->d_tile);
        if (sl->h_total) cudaFreeHost(sl->h_total);
        if (c->streams[s]) cudaStreamDestroy(c->streams[s]);
    }
    if (c->names) {
        for (int i = 0; i < c->G; i++) free(c->names[i]);
        free(c->names);
    }
    free(c->h_class_rows);
    free(c->h_class_counts);
    free(c);
}

static int enqueue_extract(bft_b200_ctx* c, uint64_t* d_kmers, uint32_t* d_cls, uint32_t* d_loc2vid);

/* The two stored-k-mer filters (bft_arena.h): enumerate the arena's k-mers on the device ONCE, set their bits in kfilter (when
 * n_blocks != 0), and build the fused root directory + filter when the k-mers are spread evenly enough over the 9-nt prefixes
 * for 192 * S bits per prefix to filter anything. Both are optional accelerators: on any shortage of memory the engine runs
 * without them. rkf_bytes_out receives the size of the fused table (0: not built). */
/* Collapsed subtrees (bft_arena.h, rootdir_fast / dbuckets): every root prefix whose suffixes live in a child Node gets one hashed
 * block with all the k-mers below it. d_k / d_cls: the enumeration (k-mers and classes). Optional: on a shortage of memory, or when
 * BFT_B200_NO_DEEP=1, the look-ups keep walking the Nodes. */
static int build_deep_blocks(bft_b200_ctx* c, const bft_entry_t* h_rootdir, const uint64_t* d_k, const uint32_t* d_cls, size_t n_kmers, size_t* bytes_out) {
    *bytes_out = 0;
    if (getenv("BFT_B200_NO_DEEP") && getenv("BFT_B200_NO_DEEP")[0] == '1') return 0;
    bool any = false;
    for (size_t p = 0; p < BFT_ROOTDIR_SIZE && !any; p++) any = (h_rootdir[p].b >> BFT_KIND_SHIFT) == BFT_KIND_NODE;
    if (!any) return 0;
    cudaStream_t st = c->streams[0];
    uint32_t* d_cnt = NULL;
    unsigned int* d_failed = NULL;
    uint32_t* h_cnt = (uint32_t*)malloc(BFT_ROOTDIR_SIZE * sizeof(uint32_t));
    bft_entry_t* h_fast = (bft_entry_t*)malloc(BFT_ROOTDIR_SIZE * sizeof(bft_entry_t));
    int rc = 0;
    bool ok = h_cnt && h_fast && cudaMalloc((void**)&d_cnt, BFT_ROOTDIR_SIZE * sizeof(uint32_t)) == cudaSuccess &&
              cudaMalloc((void**)&d_failed, sizeof(unsigned int)) == cudaSuccess;
    if (ok) {
        cudaMemsetAsync(d_cnt, 0, BFT_ROOTDIR_SIZE * sizeof(uint32_t), st);
        cudaMemsetAsync(d_failed, 0, sizeof(unsigned int), st);
#define BFT_L(W_) k_deep_count<W_><<<grid_for(c, n_kmers, BFT_TPB), BFT_TPB, 0, st>>>(c->dview.rootdir, d_k, n_kmers, d_cnt)
        BFT_BY_W(c->W, BFT_L);
#undef BFT_L
        c->launches++;
        ok = cudaMemcpyAsync(h_cnt, d_cnt, BFT_ROOTDIR_SIZE * sizeof(uint32_t), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
             cudaStreamSynchronize(st) == cudaSuccess;
    }
    size_t n_db = 0, n_deep = 0;
    if (ok) {
        /* block of a prefix: the smallest power of two of buckets that keeps the load at or below 1/2; the entry has four bits
         * for its log2, so a subtree of more than 2^16 k-mers stays a Node (and is walked) */
        /* test knob: BFT_B200_DEEP_TIGHT=1 sizes the blocks for a load of up to 1 (every slot may be taken), which makes the
         * linear probing — and its termination in a block without an empty slot — the common case instead of the rare one */
        const size_t per_bucket = (getenv("BFT_B200_DEEP_TIGHT") && getenv("BFT_B200_DEEP_TIGHT")[0] == '1') ? BFT_BUCKET_KEYS : BFT_BUCKET_KEYS / 2;
        memcpy(h_fast, h_rootdir, BFT_ROOTDIR_SIZE * sizeof(bft_entry_t));
        for (size_t p = 0; p < BFT_ROOTDIR_SIZE; p++) {
            if ((h_rootdir[p].b >> BFT_KIND_SHIFT) != BFT_KIND_NODE || h_cnt[p] == 0) continue;
            uint32_t lb = 0;
            while (lb < BFT_LB_MASK && ((size_t)1 << lb) * per_bucket < (size_t)h_cnt[p]) lb++;
            if (((size_t)1 << lb) * per_bucket < (size_t)h_cnt[p]) continue;
            if (n_db + ((size_t)1 << lb) >= 0xfffffff0u) { ok = false; break; }
            h_fast[p] = bft_mk_entry(BFT_KIND_DEEP, (uint32_t)n_db, h_cnt[p] > BFT_CNT_MASK ? BFT_CNT_MASK : h_cnt[p]);
            h_fast[p].b = (BFT_KIND_DEEP << BFT_KIND_SHIFT) | (lb << BFT_LB_SHIFT) | (h_fast[p].b & ((1u << BFT_LB_SHIFT) - 1u));
            n_db += (size_t)1 << lb;
            n_deep += h_cnt[p];
        }
        ok = ok && n_db > 0;
    }
    const size_t bk_bytes = n_db * (size_t)(BFT_BUCKET_KEYS * c->W) * 8, sc_bytes = c->dview.cls_shift ? 0 : n_db * BFT_BUCKET_KEYS * 4;
    if (ok) {
        ok = cudaMalloc(&c->d_deep[0], BFT_ROOTDIR_SIZE * sizeof(bft_entry_t)) == cudaSuccess && cudaMalloc(&c->d_deep[1], bk_bytes + 64) == cudaSuccess &&
             (!sc_bytes || cudaMalloc(&c->d_deep[2], sc_bytes + 64) == cudaSuccess);
    }
    if (ok) {
        cudaMemcpyAsync(c->d_deep[0], h_fast, BFT_ROOTDIR_SIZE * sizeof(bft_entry_t), cudaMemcpyHostToDevice, st);
        cudaMemsetAsync(c->d_deep[1], 0xff, bk_bytes + 64, st);
        if (sc_bytes) cudaMemsetAsync(c->d_deep[2], 0xff, sc_bytes + 64, st);
#define BFT_L(W_) k_deep_insert<W_><<<grid_for(c, n_kmers, BFT_TPB), BFT_TPB, 0, st>>>((const bft_entry_t*)c->d_deep[0], d_k, d_cls, n_kmers, \
            (unsigned long long*)c->d_deep[1], (uint32_t*)c->d_deep[2], c->dview.cls_shift, d_failed)
        BFT_BY_W(c->W, BFT_L);
#undef BFT_L
        c->launches++;
        unsigned int h_failed = 0;
        cudaMemcpyAsync(&h_failed, d_failed, sizeof h_failed, cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "collapsed-subtree build failed: %s", cudaGetErrorString(e));
        else if (h_failed) rc = set_err(BFT_B200_ERR_CUDA, "collapsed-subtree build: %u k-mers found no slot", h_failed);
        else {
            c->dview.rootdir_fast = (const bft_entry_t*)c->d_deep[0];
            c->dview.dbuckets = (const uint64_t*)c->d_deep[1];
            c->dview.dslotcls = (const uint32_t*)c->d_deep[2];
            *bytes_out = bk_bytes + sc_bytes + BFT_ROOTDIR_SIZE * sizeof(bft_entry_t);
        }
    }
    if (!ok || rc) {
        (void)cudaGetLastError();
        for (int i = 0; i < 3; i++) { if (c->d_deep[i]) cudaFree(c->d_deep[i]); c->d_deep[i] = NULL; }
    }
    if (d_cnt) cudaFree(d_cnt);
    if (d_failed) cudaFree(d_failed);
    free(h_cnt);
    free(h_fast);
    return rc;
}

static int build_filters(bft_b200_ctx* c, const bft_entry_t* h_rootdir, size_t n_kmers, unsigned long long* d_filter, uint32_t n_blocks, int quirk_safe,
                         size_t* rkf_bytes_out, size_t* deep_bytes_out) {
    *rkf_bytes_out = 0;
    *deep_bytes_out = 0;
    c->dview.kf_quirk_safe = (uint32_t)(quirk_safe != 0);
    /* sectors per prefix of the fused table: enough for BFT_B200_RKF_BITS (default 5.5) filter bits per stored k-mer, at most 4
     * (33.5 MB: it must stay in L2 next to the class rows), at least 3.5 bits per k-mer or not at all. BFT_B200_RKF_SECTORS forces S. */
    uint32_t S = 0;
    {
        const char* es = getenv("BFT_B200_RKF_SECTORS");
        const char* eb = getenv("BFT_B200_RKF_BITS");
        const double want = (eb ? atof(eb) : 5.5) * (double)n_kmers, per_s = 192.0 * BFT_ROOTDIR_SIZE;
        if (es) S = (uint32_t)atoi(es);
        else if (n_kmers) {
            for (S = 1; S < 4 && per_s * S < want; S++) {}
            if (per_s * S < 3.5 * (double)n_kmers) S = 0;
        }
        if (S > 16) S = 16;
    }
    uint64_t* d_k = NULL;
    uint32_t* d_cls = NULL;
    if (cudaMalloc((void**)&d_k, (n_kmers + 1) * (size_t)c->W * 8) != cudaSuccess || cudaMalloc((void**)&d_cls, (n_kmers + 1) * 4) != cudaSuccess) {
        (void)cudaGetLastError();
        if (d_k) cudaFree(d_k);
        return 0; /* no room for the scratch: run without the accelerators */
    }
    c->stats.n_kmers = n_kmers; /* enqueue_extract sizes nothing by it, but keep the context coherent */
    cudaStream_t st = c->streams[0];
    int rc = enqueue_extract(c, d_k, d_cls, NULL);
    if (!rc) rc = build_deep_blocks(c, h_rootdir, d_k, d_cls, n_kmers, deep_bytes_out);
    cudaFree(d_cls);
    if (!rc && n_blocks) {
#define BFT_L(W_) k_kf_insert<W_><<<grid_for(c, n_kmers, BFT_TPB), BFT_TPB, 0, st>>>(d_k, n_kmers, c->k, d_filter, n_blocks)
        BFT_BY_W(c->W, BFT_L);
#undef BFT_L
        c->launches++;
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "stored-k-mer filter build failed: %s", cudaGetErrorString(e));
        if (!rc) {
            c->dview.kfilter = (const uint64_t*)d_filter;
            c->dview.kf_blocks = n_blocks;
        }
    }
    if (!rc && S) {
        /* how many stored k-mers sit under each prefix: expected false-positive rate of a query distributed like the stored
         * k-mers, sum_p c_p (1 - exp(-3 c_p / (192 S)))^3 / n. A trie whose k-mers crowd under a few prefixes (a forced-deep
         * one) would saturate their sectors: it keeps the two-table path, whose filter blocks are shared by all prefixes. */
        uint32_t* d_cnt = NULL;
        uint32_t* h_cnt = (uint32_t*)malloc(BFT_ROOTDIR_SIZE * sizeof(uint32_t));
        bool use = h_cnt && cudaMalloc((void**)&d_cnt, BFT_ROOTDIR_SIZE * sizeof(uint32_t)) == cudaSuccess;
        if (use) {
            cudaMemsetAsync(d_cnt, 0, BFT_ROOTDIR_SIZE * sizeof(uint32_t), st);
#define BFT_L(W_) k_rkf_count<W_><<<grid_for(c, n_kmers, BFT_TPB), BFT_TPB, 0, st>>>(d_k, n_kmers, d_cnt)
            BFT_BY_W(c->W, BFT_L);
#undef BFT_L
            c->launches++;
            use = cudaMemcpyAsync(h_cnt, d_cnt, BFT_ROOTDIR_SIZE * sizeof(uint32_t), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                  cudaStreamSynchronize(st) == cudaSuccess;
        }
        if (use && !getenv("BFT_B200_RKF_SECTORS")) {
            double fp = 0;
            for (size_t p = 0; p < BFT_ROOTDIR_SIZE; p++) {
                const double cp = (double)h_cnt[p], t = 1.0 - exp(-3.0 * cp / (192.0 * S));
                fp += cp * t * t * t;
            }
            use = fp / (double)n_kmers <= 0.15;
        }
        const size_t bytes = (size_t)BFT_ROOTDIR_SIZE * S * 32;
        if (use && cudaMalloc(&c->d_rootkf, bytes + 32) != cudaSuccess) { c->d_rootkf = NULL; use = false; }
        if (use) {
            cudaMemsetAsync(c->d_rootkf, 0, bytes + 32, st);
            k_rkf_fill_entries<<<grid_for(c, (size_t)BFT_ROOTDIR_SIZE * S, BFT_TPB), BFT_TPB, 0, st>>>(
                c->dview.rootdir_fast ? c->dview.rootdir_fast : c->dview.rootdir, (unsigned long long*)c->d_rootkf, S);
#define BFT_L(W_) k_rkf_insert<W_><<<grid_for(c, n_kmers, BFT_TPB), BFT_TPB, 0, st>>>(d_k, n_kmers, (unsigned long long*)c->d_rootkf, S)
            BFT_BY_W(c->W, BFT_L);
#undef BFT_L
            c->launches += 2;
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "fused root directory + filter build failed: %s", cudaGetErrorString(e));
            else {
                c->dview.rootkf = (const uint64_t*)c->d_rootkf;
                c->dview.rkf_sectors = S;
                *rkf_bytes_out = bytes;
            }
        }
        (void)cudaGetLastError();
        if (d_cnt) cudaFree(d_cnt);
        free(h_cnt);
    }
    cudaFree(d_k);
    return rc;
}

extern "C" int bft_b200_open(const char* path, int device, bft_b200_ctx** out) {
    if (!path || !out) return set_err(BFT_B200_ERR_ARG, "bft_b200_open: NULL argument");
    *out = NULL;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(BFT_B200_ERR_CUDA, "bft_b200_open: no CUDA device (%s); this engine has no CPU fallback",
                       e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= ndev) return set_err(BFT_B200_ERR_ARG, "bft_b200_open: device %d out of range (0..%d)", device, ndev - 1);
    CK(cudaSetDevice(device));

    char ferr[256];
    double t0 = now_s();
    bft_arena_t* a = bft_arena_from_file(path, ferr, sizeof ferr);
    if (!a) return set_err(BFT_B200_ERR_FILE, "%s", ferr);
    double t1 = now_s();

    bft_b200_ctx* c = (bft_b200_ctx*)calloc(1, sizeof(bft_b200_ctx));
    if (!c) { bft_arena_free(a); return set_err(BFT_B200_ERR_NOMEM, "out of host memory"); }
    c->device = device;
    c->ref_quirks = 1;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { free(c); bft_arena_free(a); return set_err(BFT_B200_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    c->sm_count = prop.multiProcessorCount;
    c->k = a->k; c->W = a->W; c->G = a->n_genomes; c->rw = (a->n_genomes + 31) / 32;
    if (c->rw < 1) c->rw = 1;
    c->names = a->filenames;
    a->filenames = NULL; /* ownership moved */
    c->n_classes = a->n_classes;

    int rc = 0;
#define UP(i, field, bytes) if (!rc) rc = upload(&c->d_arena[i], a->field, (bytes))
    UP(1, nodes, a->n_nodes * sizeof(bft_node_t));
    UP(2, ccs, a->n_ccs * sizeof(bft_cc_t));
    UP(3, firstcc, a->firstcc_bytes);
    UP(4, csr, a->n_csr * sizeof(uint16_t));
    UP(5, filter3, a->filter3_bytes);
    UP(6, pref, a->n_pref * sizeof(bft_entry_t));
    UP(7, buckets, a->n_buckets * (size_t)(BFT_BUCKET_KEYS * a->W) * sizeof(uint64_t));
    UP(8, slotcls, a->slotcls ? a->n_buckets * BFT_BUCKET_KEYS * sizeof(uint32_t) : 0);
    UP(9, ovf, a->n_ovf * (size_t)a->W * sizeof(uint64_t));
    UP(10, ovfcls, a->ovfcls ? a->n_ovf * sizeof(uint32_t) : 0);
    UP(11, uckeys, a->n_uc_lines * (size_t)a->W * sizeof(uint64_t));
    UP(12, uccls, a->n_uc_lines * sizeof(uint32_t));
    UP(13, pref_low18, a->n_pref * sizeof(uint32_t));
    UP(14, pref_node, a->n_pref * sizeof(uint32_t));
    UP(15, node_path, a->n_nodes * sizeof(bft_path_t));
    UP(16, pref_out, (a->n_pref + 1) * sizeof(uint64_t));
    UP(17, uc_rank, a->n_uc_lines);
    UP(18, hpos, a->n_hpos * sizeof(uint32_t));
    c->n_pref = a->n_pref;
    c->n_nodes = a->n_nodes;
#undef UP
    void* d_cls_off = NULL; void* d_cls_bytes = NULL;
    if (!rc) rc = upload(&d_cls_off, a->cls_off, (a->n_classes + 1) * sizeof(uint32_t));
    if (!rc) rc = upload(&d_cls_bytes, a->cls_bytes, a->cls_bytes_len);
    if (!rc) rc = upload(&c->d_pool[0], a->pool_last_index, (size_t)a->n_pools * sizeof(int64_t));
    if (!rc) rc = upload(&c->d_pool[1], a->pool_size_annot, (size_t)a->n_pools * sizeof(int32_t));
    if (!rc) rc = upload(&c->d_pool[2], a->pool_off, (size_t)a->n_pools * sizeof(uint64_t));
    if (!rc) rc = upload(&c->d_pool[3], a->pool_bytes, a->pool_bytes_len);
    double t2 = now_s();
    if (!rc) {
        c->dview.nodes = (const bft_node_t*)c->d_arena[1];
        c->dview.ccs = (const bft_cc_t*)c->d_arena[2];
        c->dview.firstcc = (const uint8_t*)c->d_arena[3];
        c->dview.csr = (const uint16_t*)c->d_arena[4];
        c->dview.filter3 = (const uint8_t*)c->d_arena[5];
        c->dview.pref = (const bft_entry_t*)c->d_arena[6];
        c->dview.buckets = (const uint64_t*)c->d_arena[7];
        c->dview.slotcls = (const uint32_t*)c->d_arena[8];
        c->dview.ovf = (const uint64_t*)c->d_arena[9];
        c->dview.ovfcls = (const uint32_t*)c->d_arena[10];
        c->dview.uckeys = (const uint64_t*)c->d_arena[11];
        c->dview.uccls = (const uint32_t*)c->d_arena[12];
        c->dview.pref_low18 = (const uint32_t*)c->d_arena[13];
        c->dview.pref_node = (const uint32_t*)c->d_arena[14];
        c->dview.node_path = (const bft_path_t*)c->d_arena[15];
        c->dview.pref_out = (const uint64_t*)c->d_arena[16];
        c->dview.uc_rank = (const uint8_t*)c->d_arena[17];
        c->dview.hpos = (const uint32_t*)c->d_arena[18];
        c->dview.cls_shift = a->cls_shift;
        c->dview.cls_mask = a->cls_mask;
        c->dview.k = a->k;
        c->dview.W = a->W;
        /* storage locations are 32-bit in the kernels; counted here in 64 bits so a BFT with 2^32 or more of them is
         * caught (bft_b200_graph_prepare refuses it; queries do not use locations) instead of wrapping silently */
        c->n_loc = (uint64_t)a->n_buckets * BFT_BUCKET_KEYS + a->n_ovf + a->n_uc_lines + a->n_pref;
        c->dview.loc_ovf = (uint32_t)(a->n_buckets * BFT_BUCKET_KEYS);
        c->dview.loc_uc = c->dview.loc_ovf + (uint32_t)a->n_ovf;
        c->dview.loc_leaf = c->dview.loc_uc + (uint32_t)a->n_uc_lines;
        c->dpools.n_pools = a->n_pools;
        c->dpools.last_index = (const int64_t*)c->d_pool[0];
        c->dpools.size_annot = (const int32_t*)c->d_pool[1];
        c->dpools.off = (const uint64_t*)c->d_pool[2];
        c->dpools.bytes = (const uint8_t*)c->d_pool[3];
    }
    /* decode every distinct annotation on the device */
    int* d_bad = NULL;
    int h_bad = 0;
    const size_t row_bytes = (a->n_classes + 1) * (size_t)c->rw * sizeof(uint32_t);
    /* hot block: rootdir, the class rows, the stored-k-mer filter */
    const size_t rootdir_bytes = BFT_ROOTDIR_SIZE * sizeof(bft_entry_t);
    const size_t rows_padded = (row_bytes + 31) & ~(size_t)31;
    /* filter size: BFT_B200_KF_BITS bits per stored k-mer (default 6: 5.6 % false positives; 0 = no filter), shrunk down
     * to 3 bits per k-mer (29 %) to stay within BFT_B200_KF_MAX_MB (default 36 MB). Measured on C3 (33 M k-mers, B200):
     * 25 MB -> 45.2 G k-mers/s, 33 MB -> 44.4, 45 MB -> 41.4, 58 MB -> 33.2 (no better than without a filter): the filter
     * shares the 126 MB L2 with the class rows, the root directory and the streamed batch. A BFT too large for that
     * gets no filter — out of L2 it would cost a second HBM access per present k-mer instead of saving one per absent one */
    size_t kf_blocks = 0;
    {
        const char* eb = getenv("BFT_B200_KF_BITS");
        const char* em = getenv("BFT_B200_KF_MAX_MB");
        double bits = eb ? atof(eb) : 6.0;
        const double max_bytes = (em ? atof(em) : 36.0) * 1048576.0;
        if (bits > 0 && a->n_kmers > 0) {
            if (bits * (double)a->n_kmers / 8.0 > max_bytes) bits = max_bytes * 8.0 / (double)a->n_kmers;
            if (bits >= 3.0) {
                kf_blocks = (size_t)(bits * (double)a->n_kmers / 256.0) + 1;
                if (kf_blocks > 0xffffffffu) kf_blocks = 0;
            }
        }
    }
    const size_t kf_bytes = kf_blocks * 32;
    c->hot_bytes = rootdir_bytes + rows_padded + kf_bytes;
    if (!rc && cudaMalloc(&c->d_hot, c->hot_bytes + 32) != cudaSuccess) rc = set_err(BFT_B200_ERR_NOMEM, "cudaMalloc(hot tables, %zu) failed", c->hot_bytes);
    if (!rc && cudaMemcpy(c->d_hot, a->rootdir, rootdir_bytes, cudaMemcpyHostToDevice) != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "rootdir upload failed");
    if (!rc && kf_bytes && cudaMemset((char*)c->d_hot + rootdir_bytes + rows_padded, 0, kf_bytes) != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "filter clear failed");
    if (!rc) {
        c->d_class_rows = (uint32_t*)((char*)c->d_hot + rootdir_bytes);
        c->dview.rootdir = (const bft_entry_t*)c->d_hot;
        c->dview.rows_keep = row_bytes <= ((size_t)48 << 20); /* evict-last priority only for a table that can stay in L2 */
    }
    if (!rc && cudaMalloc((void**)&c->d_class_counts, (a->n_classes + 1) * sizeof(uint32_t)) != cudaSuccess) rc = set_err(BFT_B200_ERR_NOMEM, "cudaMalloc(class counts) failed");
    if (!rc && cudaMalloc((void**)&d_bad, sizeof(int)) != cudaSuccess) rc = set_err(BFT_B200_ERR_NOMEM, "cudaMalloc failed");
    if (!rc && cudaMalloc((void**)&c->d_counter, sizeof(unsigned long long)) != cudaSuccess) rc = set_err(BFT_B200_ERR_NOMEM, "cudaMalloc failed");
    for (int s = 0; s < BFT_N_SLOTS && !rc; s++)
        if (cudaStreamCreateWithFlags(&c->streams[s], cudaStreamNonBlocking) != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "cudaStreamCreate failed");
    if (!rc && getenv("BFT_B200_L2_FETCH")) { /* experiment knob (device-wide hint): DRAM -> L2 fetch granularity in bytes (32, 64 or 128) */
        const int gran = atoi(getenv("BFT_B200_L2_FETCH"));
        if (gran == 32 || gran == 64 || gran == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
        cudaGetLastError();
    }
    if (!rc && getenv("BFT_B200_L2_PERSIST") && getenv("BFT_B200_L2_PERSIST")[0] == '1') {
        /* Opt-in experiment knob, OFF by default: a persisting access-policy window over the hot block. Measured on B200 it buys
         * nothing (C3: 46.8 vs 46.7 G k-mers/s — the evict_last / evict_first hints on the loads already keep the hot tables in
         * L2) and it costs: the set-aside is a DEVICE-wide limit that outlives the context, so every later context in the
         * process (and the host application) runs with that much less L2 — the 1000-colour config dropped from 28.0 to 18.9 G
         * k-mers/s, the 4-genome one from 74.6 to 59.2, when opened after the 100-genome BFT (profiles/r02_l2_persist_ab.md). */
        size_t win = c->hot_bytes;
        if (win > (size_t)prop.accessPolicyMaxWindowSize) win = (size_t)prop.accessPolicyMaxWindowSize;
        size_t persist = win; /* measured on B200: a set-aside larger than the window buys nothing, 64 MB costs 5 % */
        if (persist > (size_t)prop.persistingL2CacheMaxSize) persist = (size_t)prop.persistingL2CacheMaxSize;
        size_t have = 0; /* the limit is device-wide: raise it if ours is larger, never shrink what another context set */
        if (cudaDeviceGetLimit(&have, cudaLimitPersistingL2CacheSize) != cudaSuccess) have = 0;
        if (win && persist && (have >= persist || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess)) {
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof attr);
            attr.accessPolicyWindow.base_ptr = c->d_hot;
            attr.accessPolicyWindow.num_bytes = win;
            attr.accessPolicyWindow.hitRatio = persist >= win ? 1.0f : (float)((double)persist / (double)win);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            for (int s2 = 0; s2 < BFT_N_SLOTS; s2++) cudaStreamSetAttribute(c->streams[s2], cudaStreamAttributeAccessPolicyWindow, &attr);
        }
        cudaGetLastError(); /* a refusal only costs performance */
    }
    if (!rc) {
        cudaMemsetAsync(d_bad, 0, sizeof(int), c->streams[0]);
        k_decode_classes<<<grid_for(c, a->n_classes, BFT_TPB), BFT_TPB, 0, c->streams[0]>>>(
            (const uint32_t*)d_cls_off, (const uint8_t*)d_cls_bytes, a->n_classes, c->dpools, c->d_class_rows, c->d_class_counts, c->rw, d_bad);
        c->launches++;
        cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c->streams[0]);
        e = cudaStreamSynchronize(c->streams[0]);
        if (e != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "k_decode_classes failed: %s", cudaGetErrorString(e));
        else if (h_bad) rc = set_err(BFT_B200_ERR_FILE, "%d colour annotations are malformed (mode 3 pointing outside the colour pools)", h_bad);
    }
    size_t rkf_bytes = 0, deep_bytes = 0;
    if (!rc && a->n_kmers) rc = build_filters(c, a->rootdir, a->n_kmers, (unsigned long long*)((char*)c->d_hot + rootdir_bytes + rows_padded), (uint32_t)kf_blocks,
                                              a->max_depth < a->k / BFT_NB_CHAR_SUF_PREF || a->k == BFT_NB_CHAR_SUF_PREF, &rkf_bytes, &deep_bytes);
    double t3 = now_s();
    if (d_bad) cudaFree(d_bad);
    if (d_cls_off) cudaFree(d_cls_off);
    if (d_cls_bytes) cudaFree(d_cls_bytes);

    c->stats.n_kmers = a->n_kmers; c->stats.n_nodes = a->n_nodes; c->stats.n_ccs = a->n_ccs; c->stats.n_lines = a->n_lines + a->n_uc_lines;
    c->stats.n_prefixes = a->n_pref; c->stats.n_classes = a->n_classes; c->stats.arena_bytes = bft_arena_bytes(a);
    c->stats.class_row_bytes = row_bytes; c->stats.max_cc_per_node = a->max_cc_per_node; c->stats.max_depth = a->max_depth;
    c->stats.n_pools = a->n_pools;
    c->stats.filter_bytes = rc ? 0 : kf_bytes;
    c->stats.rootkf_bytes = rc ? 0 : rkf_bytes;
    c->stats.deep_bytes = rc ? 0 : deep_bytes;
    c->stats.flatten_seconds = t1 - t0; c->stats.upload_seconds = t2 - t1; c->stats.decode_seconds = t3 - t2;
    bft_arena_free(a);

    /* shared memory for the sequence kernel */
    c->seq_smem = bft_seq_smem_per_warp(c->G) * BFT_SEQ_WARPS;
    if (!rc && c->seq_smem > 48 * 1024) {
        if (c->seq_smem > (size_t)prop.sharedMemPerBlockOptin)
            c->seq_smem = 0; /* too many genomes for the shared-memory counters: sequence queries will refuse */
        else {
            cudaFuncSetAttribute(k_query_sequences<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->seq_smem);
            cudaFuncSetAttribute(k_query_sequences<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->seq_smem);
            cudaFuncSetAttribute(k_query_sequences<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->seq_smem);
        }
    }
    if (rc) { bft_b200_close(c); return rc; }
    *out = c;
    return BFT_B200_OK;
}

extern "C" int bft_b200_k(const bft_b200_ctx* c) { return c ? c->k : 0; }
extern "C" int bft_b200_n_genomes(const bft_b200_ctx* c) { return c ? c->G : 0; }
extern "C" const char* bft_b200_genome_name(const bft_b200_ctx* c, int i) { return (c && i >= 0 && i < c->G) ? c->names[i] : NULL; }
extern "C" int bft_b200_kmer_words(const bft_b200_ctx* c) { return c ? c->W : 0; }
extern "C" int bft_b200_row_words(const bft_b200_ctx* c) { return c ? c->rw : 0; }
extern "C" int bft_b200_device(const bft_b200_ctx* c) { return c ? c->device : -1; }
extern "C" void* bft_b200_stream(const bft_b200_ctx* c) { return c ? (void*)c->streams[0] : NULL; }
extern "C" uint64_t bft_b200_launch_count(const bft_b200_ctx* c) { return c ? c->launches : 0; }
extern "C" int bft_b200_get_stats(const bft_b200_ctx* c, bft_b200_stats* out) {
    if (!c || !out) return set_err(BFT_B200_ERR_ARG, "bft_b200_get_stats: NULL argument");
    *out = c->stats;
    return 0;
}

extern "C" void* bft_b200_host_alloc(size_t bytes) {
    void* p = NULL;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { set_err(BFT_B200_ERR_NOMEM, "cudaMallocHost(%zu) failed", bytes); return NULL; }
    return p;
}
extern "C" void bft_b200_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int bft_b200_sync(bft_b200_ctx* c) {
    if (!c) return set_err(BFT_B200_ERR_ARG, "bft_b200_sync: NULL context");
    CK(cudaSetDevice(c->device));
    for (int s = 0; s < BFT_N_SLOTS; s++) CK(cudaStreamSynchronize(c->streams[s]));
    return 0;
}

/* ---- enqueue helpers (device pointers, one stream) ---------------------------------------------------------- */
static int enqueue_kmers(bft_b200_ctx* c, cudaStream_t st, const uint64_t* d_kmers, size_t n, uint8_t* d_present, uint32_t* d_rows,
                         uint32_t* d_cls, unsigned long long* d_n_present = NULL) {
    if (n == 0) return 0;
    const int grid = grid_for(c, n, BFT_TPB);
    if (d_rows && (c->rw == 1 || c->rw == 2 || c->rw == 4) && ((uintptr_t)d_rows & 15) == 0) { /* narrow rows: one fused kernel */
#define BFT_FUSED(W_, RW_) k_query_kmers_rows<W_, RW_><<<grid, BFT_TPB, 0, st>>>(c->dview, d_kmers, n, d_present, d_cls, c->d_class_rows, d_rows, d_n_present)
#define BFT_FUSED_W(W_) do { if (c->rw == 4) BFT_FUSED(W_, 4); else if (c->rw == 2) BFT_FUSED(W_, 2); else BFT_FUSED(W_, 1); } while (0)
        BFT_BY_W(c->W, BFT_FUSED_W);
#undef BFT_FUSED_W
#undef BFT_FUSED
        c->launches++;
        CK(cudaGetLastError());
        return 0;
    }
    if (d_rows) { /* wide rows: walk and row expansion in one kernel, 32 rows per warp written together */
        if (c->rw % 4 == 0 && ((uintptr_t)d_rows & 15) == 0 && ((uintptr_t)c->d_class_rows & 15) == 0) {
#define BFT_L(W_) k_query_kmers_wide<W_, uint4><<<grid, BFT_TPB, 0, st>>>(c->dview, d_kmers, n, d_present, d_cls, (const uint4*)c->d_class_rows, c->rw / 4, (uint4*)d_rows, d_n_present)
            BFT_BY_W(c->W, BFT_L);
#undef BFT_L
        } else {
#define BFT_L(W_) k_query_kmers_wide<W_, uint32_t><<<grid, BFT_TPB, 0, st>>>(c->dview, d_kmers, n, d_present, d_cls, c->d_class_rows, c->rw, d_rows, d_n_present)
            BFT_BY_W(c->W, BFT_L);
#undef BFT_L
        }
        c->launches++;
        CK(cudaGetLastError());
        return 0;
    }
#define BFT_L(W_) k_query_kmers<W_><<<grid, BFT_TPB, 0, st>>>(c->dview, d_kmers, n, d_present, d_cls)
    BFT_BY_W(c->W, BFT_L);
#undef BFT_L
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int bft_b200_query_kmers_device(bft_b200_ctx* c, const uint64_t* d_kmers, size_t n, uint8_t* d_present, uint32_t* d_rows,
                                           uint32_t* d_cls) {
    if (!c || (!d_kmers && n)) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_kmers_device: NULL argument");
    CK(cudaSetDevice(c->device));
    return enqueue_kmers(c, c->streams[0], d_kmers, n, d_present, d_rows, d_cls);
}

extern "C" int bft_b200_query_kmers_device_counted(bft_b200_ctx* c, const uint64_t* d_kmers, size_t n, uint8_t* d_present, uint32_t* d_rows,
                                                   uint64_t* d_n_present) {
    if (!c || (!d_kmers && n) || !d_rows || !d_n_present) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_kmers_device_counted: NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(d_n_present, 0, sizeof(uint64_t), c->streams[0]));
    return enqueue_kmers(c, c->streams[0], d_kmers, n, d_present, d_rows, NULL, (unsigned long long*)d_n_present);
}

extern "C" int bft_b200_query_kmers_device_accumulate(bft_b200_ctx* c, const uint64_t* d_kmers, size_t n, uint8_t* d_present, uint32_t* d_rows,
                                                      uint64_t* d_counter) {
    if (!c || (!d_kmers && n) || !d_rows || !d_counter) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_kmers_device_accumulate: NULL argument");
    CK(cudaSetDevice(c->device));
    return enqueue_kmers(c, c->streams[0], d_kmers, n, d_present, d_rows, NULL, (unsigned long long*)d_counter);
}

static int query_kmers_host(bft_b200_ctx* c, const uint64_t* kmers, const char* ascii, size_t n, uint8_t* valid, uint8_t* present,
                            uint32_t* rows, uint32_t* class_ids) {
    CK(cudaSetDevice(c->device));
    const size_t rw = (size_t)c->rw, W = (size_t)c->W, k = (size_t)c->k;
    size_t done = 0;
    int it = 0;
    while (done < n) {
        const size_t m = n - done < BFT_CHUNK_KMERS ? n - done : BFT_CHUNK_KMERS;
        const int s = it & 1;
        slot_t* sl = &c->slot[s];
        cudaStream_t st = c->streams[s];
        CKP(cudaStreamSynchronize(st)); /* slot buffers free again */
        const uint64_t* d_k;
        if (ascii) {
            ENSUREP(sl->d_in, sl->cap_in, m * k);
            ENSUREP(sl->d_kmers, sl->cap_kmers, m * W * 8);
            ENSUREP(sl->d_u8b, sl->cap_u8b, m);
            CKP(cudaMemcpyAsync(sl->d_in, ascii + done * k, m * k, cudaMemcpyHostToDevice, st));
#define BFT_L(W_) k_encode_ascii<W_><<<grid_for(c, m, BFT_TPB), BFT_TPB, 0, st>>>((const char*)sl->d_in, m, c->k, sl->d_kmers, sl->d_u8b)
            BFT_BY_W(c->W, BFT_L);
#undef BFT_L
            c->launches++;
            d_k = sl->d_kmers;
        } else {
            ENSUREP(sl->d_in, sl->cap_in, m * W * 8);
            CKP(cudaMemcpyAsync(sl->d_in, kmers + done * W, m * W * 8, cudaMemcpyHostToDevice, st));
            d_k = (const uint64_t*)sl->d_in;
        }
        ENSUREP(sl->d_u8a, sl->cap_u8a, m);
        ENSUREP(sl->d_cls, sl->cap_cls, m * sizeof(uint32_t));
        if (rows) ENSUREP(sl->d_rows, sl->cap_rows, m * rw * sizeof(uint32_t));
        /* class ids are materialised only when asked for, or as the intermediate of the wide-row path */
        const int fused = rows && (c->rw == 1 || c->rw == 2 || c->rw == 4);
        uint32_t* d_cls = (class_ids || (rows && !fused) || !rows) ? sl->d_cls : NULL;
        int rc = enqueue_kmers(c, st, d_k, m, sl->d_u8a, rows ? sl->d_rows : NULL, d_cls);
        if (rc) return drain_ret(c, rc);
        if (ascii) {
            k_blank_invalid<<<grid_for(c, m, BFT_TPB), BFT_TPB, 0, st>>>(sl->d_u8b, m, c->rw, sl->d_u8a, d_cls, rows ? sl->d_rows : NULL);
            c->launches++;
        }
        if (present) CKP(cudaMemcpyAsync(present + done, sl->d_u8a, m, cudaMemcpyDeviceToHost, st));
        if (valid) CKP(cudaMemcpyAsync(valid + done, sl->d_u8b, m, cudaMemcpyDeviceToHost, st));
        if (class_ids) CKP(cudaMemcpyAsync(class_ids + done, sl->d_cls, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (rows) CKP(cudaMemcpyAsync(rows + done * rw, sl->d_rows, m * rw * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        done += m;
        it++;
    }
    CKP(cudaStreamSynchronize(c->streams[0]));
    CKP(cudaStreamSynchronize(c->streams[1]));
    return 0;
}

extern "C" int bft_b200_query_kmers(bft_b200_ctx* c, const uint64_t* kmers, size_t n, uint8_t* present, uint32_t* rows, uint32_t* class_ids) {
    if (!c || (!kmers && n)) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_kmers: NULL argument");
    return query_kmers_host(c, kmers, NULL, n, NULL, present, rows, class_ids);
}

/* ---- the reference's record format in, byte rows out ------------------------------------------------------------ */
static int enqueue_records(bft_b200_ctx* c, cudaStream_t st, const uint8_t* d_records, size_t n, uint8_t* d_present, uint8_t* d_rows,
                           unsigned long long* d_n_present) {
    if (n == 0) return 0;
    const int nb = (2 * c->k + 7) / 8, rb = (c->G + 7) / 8 > 0 ? (c->G + 7) / 8 : 1;
    const size_t smem = bft_records_smem(nb, rb);
    if (smem > 227 * 1024) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_records: %d genomes exceed the shared-memory tile; use bft_b200_query_kmers", c->G);
    if (smem > 48 * 1024) { /* opt in to large dynamic shared memory (wide colour rows) */
#define BFT_L(W_) CK(cudaFuncSetAttribute(k_query_records<W_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem))
        BFT_BY_W(c->W, BFT_L);
#undef BFT_L
    }
    const size_t n_tiles = (n + BFT_TPB - 1) / BFT_TPB;
    const size_t want = (size_t)c->sm_count * 8;
    const int grid = (int)(n_tiles < want ? n_tiles : want);
#define BFT_L(W_) k_query_records<W_, false><<<grid, BFT_TPB, smem, st>>>(c->dview, d_records, n, nb, rb, c->rw, c->d_class_rows, d_present, d_rows, d_n_present, NULL, NULL, NULL)
    BFT_BY_W(c->W, BFT_L);
#undef BFT_L
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

/* compact form, device side of one chunk: class ids + presence bits + per-tile counts, prefix sums, rows of the present
 * k-mers back to back. d_tile holds [counts | offsets | total]. */
static int enqueue_records_compact(bft_b200_ctx* c, cudaStream_t st, slot_t* sl, size_t n) {
    const int nb = (2 * c->k + 7) / 8, rb = bft_b200_row_bytes(c);
    const size_t n_tiles = (n + BFT_TPB - 1) / BFT_TPB;
    const size_t smem_in = 2 * (size_t)BFT_TPB * (size_t)nb + 16, smem_rows = (size_t)BFT_TPB * (size_t)rb + 16;
    if (smem_rows > 227 * 1024) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_records_compact: %d genomes exceed the shared-memory tile", c->G);
    if (smem_rows > 48 * 1024) CK(cudaFuncSetAttribute(k_compact_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    const size_t want = (size_t)c->sm_count * 8;
    const int grid = (int)(n_tiles < want ? n_tiles : want);
    uint32_t* d_cnt = sl->d_tile;
    uint32_t* d_off = sl->d_tile + n_tiles;
    uint32_t* d_total = sl->d_tile + 2 * n_tiles;
    /* rb = 0 in the first pass: no row tiles in shared memory */
#define BFT_L(W_) k_query_records<W_, true><<<grid, BFT_TPB, smem_in, st>>>(c->dview, (const uint8_t*)sl->d_in, n, nb, 0, c->rw, c->d_class_rows, NULL, NULL, NULL, sl->d_cls, (uint32_t*)sl->d_u8a, d_cnt)
    BFT_BY_W(c->W, BFT_L);
#undef BFT_L
    k_scan_tile_counts<<<1, 1024, 0, st>>>(d_cnt, (uint32_t)n_tiles, d_off, d_total);
    k_compact_rows<<<grid, BFT_TPB, smem_rows, st>>>(sl->d_cls, n, d_off, c->d_class_rows, c->rw, rb, (uint8_t*)sl->d_rows);
    c->launches += 3;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int bft_b200_row_bytes(const bft_b200_ctx* c) { return c ? ((c->G + 7) / 8 > 0 ? (c->G + 7) / 8 : 1) : 0; }
extern "C" int bft_b200_record_bytes(const bft_b200_ctx* c) { return c ? (2 * c->k + 7) / 8 : 0; }

extern "C" int bft_b200_query_records_device(bft_b200_ctx* c, const uint8_t* d_records, size_t n, uint8_t* d_present, uint8_t* d_rows,
                                             uint64_t* d_n_present) {
    if (!c || (!d_records && n) || !d_rows) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_records_device: NULL argument");
    if (((uintptr_t)d_records & 15) || ((uintptr_t)d_rows & 15))
        return set_err(BFT_B200_ERR_ARG, "bft_b200_query_records_device: record and row buffers must be 16-byte aligned");
    CK(cudaSetDevice(c->device));
    if (d_n_present) CK(cudaMemsetAsync(d_n_present, 0, sizeof(uint64_t), c->streams[0]));
    return enqueue_records(c, c->streams[0], d_records, n, d_present, d_rows, (unsigned long long*)d_n_present);
}

extern "C" int bft_b200_query_records(bft_b200_ctx* c, const uint8_t* records, size_t n, uint8_t* present, uint8_t* rows, uint64_t* n_present) {
    if (!c || (!records && n) || !rows) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_records: NULL argument");
    CK(cudaSetDevice(c->device));
    const size_t nb = (size_t)bft_b200_record_bytes(c), rb = (size_t)bft_b200_row_bytes(c);
    CKP(cudaStreamSynchronize(c->streams[0]));
    CKP(cudaStreamSynchronize(c->streams[1]));
    CKP(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), c->streams[0]));
    CKP(cudaStreamSynchronize(c->streams[0]));
    size_t done = 0;
    int it = 0;
    while (done < n) {
        const size_t m = n - done < BFT_CHUNK_KMERS ? n - done : BFT_CHUNK_KMERS;
        const int s = it & 1;
        slot_t* sl = &c->slot[s];
        cudaStream_t st = c->streams[s];
        CKP(cudaStreamSynchronize(st)); /* slot buffers free again */
        ENSUREP(sl->d_in, sl->cap_in, m * nb);
        ENSUREP(sl->d_u8a, sl->cap_u8a, m);
        ENSUREP(sl->d_rows, sl->cap_rows, m * rb);
        CKP(cudaMemcpyAsync(sl->d_in, records + done * nb, m * nb, cudaMemcpyHostToDevice, st));
        int rc = enqueue_records(c, st, (const uint8_t*)sl->d_in, m, present ? sl->d_u8a : NULL, (uint8_t*)sl->d_rows, n_present ? c->d_counter : NULL);
        if (rc) return drain_ret(c, rc);
        if (present) CKP(cudaMemcpyAsync(present + done, sl->d_u8a, m, cudaMemcpyDeviceToHost, st));
        CKP(cudaMemcpyAsync(rows + done * rb, sl->d_rows, m * rb, cudaMemcpyDeviceToHost, st));
        done += m;
        it++;
    }
    CKP(cudaStreamSynchronize(c->streams[0]));
    CKP(cudaStreamSynchronize(c->streams[1]));
    if (n_present) {
        unsigned long long h = 0;
        CKP(cudaMemcpy(&h, c->d_counter, sizeof h, cudaMemcpyDeviceToHost));
        *n_present = h;
    }
    return 0;
}

/* Rows of a chunk can only be copied once its hit count is known on the host. Four slots on four streams, and the
 * copy of chunk i - 2 is issued after the front half of chunk i has been enqueued: by then the count is there, so the
 * host does not wait and neither copy engine idles. */
extern "C" int bft_b200_query_records_compact(bft_b200_ctx* c, const uint8_t* records, size_t n, uint8_t* present_bits, uint8_t* rows,
                                              uint64_t* n_present) {
    if (!c || (!records && n) || !present_bits || !rows || !n_present) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_records_compact: NULL argument");
    CK(cudaSetDevice(c->device));
    const size_t nb = (size_t)bft_b200_record_bytes(c), rb = (size_t)bft_b200_row_bytes(c);
    for (int s = 0; s < BFT_N_SLOTS; s++) CK(cudaStreamSynchronize(c->streams[s]));
    const size_t n_chunks = (n + BFT_CHUNK_KMERS - 1) / BFT_CHUNK_KMERS;
    const size_t lag = 2;
    size_t rows_done = 0;
    *n_present = 0;
    for (size_t i = 0; i < n_chunks + lag; i++) {
        if (i < n_chunks) { /* front half of chunk i: records in, look-ups, compaction, presence bits and hit count out */
            slot_t* sl = &c->slot[i % BFT_N_SLOTS];
            cudaStream_t st = c->streams[i % BFT_N_SLOTS];
            const size_t first = i * BFT_CHUNK_KMERS;
            const size_t m = n - first < BFT_CHUNK_KMERS ? n - first : BFT_CHUNK_KMERS;
            const size_t n_tiles = (m + BFT_TPB - 1) / BFT_TPB;
            CKP(cudaStreamSynchronize(st)); /* the slot's rows of four chunks ago reached the host long ago */
            ENSUREP(sl->d_in, sl->cap_in, m * nb);
            ENSUREP(sl->d_cls, sl->cap_cls, m * sizeof(uint32_t));
            ENSUREP(sl->d_u8a, sl->cap_u8a, (m + 31) / 32 * 4);
            ENSUREP(sl->d_rows, sl->cap_rows, m * rb);
            ENSUREP(sl->d_tile, sl->cap_tile, (2 * n_tiles + 1) * sizeof(uint32_t));
            if (!sl->h_total) CKP(cudaMallocHost((void**)&sl->h_total, 16));
            CKP(cudaMemcpyAsync(sl->d_in, records + first * nb, m * nb, cudaMemcpyHostToDevice, st));
            int rc = enqueue_records_compact(c, st, sl, m);
            if (rc) return drain_ret(c, rc);
            CKP(cudaMemcpyAsync(present_bits + first / 8, sl->d_u8a, (m + 7) / 8, cudaMemcpyDeviceToHost, st)); /* chunks are multiples of 8 */
            CKP(cudaMemcpyAsync(sl->h_total, sl->d_tile + 2 * n_tiles, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        }
        if (i >= lag) { /* back half of chunk i - lag: its rows, now that their number is known */
            const size_t j = i - lag;
            slot_t* pl = &c->slot[j % BFT_N_SLOTS];
            cudaStream_t pst = c->streams[j % BFT_N_SLOTS];
            CKP(cudaStreamSynchronize(pst));
            const size_t hits = pl->h_total[0];
            if (hits) CKP(cudaMemcpyAsync(rows + rows_done * rb, pl->d_rows, hits * rb, cudaMemcpyDeviceToHost, pst));
            rows_done += hits;
        }
    }
    for (int s = 0; s < BFT_N_SLOTS; s++) CK(cudaStreamSynchronize(c->streams[s]));
    *n_present = rows_done;
    return 0;
}

extern "C" int bft_b200_query_kmers_ascii(bft_b200_ctx* c, const char* ascii, size_t n, uint8_t* valid, uint8_t* present, uint32_t* rows,
                                          uint32_t* class_ids) {
    if (!c || (!ascii && n)) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_kmers_ascii: NULL argument");
    return query_kmers_host(c, NULL, ascii, n, valid, present, rows, class_ids);
}

extern "C" int bft_b200_class_rows(bft_b200_ctx* c, const uint32_t** rows, uint64_t* n_classes) {
    if (!c || !rows) return set_err(BFT_B200_ERR_ARG, "bft_b200_class_rows: NULL argument");
    CK(cudaSetDevice(c->device));
    if (!c->h_class_rows) {
        const size_t bytes = (c->n_classes + 1) * (size_t)c->rw * sizeof(uint32_t);
        c->h_class_rows = (uint32_t*)malloc(bytes);
        if (!c->h_class_rows) return set_err(BFT_B200_ERR_NOMEM, "out of host memory");
        CK(cudaMemcpy(c->h_class_rows, c->d_class_rows, bytes, cudaMemcpyDeviceToHost));
    }
    *rows = c->h_class_rows;
    if (n_classes) *n_classes = c->n_classes;
    return 0;
}

extern "C" int bft_b200_class_counts(bft_b200_ctx* c, const uint32_t** counts, uint64_t* n_classes) {
    if (!c || !counts) return set_err(BFT_B200_ERR_ARG, "bft_b200_class_counts: NULL argument");
    CK(cudaSetDevice(c->device));
    if (!c->h_class_counts) {
        const size_t bytes = (c->n_classes + 1) * sizeof(uint32_t);
        c->h_class_counts = (uint32_t*)malloc(bytes);
        if (!c->h_class_counts) return set_err(BFT_B200_ERR_NOMEM, "out of host memory");
        CK(cudaMemcpy(c->h_class_counts, c->d_class_counts, bytes, cudaMemcpyDeviceToHost));
    }
    *counts = c->h_class_counts;
    if (n_classes) *n_classes = c->n_classes;
    return 0;
}

/* ---- annotation set algebra ----------------------------------------------------------------------------------------- */
extern "C" int bft_b200_annotation_setop_device(bft_b200_ctx* c, int op, const uint32_t* d_class_ids, const uint64_t* d_group_offs, size_t n_groups,
                                                uint32_t* d_rows, uint32_t* d_counts) {
    if (!c || (n_groups && (!d_class_ids || !d_group_offs)) || (!d_rows && !d_counts)) return set_err(BFT_B200_ERR_ARG, "bft_b200_annotation_setop_device: NULL argument");
    if (op < 0 || op > 2) return set_err(BFT_B200_ERR_ARG, "bft_b200_annotation_setop: op %d is not BFT_B200_SET_INTERSECTION / _UNION / _SYM_DIFFERENCE", op);
    CK(cudaSetDevice(c->device));
    if (!n_groups) return 0;
    cudaStream_t st = c->streams[0];
    if (d_counts && c->rw > 1) CK(cudaMemsetAsync(d_counts, 0, n_groups * sizeof(uint32_t), st));
    k_annotation_setop<<<grid_for(c, n_groups * (size_t)c->rw, BFT_TPB), BFT_TPB, 0, st>>>(c->d_class_rows, c->rw, d_class_ids, d_group_offs, n_groups, op, d_rows, d_counts);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int bft_b200_annotation_setop(bft_b200_ctx* c, int op, const uint32_t* class_ids, const uint64_t* group_offs, size_t n_groups, uint32_t* rows,
                                         uint32_t* counts) {
    if (!c || (n_groups && (!class_ids || !group_offs)) || (!rows && !counts)) return set_err(BFT_B200_ERR_ARG, "bft_b200_annotation_setop: NULL argument");
    if (op < 0 || op > 2) return set_err(BFT_B200_ERR_ARG, "bft_b200_annotation_setop: op %d is not BFT_B200_SET_INTERSECTION / _UNION / _SYM_DIFFERENCE", op);
    if (!n_groups) return 0;
    for (size_t g = 0; g < n_groups; g++)
        if (group_offs[g + 1] < group_offs[g]) return set_err(BFT_B200_ERR_ARG, "bft_b200_annotation_setop: group offsets must be non-decreasing");
    const size_t n_ids = (size_t)(group_offs[n_groups] - group_offs[0]);
    for (size_t i = 0; i < n_ids; i++) {
        const uint32_t id = class_ids[group_offs[0] + i];
        if (id != BFT_CLS_NONE && id >= c->n_classes) return set_err(BFT_B200_ERR_ARG, "bft_b200_annotation_setop: class id %u out of range (%zu classes)", id, c->n_classes);
    }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->streams[0];
    CK(cudaStreamSynchronize(st));
    slot_t* sl = &c->slot[0];
    const size_t rw = (size_t)c->rw;
    /* groups in chunks bounded by ids and by groups, so the device buffers stay small */
    size_t g0 = 0;
    while (g0 < n_groups) {
        size_t g1 = g0;
        while (g1 < n_groups && g1 - g0 < BFT_CHUNK_KMERS && (g1 == g0 || group_offs[g1 + 1] - group_offs[g0] <= 4 * BFT_CHUNK_KMERS)) g1++;
        const size_t m = g1 - g0, ids = (size_t)(group_offs[g1] - group_offs[g0]);
        ENSUREP(sl->d_cls, sl->cap_cls, (ids + 1) * sizeof(uint32_t));
        ENSUREP(sl->d_offs, sl->cap_offs, (m + 1) * sizeof(uint64_t));
        ENSUREP(sl->d_rows, sl->cap_rows, m * rw * sizeof(uint32_t));
        ENSUREP(sl->d_tile, sl->cap_tile, m * sizeof(uint32_t));
        if (ids) CKP(cudaMemcpyAsync(sl->d_cls, class_ids + group_offs[g0], ids * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        CKP(cudaMemcpyAsync(sl->d_offs, group_offs + g0, (m + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        /* offsets stay absolute: hand the kernel the id array shifted back by the chunk's first offset */
        int rc = bft_b200_annotation_setop_device(c, op, sl->d_cls - group_offs[g0], sl->d_offs, m, rows ? sl->d_rows : NULL, counts ? sl->d_tile : NULL);
        if (rc) return drain_ret(c, rc);
        if (rows) CKP(cudaMemcpyAsync(rows + g0 * rw, sl->d_rows, m * rw * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (counts) CKP(cudaMemcpyAsync(counts + g0, sl->d_tile, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CKP(cudaStreamSynchronize(st));
        g0 = g1;
    }
    return 0;
}

/* ---- sequences ---------------------------------------------------------------------------------------------- */
static int enqueue_sequences(bft_b200_ctx* c, cudaStream_t st, const char* d_chars, const uint64_t* d_offs, size_t n_seq, double thr,
                             int canonical, uint32_t* d_rows, uint8_t* d_status) {
    if (n_seq == 0) return 0;
    if (!c->seq_smem) return set_err(BFT_B200_ERR_ARG, "sequence queries: %d genomes exceed the shared-memory counter budget", c->G);
    size_t blocks = (n_seq + BFT_SEQ_WARPS - 1) / BFT_SEQ_WARPS;
    const size_t cap = (size_t)c->sm_count * 16;
    if (blocks > cap) blocks = cap;
#define BFT_L2(W_, G32_) k_query_sequences<W_, G32_><<<(int)blocks, 32 * BFT_SEQ_WARPS, c->seq_smem, st>>>(c->dview, d_chars, d_offs, n_seq, thr, canonical, \
                                                                                                      c->d_class_rows, c->rw, c->G, d_rows, d_status)
#define BFT_L(W_) do { if (c->rw == 1) BFT_L2(W_, true); else BFT_L2(W_, false); } while (0)
    BFT_BY_W(c->W, BFT_L);
#undef BFT_L
#undef BFT_L2
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

static int check_threshold(double thr) {
    if (!(thr > 0)) return set_err(BFT_B200_ERR_ARG, "query_sequence(): the threshold must be superior to 0.");
    if (thr > 1) return set_err(BFT_B200_ERR_ARG, "query_sequence(): the threshold must be inferior or equal to 1.");
    return 0;
}

extern "C" int bft_b200_query_sequences_device(bft_b200_ctx* c, const char* d_chars, const uint64_t* d_offs, size_t n_seq, double thr,
                                               int canonical, uint32_t* d_rows, uint8_t* d_status) {
    if (!c || !d_offs || !d_rows) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_sequences_device: NULL argument");
    int rc = check_threshold(thr);
    if (rc) return rc;
    CK(cudaSetDevice(c->device));
    return enqueue_sequences(c, c->streams[0], d_chars, d_offs, n_seq, thr, canonical, d_rows, d_status);
}

extern "C" int bft_b200_query_sequences(bft_b200_ctx* c, const char* chars, const uint64_t* offs, size_t n_seq, double thr, int canonical,
                                        uint32_t* rows, uint8_t* status) {
    if (!c || !offs || !rows) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_sequences: NULL argument");
    int rc = check_threshold(thr);
    if (rc) return drain_ret(c, rc);
    CK(cudaSetDevice(c->device));
    const size_t rw = (size_t)c->rw;
    size_t done = 0;
    int it = 0;
    while (done < n_seq) {
        size_t m = 0;
        while (done + m < n_seq && m < BFT_CHUNK_SEQS && (m == 0 || offs[done + m + 1] - offs[done] <= BFT_CHUNK_SEQ_CHARS)) m++;
        const uint64_t c0 = offs[done], c1 = offs[done + m];
        const int s = it & 1;
        slot_t* sl = &c->slot[s];
        cudaStream_t st = c->streams[s];
        CKP(cudaStreamSynchronize(st));
        ENSUREP(sl->d_in, sl->cap_in, (size_t)(c1 - c0) + 64);
        ENSUREP(sl->d_offs, sl->cap_offs, (m + 1) * sizeof(uint64_t));
        ENSUREP(sl->d_rows, sl->cap_rows, m * rw * sizeof(uint32_t));
        ENSUREP(sl->d_u8a, sl->cap_u8a, m);
        if (c1 > c0) CKP(cudaMemcpyAsync(sl->d_in, chars + c0, (size_t)(c1 - c0), cudaMemcpyHostToDevice, st));
        CKP(cudaMemcpyAsync(sl->d_offs, offs + done, (m + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        /* offsets stay absolute: hand the kernel a base pointer shifted back by the chunk's first offset */
        rc = enqueue_sequences(c, st, (const char*)sl->d_in - c0, sl->d_offs, m, thr, canonical, sl->d_rows, sl->d_u8a);
        if (rc) return drain_ret(c, rc);
        CKP(cudaMemcpyAsync(rows + done * rw, sl->d_rows, m * rw * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (status) CKP(cudaMemcpyAsync(status + done, sl->d_u8a, m, cudaMemcpyDeviceToHost, st));
        done += m;
        it++;
    }
    CKP(cudaStreamSynchronize(c->streams[0]));
    CKP(cudaStreamSynchronize(c->streams[1]));
    return 0;
}

/* ---- branching ---------------------------------------------------------------------------------------------- */
static int enqueue_branching(bft_b200_ctx* c, cudaStream_t st, const uint64_t* d_kmers, size_t n, uint8_t* d_succ, uint8_t* d_pred,
                             unsigned long long* d_count, uint32_t* d_nbr) {
    if (n == 0) return 0;
    const int grid = grid_for(c, (n + BFT_NBR_Q - 1) / BFT_NBR_Q, BFT_TPB); /* one lane per BFT_NBR_Q queries */
#define BFT_L(W_) k_query_branching<W_><<<grid, BFT_TPB, 0, st>>>(c->dview, d_kmers, n, d_succ, d_pred, d_count, d_nbr, c->ref_quirks)
    BFT_BY_W(c->W, BFT_L);
#undef BFT_L
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int bft_b200_query_branching_device(bft_b200_ctx* c, const uint64_t* d_kmers, size_t n, uint8_t* d_succ, uint8_t* d_pred,
                                               uint64_t* d_n_branching) {
    if (!c || (!d_kmers && n)) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_branching_device: NULL argument");
    CK(cudaSetDevice(c->device));
    if (d_n_branching) CK(cudaMemsetAsync(d_n_branching, 0, sizeof(uint64_t), c->streams[0]));
    return enqueue_branching(c, c->streams[0], d_kmers, n, d_succ, d_pred, (unsigned long long*)d_n_branching, NULL);
}

extern "C" int bft_b200_query_branching(bft_b200_ctx* c, const uint64_t* kmers, size_t n, uint8_t* succ, uint8_t* pred, uint64_t* n_branching) {
    if (!c || (!kmers && n)) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_branching: NULL argument");
    CK(cudaSetDevice(c->device));
    const size_t W = (size_t)c->W;
    CKP(cudaStreamSynchronize(c->streams[0]));
    CKP(cudaStreamSynchronize(c->streams[1]));
    CKP(cudaMemsetAsync(c->d_counter, 0, sizeof(unsigned long long), c->streams[0]));
    CKP(cudaStreamSynchronize(c->streams[0]));
    size_t done = 0;
    int it = 0;
    while (done < n) {
        const size_t m = n - done < BFT_CHUNK_KMERS ? n - done : BFT_CHUNK_KMERS;
        const int s = it & 1;
        slot_t* sl = &c->slot[s];
        cudaStream_t st = c->streams[s];
        CKP(cudaStreamSynchronize(st));
        ENSUREP(sl->d_in, sl->cap_in, m * W * 8);
        ENSUREP(sl->d_u8a, sl->cap_u8a, m);
        ENSUREP(sl->d_u8b, sl->cap_u8b, m);
        CKP(cudaMemcpyAsync(sl->d_in, kmers + done * W, m * W * 8, cudaMemcpyHostToDevice, st));
        int rc = enqueue_branching(c, st, (const uint64_t*)sl->d_in, m, sl->d_u8a, sl->d_u8b, c->d_counter, NULL);
        if (rc) return drain_ret(c, rc);
        if (succ) CKP(cudaMemcpyAsync(succ + done, sl->d_u8a, m, cudaMemcpyDeviceToHost, st));
        if (pred) CKP(cudaMemcpyAsync(pred + done, sl->d_u8b, m, cudaMemcpyDeviceToHost, st));
        done += m;
        it++;
    }
    CKP(cudaStreamSynchronize(c->streams[0]));
    CKP(cudaStreamSynchronize(c->streams[1]));
    if (n_branching) {
        unsigned long long h = 0;
        CKP(cudaMemcpy(&h, c->d_counter, sizeof h, cudaMemcpyDeviceToHost));
        *n_branching = h;
    }
    return 0;
}

extern "C" int bft_b200_set_reference_exact_branching(bft_b200_ctx* c, int exact) {
    if (!c) return set_err(BFT_B200_ERR_ARG, "bft_b200_set_reference_exact_branching: NULL context");
    c->ref_quirks = exact != 0;
    return 0;
}

extern "C" int bft_b200_query_neighbors(bft_b200_ctx* c, const uint64_t* kmers, size_t n, uint32_t* nbr) {
    if (!c || !nbr || (!kmers && n)) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_neighbors: NULL argument");
    CK(cudaSetDevice(c->device));
    const size_t W = (size_t)c->W;
    size_t done = 0;
    int it = 0;
    while (done < n) { /* two slots on two streams: the copies of one chunk overlap the look-ups of the other */
        const size_t m = n - done < BFT_CHUNK_KMERS ? n - done : BFT_CHUNK_KMERS;
        const int s = it & 1;
        slot_t* sl = &c->slot[s];
        cudaStream_t st = c->streams[s];
        CKP(cudaStreamSynchronize(st)); /* slot buffers free again */
        ENSUREP(sl->d_in, sl->cap_in, m * W * 8);
        ENSUREP(sl->d_rows, sl->cap_rows, m * 8 * sizeof(uint32_t));
        CKP(cudaMemcpyAsync(sl->d_in, kmers + done * W, m * W * 8, cudaMemcpyHostToDevice, st));
        int rc = enqueue_branching(c, st, (const uint64_t*)sl->d_in, m, NULL, NULL, NULL, sl->d_rows);
        if (rc) return drain_ret(c, rc);
        CKP(cudaMemcpyAsync(nbr + done * 8, sl->d_rows, m * 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        done += m;
        it++;
    }
    CKP(cudaStreamSynchronize(c->streams[0]));
    CKP(cudaStreamSynchronize(c->streams[1]));
    return 0;
}

extern "C" int bft_b200_kmer_walk_stats_device(bft_b200_ctx* c, const uint64_t* d_kmers, size_t n, uint64_t out[8]) {
    if (!c || !out || (!d_kmers && n)) return set_err(BFT_B200_ERR_ARG, "bft_b200_kmer_walk_stats_device: NULL argument");
    CK(cudaSetDevice(c->device));
    unsigned long long* d_acc = NULL;
    CK(cudaMalloc((void**)&d_acc, BFT_N_WALK_STATS * sizeof(unsigned long long)));
    cudaMemsetAsync(d_acc, 0, BFT_N_WALK_STATS * sizeof(unsigned long long), c->streams[0]);
    if (n) {
        const int grid = grid_for(c, n, BFT_TPB);
#define BFT_L(W_) k_kmer_walk_stats<W_><<<grid, BFT_TPB, 0, c->streams[0]>>>(c->dview, d_kmers, n, d_acc)
        BFT_BY_W(c->W, BFT_L);
#undef BFT_L
        c->launches++;
    }
    unsigned long long h[BFT_N_WALK_STATS] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(h, d_acc, sizeof h, cudaMemcpyDeviceToHost, c->streams[0]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->streams[0]);
    cudaFree(d_acc);
    if (e != cudaSuccess) return set_err(BFT_B200_ERR_CUDA, "k_kmer_walk_stats failed: %s", cudaGetErrorString(e));
    for (int j = 0; j < BFT_N_WALK_STATS; j++) out[j] = h[j];
    return 0;
}

extern "C" int bft_b200_random_gather_probe(bft_b200_ctx* c, size_t table_bytes, size_t n_loads, double* loads_per_sec) {
    if (!c || !loads_per_sec || table_bytes < 4096) return set_err(BFT_B200_ERR_ARG, "bft_b200_random_gather_probe: bad argument");
    CK(cudaSetDevice(c->device));
    uint64_t* table = NULL;
    unsigned long long* sink = NULL;
    CK(cudaMalloc((void**)&table, table_bytes));
    if (cudaMalloc((void**)&sink, 8) != cudaSuccess) { cudaFree(table); return set_err(BFT_B200_ERR_NOMEM, "cudaMalloc failed"); }
    cudaMemsetAsync(table, 1, table_bytes, c->streams[0]);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = grid_for(c, n_loads, BFT_TPB);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, c->streams[0]);
        k_random_gather<<<grid, BFT_TPB, 0, c->streams[0]>>>(table, table_bytes / 8, n_loads, sink);
        cudaEventRecord(e1, c->streams[0]);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
        c->launches++;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(table);
    cudaFree(sink);
    if (e != cudaSuccess) return set_err(BFT_B200_ERR_CUDA, "k_random_gather failed: %s", cudaGetErrorString(e));
    *loads_per_sec = (double)n_loads / (best * 1e-3);
    return 0;
}

/* ---- peer (NVLink) result buffers --------------------------------------------------------------------------- */
extern "C" int bft_b200_device_alloc(bft_b200_ctx* c, size_t bytes, void** d_ptr) {
    if (!c || !d_ptr) return set_err(BFT_B200_ERR_ARG, "bft_b200_device_alloc: NULL argument");
    CK(cudaSetDevice(c->device));
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) return set_err(BFT_B200_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    CK(cudaMemset(*d_ptr, 0, bytes ? bytes : 1));
    return 0;
}

extern "C" int bft_b200_device_free(bft_b200_ctx* c, void* d_ptr) {
    if (!c) return set_err(BFT_B200_ERR_ARG, "bft_b200_device_free: NULL context");
    CK(cudaSetDevice(c->device));
    CK(cudaFree(d_ptr));
    return 0;
}

extern "C" int bft_b200_peer_export(bft_b200_ctx* c, void* d_ptr, unsigned char handle[64]) {
    if (!c || !d_ptr || !handle) return set_err(BFT_B200_ERR_ARG, "bft_b200_peer_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    CK(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, 64);
    return 0;
}

extern "C" int bft_b200_peer_import(bft_b200_ctx* c, const unsigned char handle[64], void** d_ptr) {
    if (!c || !d_ptr || !handle) return set_err(BFT_B200_ERR_ARG, "bft_b200_peer_import: NULL argument");
    CK(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess)); /* maps the peer allocation; stores go over NVLink */
    return 0;
}

extern "C" int bft_b200_peer_close(bft_b200_ctx* c, void* d_ptr) {
    if (!c) return set_err(BFT_B200_ERR_ARG, "bft_b200_peer_close: NULL context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->streams[0]));
    CK(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}

/* ---- enumeration --------------------------------------------------------------------------------------------- */
static int enqueue_extract(bft_b200_ctx* c, uint64_t* d_kmers, uint32_t* d_cls, uint32_t* d_loc2vid) {
    cudaStream_t st = c->streams[0];
    if (c->n_pref) {
        const int tpb = 32 * BFT_EXTRACT_WARPS;
        const int grid = grid_for(c, c->n_pref * 32, tpb);
        const size_t smem = bft_extract_smem(c->W);
#define BFT_L(W_) k_extract_prefix_kmers<W_><<<grid, tpb, smem, st>>>(c->dview, c->n_pref, d_kmers, d_cls, d_loc2vid)
        BFT_BY_W(c->W, BFT_L);
#undef BFT_L
        c->launches++;
    }
#define BFT_L(W_) k_extract_uc_kmers<W_><<<grid_for(c, c->n_nodes, BFT_TPB), BFT_TPB, 0, st>>>(c->dview, c->n_nodes, d_kmers, d_cls, d_loc2vid)
    BFT_BY_W(c->W, BFT_L);
#undef BFT_L
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int bft_b200_extract_kmers_device(bft_b200_ctx* c, uint64_t* d_kmers, uint32_t* d_cls, size_t capacity) {
    if (!c || !d_kmers) return set_err(BFT_B200_ERR_ARG, "bft_b200_extract_kmers_device: NULL argument");
    if (capacity < c->stats.n_kmers) return set_err(BFT_B200_ERR_ARG, "bft_b200_extract_kmers: capacity %zu < %llu stored k-mers", capacity, (unsigned long long)c->stats.n_kmers);
    CK(cudaSetDevice(c->device));
    return enqueue_extract(c, d_kmers, d_cls, NULL);
}

extern "C" int bft_b200_extract_kmers(bft_b200_ctx* c, uint64_t* kmers, uint32_t* class_ids, uint32_t* rows, size_t capacity, uint64_t* n_written) {
    if (!c || !kmers) return set_err(BFT_B200_ERR_ARG, "bft_b200_extract_kmers: NULL argument");
    const size_t n = (size_t)c->stats.n_kmers, W = (size_t)c->W, rw = (size_t)c->rw;
    if (capacity < n) return set_err(BFT_B200_ERR_ARG, "bft_b200_extract_kmers: capacity %zu < %zu stored k-mers", capacity, n);
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->streams[0]));
    uint64_t* d_k = NULL;
    uint32_t* d_c = NULL;
    uint32_t* d_r = NULL;
    int rc = 0;
    if (cudaMalloc((void**)&d_k, (n + 1) * W * 8) != cudaSuccess || cudaMalloc((void**)&d_c, (n + 1) * 4) != cudaSuccess)
        rc = set_err(BFT_B200_ERR_NOMEM, "bft_b200_extract_kmers: cudaMalloc failed for %zu k-mers", n);
    if (!rc) rc = bft_b200_extract_kmers_device(c, d_k, d_c, n);
    if (!rc && cudaMemcpyAsync(kmers, d_k, n * W * 8, cudaMemcpyDeviceToHost, c->streams[0]) != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "copy failed");
    if (!rc && class_ids && cudaMemcpyAsync(class_ids, d_c, n * 4, cudaMemcpyDeviceToHost, c->streams[0]) != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "copy failed");
    if (!rc && rows && n) { /* rows in slices so the device buffer stays bounded */
        const size_t slice = BFT_CHUNK_KMERS;
        if (cudaMalloc((void**)&d_r, slice * rw * 4) != cudaSuccess) rc = set_err(BFT_B200_ERR_NOMEM, "bft_b200_extract_kmers: cudaMalloc failed");
        for (size_t done = 0; !rc && done < n; done += slice) {
            const size_t m = n - done < slice ? n - done : slice;
            k_expand_rows<<<grid_for(c, m * rw, BFT_TPB), BFT_TPB, 0, c->streams[0]>>>(d_c + done, m, c->d_class_rows, c->rw, d_r);
            c->launches++;
            if (cudaMemcpyAsync(rows + done * rw, d_r, m * rw * 4, cudaMemcpyDeviceToHost, c->streams[0]) != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "copy failed");
            if (!rc && cudaStreamSynchronize(c->streams[0]) != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "k_expand_rows failed");
        }
    }
    if (!rc) {
        cudaError_t e = cudaStreamSynchronize(c->streams[0]);
        if (e != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "extraction kernels failed: %s", cudaGetErrorString(e));
    }
    if (d_k) cudaFree(d_k);
    if (d_c) cudaFree(d_c);
    if (d_r) cudaFree(d_r);
    if (!rc && n_written) *n_written = n;
    return rc;
}

extern "C" int bft_b200_extract_kmers_file(bft_b200_ctx* c, const char* path, int compressed_output) {
    if (!c || !path) return set_err(BFT_B200_ERR_ARG, "bft_b200_extract_kmers_file: NULL argument");
    const size_t n = (size_t)c->stats.n_kmers, W = (size_t)c->W;
    uint64_t* km = (uint64_t*)malloc((n + 1) * W * 8);
    if (!km) return set_err(BFT_B200_ERR_NOMEM, "out of host memory");
    int rc = bft_b200_extract_kmers(c, km, NULL, NULL, n, NULL);
    if (!rc) {
        FILE* f = fopen(path, "w");
        if (!f) rc = set_err(BFT_B200_ERR_FILE, "extract_kmers_to_disk(): failed to create/open output file %s", path);
        else {
            const int k = c->k;
            const size_t nb = (size_t)(2 * k + 7) / 8;
            if (compressed_output) {
                fprintf(f, "%d\n%d\n", k, (int)n);
                for (size_t i = 0; i < n; i++) fwrite(km + i * W, 1, nb, f);
            } else {
                char line[132];
                for (size_t i = 0; i < n; i++) {
                    for (int j = 0; j < k; j++) line[j] = "ACGT"[(km[i * W + (size_t)(j >> 5)] >> (2 * (j & 31))) & 3];
                    line[k] = '\n';
                    fwrite(line, 1, (size_t)k + 1, f);
                }
            }
            fclose(f);
        }
    }
    free(km);
    return rc;
}

/* ---- file-level drivers ------------------------------------------------------------------------------------- */
extern "C" int bft_b200_query_kmers_file(bft_b200_ctx* c, const char* query_path, int binary_file, const char* csv_path, uint64_t* n_present) {
    if (!c || !query_path || !csv_path) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_kmers_file: NULL argument");
    uint64_t* q = NULL;
    char* ascii = NULL;
    size_t n = 0;
    /* "kmers_comp": packed records. "kmers": the text lines go to the GPU as they are — parseKmerCount (src/fasta.c:3-53)
     * runs there (k_encode_ascii) and flags the lines the reference would drop */
    if (binary_file ? bft_read_kmer_file(query_path, 1, c->k, c->W, &q, &n) : bft_read_kmer_text_file(query_path, c->k, &ascii, &n))
        return set_err(BFT_B200_ERR_FILE, "cannot read k-mer file %s", query_path);
    uint8_t* present = (uint8_t*)malloc(n + 1);
    uint8_t* valid = binary_file ? NULL : (uint8_t*)malloc(n + 1);
    uint32_t* rows = (uint32_t*)malloc((n + 1) * (size_t)c->rw * sizeof(uint32_t));
    int rc = (!present || !rows || (!binary_file && !valid)) ? set_err(BFT_B200_ERR_NOMEM, "out of host memory")
             : binary_file ? bft_b200_query_kmers(c, q, n, present, rows, NULL)
                           : bft_b200_query_kmers_ascii(c, ascii, n, valid, present, rows, NULL);
    if (!rc) {
        size_t m = n;
        if (valid) { /* rejected lines produce no CSV row (src/file_io.c:786-862) */
            m = 0;
            for (size_t i = 0; i < n; i++) {
                if (!valid[i]) continue;
                if (m != i) {
                    memcpy(rows + m * (size_t)c->rw, rows + i * (size_t)c->rw, (size_t)c->rw * sizeof(uint32_t));
                    present[m] = present[i];
                }
                m++;
            }
        }
        FILE* f = fopen(csv_path, "w");
        if (!f) rc = set_err(BFT_B200_ERR_FILE, "cannot write %s", csv_path);
        else {
            if (bft_csv_write_header(f, c->names, c->G) || bft_csv_write_rows(f, rows, m, c->G, c->rw) || bft_csv_finish(f))
                rc = set_err(BFT_B200_ERR_FILE, "could not write output to CSV file %s", csv_path);
            fclose(f);
        }
        uint64_t np = 0;
        for (size_t i = 0; i < m; i++) np += present[i];
        if (n_present) *n_present = np;
    }
    free(q); free(ascii); free(present); free(valid); free(rows);
    return rc;
}

extern "C" int bft_b200_query_branching_file(bft_b200_ctx* c, const char* query_path, int binary_file, uint64_t* n_branching) {
    if (!c || !query_path) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_branching_file: NULL argument");
    uint64_t* q = NULL;
    size_t n = 0;
    if (bft_read_kmer_file(query_path, binary_file, c->k, c->W, &q, &n)) return set_err(BFT_B200_ERR_FILE, "cannot read k-mer file %s", query_path);
    int rc = bft_b200_query_branching(c, q, n, NULL, NULL, n_branching);
    free(q);
    return rc;
}

extern "C" int bft_b200_query_sequences_file(bft_b200_ctx* c, const char* query_path, const char* csv_path, double thr, int canonical) {
    if (!c || !query_path || !csv_path) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_sequences_file: NULL argument");
    char* chars = NULL;
    uint64_t* offs = NULL;
    size_t n = 0;
    if (bft_read_sequence_file(query_path, &chars, &offs, &n)) return set_err(BFT_B200_ERR_FILE, "cannot read sequence file %s", query_path);
    uint32_t* rows = (uint32_t*)malloc((n + 1) * (size_t)c->rw * sizeof(uint32_t));
    uint8_t* status = (uint8_t*)malloc(n + 1);
    int rc = (!rows || !status) ? set_err(BFT_B200_ERR_NOMEM, "out of host memory") : bft_b200_query_sequences(c, chars, offs, n, thr, canonical, rows, status);
    if (!rc) {
        for (size_t i = 0; i < n && !rc; i++)
            if (status[i] == BFT_B200_SEQ_BAD_CHAR) rc = set_err(BFT_B200_ERR_ARG, "get_kmer(): Unexpected character encountered in k-mer (sequence %zu).", i);
    }
    if (!rc) {
        FILE* f = fopen(csv_path, "w");
        if (!f) rc = set_err(BFT_B200_ERR_FILE, "cannot write %s", csv_path);
        else {
            if (bft_csv_write_header(f, c->names, c->G) || bft_csv_write_rows(f, rows, n, c->G, c->rw) || bft_csv_finish(f))
                rc = set_err(BFT_B200_ERR_FILE, "could not write output to CSV file %s", csv_path);
            fclose(f);
        }
    }
    free(chars); free(offs); free(rows); free(status);
    return rc;
}

/* ---- graph traversals (reference src/snippets.c) --------------------------------------------------------------- */
extern "C" int bft_b200_graph_release(bft_b200_ctx* c) {
    if (!c) return set_err(BFT_B200_ERR_ARG, "bft_b200_graph_release: NULL context");
    cudaSetDevice(c->device);
    if (c->graph.d_vk) cudaFree(c->graph.d_vk);
    if (c->graph.d_vcls) cudaFree(c->graph.d_vcls);
    if (c->graph.d_adj) cudaFree(c->graph.d_adj);
    if (c->graph.d_loc2vid) cudaFree(c->graph.d_loc2vid);
    memset(&c->graph, 0, sizeof c->graph);
    if (c->pool) { /* hand the traversal scratch back to the device */
        cudaStreamSynchronize(c->streams[0]);
        cudaMemPoolTrimTo(c->pool, 0);
    }
    return 0;
}

/* scratch device arrays of one traversal call, freed together. Stream-ordered allocations from the context's OWN
 * memory pool (release threshold unlimited, so after the first call the blocks come back without a driver round trip;
 * bft_b200_graph_release trims it). The device's default pool — which the host application may share — is not touched. */
static int ensure_pool(bft_b200_ctx* c) {
    if (c->pool) return 0;
    cudaMemPoolProps props;
    memset(&props, 0, sizeof props);
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = c->device;
    if (cudaMemPoolCreate(&c->pool, &props) != cudaSuccess) {
        c->pool = NULL;
        return set_err(BFT_B200_ERR_CUDA, "graph traversal: cudaMemPoolCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    uint64_t keep = ~0ULL;
    cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    return 0;
}

struct dev_scratch {
    void* p[24];
    int n;
    cudaStream_t st;
    cudaMemPool_t pool;
    explicit dev_scratch(bft_b200_ctx* c) : n(0), st(c->streams[0]), pool(c->pool) {}
    ~dev_scratch() { for (int i = 0; i < n; i++) cudaFreeAsync(p[i], st); }
    template <typename T>
    int get(T** out, size_t count) {
        void* q = NULL;
        const size_t bytes = (count ? count : 1) * sizeof(T) + 32;
        if (n >= 24 || !pool || cudaMallocFromPoolAsync(&q, bytes, pool, st) != cudaSuccess) {
            (void)cudaGetLastError();
            return set_err(BFT_B200_ERR_NOMEM, "graph traversal: cudaMallocAsync(%zu) failed", bytes);
        }
        p[n++] = q;
        *out = (T*)q;
        return 0;
    }
};

extern "C" int bft_b200_graph_prepare(bft_b200_ctx* c) {
    if (!c) return set_err(BFT_B200_ERR_ARG, "bft_b200_graph_prepare: NULL context");
    if (c->graph.ready) return 0;
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->stats.n_kmers, W = (size_t)c->W;
    const uint64_t n_loc = c->n_loc; /* 64-bit count of the untruncated parts (bft_b200_open) */
    if (n >= BFT_V_NONE || n_loc >= BFT_V_NONE)
        return set_err(BFT_B200_ERR_ARG, "graph traversal: %zu k-mers / %llu storage locations exceed the 32-bit vertex ids", n, (unsigned long long)n_loc);
    cudaStream_t st = c->streams[0];
    CK(cudaStreamSynchronize(st));
    bft_b200_graph_release(c);
    if (cudaMalloc((void**)&c->graph.d_vk, (n + 1) * W * 8) != cudaSuccess || cudaMalloc((void**)&c->graph.d_vcls, (n + 1) * 4) != cudaSuccess ||
        cudaMalloc((void**)&c->graph.d_adj, (n + 1) * 8 * 4) != cudaSuccess || cudaMalloc((void**)&c->graph.d_loc2vid, ((size_t)n_loc + 1) * 4) != cudaSuccess) {
        bft_b200_graph_release(c);
        return set_err(BFT_B200_ERR_NOMEM, "graph traversal: cudaMalloc failed for %zu vertices", n);
    }
    uint32_t* d_loc2vid = c->graph.d_loc2vid;
    c->graph.n = n;
    c->graph.bytes = (n + 1) * (W * 8 + 4 + 32) + ((size_t)n_loc + 1) * 4;
    CK(cudaMemsetAsync(d_loc2vid, 0xff, ((size_t)n_loc + 1) * 4, st));
    int rc = enqueue_extract(c, c->graph.d_vk, c->graph.d_vcls, d_loc2vid);
    if (!rc && n) {
        const int grid = grid_for(c, (n + BFT_NBR_Q - 1) / BFT_NBR_Q, BFT_TPB);
#define BFT_L(W_) k_graph_adjacency<W_><<<grid, BFT_TPB, 0, st>>>(c->dview, c->graph.d_vk, n, d_loc2vid, c->graph.d_adj)
        BFT_BY_W(c->W, BFT_L);
#undef BFT_L
        c->launches++;
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (!rc && e != cudaSuccess) rc = set_err(BFT_B200_ERR_CUDA, "graph construction failed: %s", cudaGetErrorString(e));
    if (rc) { bft_b200_graph_release(c); return rc; }
    c->graph.ready = 1;
    return ensure_pool(c);
}

extern "C" int bft_b200_query_vertex_ids(bft_b200_ctx* c, const uint64_t* kmers, size_t n, uint32_t* vertex_ids) {
    if (!c || !vertex_ids || (!kmers && n)) return set_err(BFT_B200_ERR_ARG, "bft_b200_query_vertex_ids: NULL argument");
    int rc = bft_b200_graph_prepare(c);
    if (rc) return rc;
    const size_t W = (size_t)c->W;
    size_t done = 0;
    int it = 0;
    while (done < n) { /* two slots on two streams, as the other host pipelines */
        const size_t m = n - done < BFT_CHUNK_KMERS ? n - done : BFT_CHUNK_KMERS;
        const int s = it & 1;
        slot_t* sl = &c->slot[s];
        cudaStream_t st = c->streams[s];
        CKP(cudaStreamSynchronize(st));
        ENSUREP(sl->d_in, sl->cap_in, m * W * 8);
        ENSUREP(sl->d_cls, sl->cap_cls, m * sizeof(uint32_t));
        CKP(cudaMemcpyAsync(sl->d_in, kmers + done * W, m * W * 8, cudaMemcpyHostToDevice, st));
#define BFT_L(W_) k_query_vertex_ids<W_><<<grid_for(c, m, BFT_TPB), BFT_TPB, 0, st>>>(c->dview, (const uint64_t*)sl->d_in, m, c->graph.d_loc2vid, (uint32_t*)sl->d_cls)
        BFT_BY_W(c->W, BFT_L);
#undef BFT_L
        c->launches++;
        CKP(cudaMemcpyAsync(vertex_ids + done, sl->d_cls, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        done += m;
        it++;
    }
    CKP(cudaStreamSynchronize(c->streams[0]));
    CKP(cudaStreamSynchronize(c->streams[1]));
    return 0;
}

extern "C" int bft_b200_graph_adjacency(bft_b200_ctx* c, uint32_t* adj, size_t capacity) {
    if (!c || !adj) return set_err(BFT_B200_ERR_ARG, "bft_b200_graph_adjacency: NULL argument");
    if (capacity < c->stats.n_kmers) return set_err(BFT_B200_ERR_ARG, "bft_b200_graph_adjacency: capacity %zu < %llu stored k-mers", capacity, (unsigned long long)c->stats.n_kmers);
    int rc = bft_b200_graph_prepare(c);
    if (rc) return rc;
    CK(cudaMemcpy(adj, c->graph.d_adj, c->graph.n * 8 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int bft_b200_connected_components(bft_b200_ctx* c, const uint32_t* genome_ids, int n_ids, uint64_t* n_components, uint32_t* labels) {
    if (!c || !n_components || n_ids < 0 || (n_ids && !genome_ids)) return set_err(BFT_B200_ERR_ARG, "bft_b200_connected_components: bad argument");
    for (int i = 1; i < n_ids; i++)
        if (genome_ids[i] <= genome_ids[i - 1])
            return set_err(BFT_B200_ERR_ARG, "bft_b200_connected_components: genome ids must be strictly ascending (is_in_subgraph() walks them in order)");
    int rc = bft_b200_graph_prepare(c);
    if (rc) return rc;
    const size_t n = c->graph.n;
    cudaStream_t st = c->streams[0];
    *n_components = 0;
    if (n_ids && genome_ids[n_ids - 1] >= (uint32_t)c->G) { /* no k-mer carries a genome that was never inserted */
        if (labels) memset(labels, 0xff, n * sizeof(uint32_t));
        return 0;
    }
    dev_scratch tmp(c);
    uint32_t *d_parent = NULL, *d_labels = NULL, *d_want = NULL;
    uint8_t* d_in = NULL;
    unsigned long long* d_cnt = NULL;
    if ((rc = tmp.get(&d_parent, n)) || (rc = tmp.get(&d_cnt, 1)) || (labels && (rc = tmp.get(&d_labels, n)))) return rc;
    if (n_ids) {
        if ((rc = tmp.get(&d_want, (size_t)c->rw)) || (rc = tmp.get(&d_in, c->n_classes + 1))) return rc;
        CK(cudaMemsetAsync(d_want, 0, (size_t)c->rw * 4, st));
        for (int i = 0; i < n_ids; i++) { /* a handful of ids: set their bits one by one */
            uint32_t word = 0;
            for (int j = 0; j < n_ids; j++)
                if ((genome_ids[j] >> 5) == (genome_ids[i] >> 5)) word |= 1u << (genome_ids[j] & 31);
            CK(cudaMemcpyAsync(d_want + (genome_ids[i] >> 5), &word, 4, cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
        }
        k_graph_class_filter<<<grid_for(c, c->n_classes, BFT_TPB), BFT_TPB, 0, st>>>(c->d_class_rows, c->rw, c->n_classes, d_want, d_in);
        c->launches++;
    }
    CK(cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), st));
    if (n) {
        const int grid = grid_for(c, n, BFT_TPB);
        k_graph_iota<<<grid, BFT_TPB, 0, st>>>(d_parent, n);
        k_graph_hook<<<grid, BFT_TPB, 0, st>>>(c->graph.d_adj, c->graph.d_vcls, d_in, n, d_parent);
        k_graph_labels<<<grid, BFT_TPB, 0, st>>>(d_parent, c->graph.d_vcls, d_in, n, d_labels, d_cnt);
        c->launches += 3;
    }
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, d_cnt, sizeof h, cudaMemcpyDeviceToHost, st));
    if (labels && n) CK(cudaMemcpyAsync(labels, d_labels, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_components = h;
    return 0;
}

extern "C" void bft_b200_free(void* p) { free(p); }

extern "C" int bft_b200_simple_paths(bft_b200_ctx* c, double core_ratio, char** paths, size_t* n_bytes, uint64_t* n_paths, uint64_t* longest) {
    if (!c || !paths || !n_bytes) return set_err(BFT_B200_ERR_ARG, "bft_b200_simple_paths: NULL argument");
    if (!(core_ratio >= 0.0 && core_ratio <= 1.0)) return set_err(BFT_B200_ERR_ARG, "bft_b200_simple_paths: core_ratio %g outside [0, 1]", core_ratio);
    *paths = NULL;
    *n_bytes = 0;
    int rc = bft_b200_graph_prepare(c);
    if (rc) return rc;
    const size_t n = c->graph.n;
    const uint32_t core = (uint32_t)(int)(core_ratio * c->G); /* nb_genomes_core, src/snippets.c:366 */
    cudaStream_t st = c->streams[0];
    dev_scratch tmp(c);
    uint8_t* d_chain = NULL;
    uint32_t *d_usucc = NULL, *d_next = NULL, *d_prev = NULL, *d_to[2] = {NULL, NULL}, *d_dist[2] = {NULL, NULL}, *d_low[2] = {NULL, NULL};
    unsigned long long *d_size = NULL, *d_offs = NULL, *d_stats = NULL;
    int* d_pending = NULL;
    if ((rc = tmp.get(&d_chain, n)) || (rc = tmp.get(&d_usucc, n)) || (rc = tmp.get(&d_next, n)) || (rc = tmp.get(&d_prev, n)) ||
        (rc = tmp.get(&d_to[0], n)) || (rc = tmp.get(&d_to[1], n)) || (rc = tmp.get(&d_dist[0], n)) || (rc = tmp.get(&d_dist[1], n)) ||
        (rc = tmp.get(&d_size, n + 1)) || (rc = tmp.get(&d_offs, n + 1)) || (rc = tmp.get(&d_stats, 2)) || (rc = tmp.get(&d_pending, 1)))
        return rc;
    const int grid = grid_for(c, n ? n : 1, BFT_TPB);
    CK(cudaMemsetAsync(d_prev, 0xff, (n + 1) * 4, st));
    CK(cudaMemsetAsync(d_size, 0, (n + 1) * 8, st));
    CK(cudaMemsetAsync(d_stats, 0, 16, st));
    k_paths_vertices<<<grid, BFT_TPB, 0, st>>>(c->graph.d_adj, c->graph.d_vcls, c->d_class_counts, core, n, d_chain, d_usucc);
    k_paths_link<<<grid, BFT_TPB, 0, st>>>(d_chain, d_usucc, c->graph.d_vcls, c->d_class_rows, c->rw, core, n, d_next, d_prev);
    c->launches += 2;
    int max_rounds = 2;
    while (((size_t)1 << (max_rounds - 2)) < n + 1) max_rounds++; /* ceil(log2(n + 1)) + 2 */
    int cur = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        /* attempt 0: rank the chains; pointers still short of a head after max_rounds sit on cycles of chain vertices:
         * redo the doubling carrying the smallest vertex id, open each cycle there, rank again */
        cur = 0;
        k_paths_rank_init<<<grid, BFT_TPB, 0, st>>>(d_chain, d_prev, n, d_to[0], d_dist[0], (uint32_t*)NULL);
        c->launches++;
        int pending = 1;
        for (int r = 0; r < max_rounds && pending; r++) {
            CK(cudaMemsetAsync(d_pending, 0, sizeof(int), st));
            k_paths_rank_step<<<grid, BFT_TPB, 0, st>>>(d_to[cur], d_dist[cur], (const uint32_t*)NULL, d_prev, n, d_to[cur ^ 1], d_dist[cur ^ 1],
                                                        (uint32_t*)NULL, d_pending);
            c->launches++;
            CK(cudaMemcpyAsync(&pending, d_pending, sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            cur ^= 1;
        }
        if (!pending) break;
        if (attempt == 1) return set_err(BFT_B200_ERR_CUDA, "bft_b200_simple_paths: chain ranking did not converge");
        uint8_t* d_cut = NULL;
        if ((rc = tmp.get(&d_low[0], n)) || (rc = tmp.get(&d_low[1], n)) || (rc = tmp.get(&d_cut, n))) return rc;
        int lc = 0;
        k_paths_rank_init<<<grid, BFT_TPB, 0, st>>>(d_chain, d_prev, n, d_to[0], d_dist[0], d_low[0]);
        for (int r = 0; r < max_rounds; r++, lc ^= 1)
            k_paths_rank_step<<<grid, BFT_TPB, 0, st>>>(d_to[lc], d_dist[lc], d_low[lc], d_prev, n, d_to[lc ^ 1], d_dist[lc ^ 1], d_low[lc ^ 1], d_pending);
        k_paths_find_cuts<<<grid, BFT_TPB, 0, st>>>(d_chain, d_to[lc], d_low[lc], d_prev, n, d_cut);
        k_paths_apply_cuts<<<grid, BFT_TPB, 0, st>>>(d_cut, n, d_next, d_prev);
        c->launches += 3 + (uint64_t)max_rounds;
    }
    k_paths_sizes<<<grid, BFT_TPB, 0, st>>>(d_chain, d_next, d_to[cur], d_dist[cur], n, c->k, d_size, d_stats);
    c->launches++;
    {
        void* d_tmp = NULL;
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_size, d_offs, (int64_t)(n + 1), st);
        unsigned char* d_scan = NULL;
        if ((rc = tmp.get(&d_scan, tmp_bytes))) return rc;
        CK(cub::DeviceScan::ExclusiveSum((void*)d_scan, tmp_bytes, d_size, d_offs, (int64_t)(n + 1), st));
    }
    unsigned long long total = 0, stats[2] = {0, 0};
    CK(cudaMemcpyAsync(&total, d_offs + n, 8, cudaMemcpyDeviceToHost, st)); /* size[n] == 0, so offs[n] is the grand total */
    CK(cudaMemcpyAsync(stats, d_stats, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    char* d_out = NULL;
    if ((rc = tmp.get(&d_out, (size_t)total))) return rc;
    if (n) {
#define BFT_L(W_) k_paths_write<W_><<<grid, BFT_TPB, 0, st>>>(d_chain, d_next, d_to[cur], d_dist[cur], d_offs, c->graph.d_vk, n, c->k, d_out)
        BFT_BY_W(c->W, BFT_L);
#undef BFT_L
        c->launches++;
    }
    char* h = (char*)malloc((size_t)total + 1);
    if (!h) return set_err(BFT_B200_ERR_NOMEM, "bft_b200_simple_paths: out of host memory (%llu bytes)", total);
    cudaError_t e = cudaMemcpyAsync(h, d_out, (size_t)total, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { free(h); return set_err(BFT_B200_ERR_CUDA, "bft_b200_simple_paths: %s", cudaGetErrorString(e)); }
    h[total] = 0;
    *paths = h;
    *n_bytes = (size_t)total;
    if (n_paths) *n_paths = stats[0];
    if (longest) *longest = stats[1];
    return 0;
}

extern "C" int bft_b200_simple_paths_file(bft_b200_ctx* c, double core_ratio, const char* out_path, uint64_t* n_paths, uint64_t* longest) {
    if (!c || !out_path) return set_err(BFT_B200_ERR_ARG, "bft_b200_simple_paths_file: NULL argument");
    char* buf = NULL;
    size_t nb = 0;
    int rc = bft_b200_simple_paths(c, core_ratio, &buf, &nb, n_paths, longest);
    if (rc) return rc;
    FILE* f = fopen(out_path, "w");
    if (!f) rc = set_err(BFT_B200_ERR_FILE, "extract_simple_paths_to_disk(): failed to create output file %s", out_path);
    else {
        if (fwrite(buf, 1, nb, f) != nb) rc = set_err(BFT_B200_ERR_FILE, "could not write %s", out_path);
        fclose(f);
    }
    free(buf);
    return rc;
}
