/* arena_accel_check — TEST TOOL (host, no GPU). The engine builds three accelerator tables on the DEVICE from the arena's own
 * enumeration (bft_b200.cu: k_kf_insert, k_rkf_fill_entries / k_rkf_insert, k_deep_count / k_deep_insert); the look-up code that
 * reads them lives in bft_arena.h and compiles for the host as well. This tool rebuilds the same tables sequentially on the host —
 * same hash functions, same block sizing, same slot format — and checks that a look-up through them (bft_lookup_loc's fast path:
 * fused root directory + filter, collapsed subtrees, stored-k-mer filter) answers exactly as the walk over the structure does:
 *   arena_accel_check file.bft queries.kc sectors_per_prefix tight
 * for every stored k-mer (found, same class) and every query of the file (same answer), flags 0 and BFT_LK_PRESENCE. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bft_flatten.h"
#include "bft_io.h"

static int W;

static void or_shl(uint64_t* dst, const uint64_t* src, int sh) { /* dst |= src << sh over BFT_MAX_WORDS words */
    const int ws = sh >> 6, bs = sh & 63;
    for (int w = BFT_MAX_WORDS - 1; w >= 0; w--) {
        uint64_t x = 0;
        if (w - ws >= 0) x = src[w - ws] << bs;
        if (bs && w - ws - 1 >= 0) x |= src[w - ws - 1] >> (64 - bs);
        dst[w] |= x;
    }
}

typedef struct { uint64_t km[BFT_MAX_WORDS]; uint32_t cls; } stored_t;

/* every stored k-mer with its class, in any order (the walk of k_extract_prefix_kmers / k_extract_uc_kmers without the ranking) */
static size_t enumerate(const bft_arena_t* a, stored_t* out) {
    const uint64_t top_mask = a->cls_shift ? ((1ULL << a->cls_shift) - 1ULL) : ~BFT_SLOT_SPECIAL;
    size_t n = 0;
    for (size_t j = 0; j < a->n_pref; j++) {
        const bft_entry_t e = a->pref[j];
        const uint32_t kind = e.b >> BFT_KIND_SHIFT;
        if (kind != BFT_KIND_INLINE && kind != BFT_KIND_LEAF) continue;
        const bft_path_t* path = &a->node_path[a->pref_node[j]];
        uint64_t base[BFT_MAX_WORDS], lw[BFT_MAX_WORDS] = {0};
        memcpy(base, path->acc, sizeof base);
        lw[0] = a->pref_low18[j];
        or_shl(base, lw, (int)(BFT_PREFIX_BITS * path->depth));
        if (kind == BFT_KIND_LEAF) {
            memcpy(out[n].km, base, sizeof base);
            out[n++].cls = e.a;
            continue;
        }
        const uint32_t n_slots = BFT_BUCKET_KEYS * BFT_INLINE_NBK(e);
        for (uint32_t s = 0; s < n_slots; s++) {
            const size_t gslot = (size_t)e.a * BFT_BUCKET_KEYS + s;
            const uint64_t* p = a->buckets + gslot * W;
            const uint64_t top = p[W - 1];
            if (top == BFT_SLOT_EMPTY) continue;
            uint32_t m = 1, start = 0;
            const int is_ovf = (top & BFT_SLOT_SPECIAL) != 0;
            if (is_ovf) { m = (uint32_t)(top >> 32) & 0x7fffffffu; start = (uint32_t)top; }
            for (uint32_t i = 0; i < m; i++) {
                const uint64_t* q = is_ovf ? a->ovf + ((size_t)start + i) * W : p;
                uint64_t key[BFT_MAX_WORDS] = {0};
                for (int w = 0; w < W; w++) key[w] = q[w];
                const uint32_t cls = a->cls_shift ? ((uint32_t)(key[W - 1] >> a->cls_shift) & a->cls_mask)
                                                  : (is_ovf ? a->ovfcls[start + i] : a->slotcls[gslot]);
                key[W - 1] &= top_mask;
                memcpy(out[n].km, base, sizeof base);
                or_shl(out[n].km, key, BFT_PREFIX_BITS * (int)(path->depth + 1));
                out[n++].cls = cls;
            }
        }
    }
    for (size_t nid = 0; nid < a->n_nodes; nid++) {
        const bft_node_t* nd = &a->nodes[nid];
        const bft_path_t* path = &a->node_path[nid];
        for (uint32_t i = 0; i < nd->uc_n; i++) {
            uint64_t key[BFT_MAX_WORDS] = {0};
            for (int w = 0; w < W; w++) key[w] = a->uckeys[((size_t)nd->uc_begin + i) * W + w];
            memcpy(out[n].km, path->acc, sizeof out[n].km);
            or_shl(out[n].km, key, BFT_PREFIX_BITS * (int)path->depth);
            out[n++].cls = a->uccls[nd->uc_begin + i];
        }
    }
    return n;
}

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s file.bft queries.kc sectors_per_prefix tight\n", argv[0]); return 2; }
    char err[256];
    bft_arena_t* a = bft_arena_from_file(argv[1], err, sizeof err);
    if (!a) { fprintf(stderr, "%s\n", err); return 1; }
    W = a->W;
    const uint32_t S = (uint32_t)atoi(argv[3]);
    const size_t per_bucket = atoi(argv[4]) ? BFT_BUCKET_KEYS : BFT_BUCKET_KEYS / 2;
    bft_view_t plain, fast;
    bft_arena_view(a, &plain);
    fast = plain;
    stored_t* st = calloc(a->n_kmers + 1, sizeof(stored_t));
    const size_t n = enumerate(a, st);
    if (n != a->n_kmers) { fprintf(stderr, "enumerated %zu of %zu k-mers\n", n, (size_t)a->n_kmers); return 1; }

    /* --- collapsed subtrees (build_deep_blocks + k_deep_insert) */
    uint32_t* cnt = calloc(BFT_ROOTDIR_SIZE, 4);
    for (size_t i = 0; i < n; i++) {
        const uint32_t p = (uint32_t)st[i].km[0] & (BFT_ROOTDIR_SIZE - 1u);
        if ((a->rootdir[p].b >> BFT_KIND_SHIFT) == BFT_KIND_NODE) cnt[p]++;
    }
    bft_entry_t* fastdir = malloc(BFT_ROOTDIR_SIZE * sizeof(bft_entry_t));
    memcpy(fastdir, a->rootdir, BFT_ROOTDIR_SIZE * sizeof(bft_entry_t));
    size_t n_db = 0, n_deep_prefixes = 0;
    for (size_t p = 0; p < BFT_ROOTDIR_SIZE; p++) {
        if ((a->rootdir[p].b >> BFT_KIND_SHIFT) != BFT_KIND_NODE || cnt[p] == 0) continue;
        uint32_t lb = 0;
        while (lb < BFT_LB_MASK && ((size_t)1 << lb) * per_bucket < (size_t)cnt[p]) lb++;
        if (((size_t)1 << lb) * per_bucket < (size_t)cnt[p]) continue;
        fastdir[p].a = (uint32_t)n_db;
        fastdir[p].b = (BFT_KIND_DEEP << BFT_KIND_SHIFT) | (lb << BFT_LB_SHIFT) | (cnt[p] > 0xffffffu ? 0xffffffu : cnt[p]);
        n_db += (size_t)1 << lb;
        n_deep_prefixes++;
    }
    uint64_t* db = NULL;
    uint32_t* dcls = NULL;
    size_t deep_kmers = 0, probes_moved = 0;
    if (n_db) {
        db = malloc(n_db * BFT_BUCKET_KEYS * W * 8);
        memset(db, 0xff, n_db * BFT_BUCKET_KEYS * W * 8);
        dcls = malloc(n_db * BFT_BUCKET_KEYS * 4);
        memset(dcls, 0xff, n_db * BFT_BUCKET_KEYS * 4);
        for (size_t i = 0; i < n; i++) {
            const bft_entry_t e = fastdir[(uint32_t)st[i].km[0] & (BFT_ROOTDIR_SIZE - 1u)];
            if ((e.b >> BFT_KIND_SHIFT) != BFT_KIND_DEEP) continue;
            uint64_t key[BFT_MAX_WORDS];
            memcpy(key, st[i].km, sizeof key);
            bft_shift18(key, W);
            const uint32_t lb = (e.b >> BFT_LB_SHIFT) & BFT_LB_MASK, mask = (1u << lb) - 1u;
            const uint64_t top = key[W - 1] | (a->cls_shift ? (uint64_t)st[i].cls << a->cls_shift : 0ULL);
            uint32_t b = bft_bucket_of(key, W, lb);
            int placed = 0;
            for (uint32_t probe = 0; probe <= mask && !placed; probe++, b = (b + 1u) & mask) {
                for (int j = 0; j < BFT_BUCKET_KEYS && !placed; j++) {
                    uint64_t* slot = db + (((size_t)e.a + b) * BFT_BUCKET_KEYS + j) * W;
                    if (slot[W - 1] != BFT_SLOT_EMPTY) continue;
                    for (int w = 0; w < W - 1; w++) slot[w] = key[w];
                    slot[W - 1] = top;
                    if (!a->cls_shift) dcls[((size_t)e.a + b) * BFT_BUCKET_KEYS + j] = st[i].cls;
                    placed = 1;
                    probes_moved += probe != 0;
                }
            }
            if (!placed) { fprintf(stderr, "no slot for a k-mer of a collapsed block\n"); return 1; }
            deep_kmers++;
        }
        fast.rootdir_fast = fastdir;
        fast.dbuckets = db;
        fast.dslotcls = dcls;
    }
    /* no leaf-level Node: the successor quirk never applies (what bft_b200_open passes as quirk_safe) */
    fast.kf_quirk_safe = (uint32_t)(a->max_depth < a->k / BFT_NB_CHAR_SUF_PREF || a->k == BFT_NB_CHAR_SUF_PREF);

    /* --- fused root directory + filter (k_rkf_fill_entries + k_rkf_insert) */
    uint64_t* rkf = NULL;
    if (S) {
        rkf = calloc((size_t)BFT_ROOTDIR_SIZE * S * 4 + 4, 8);
        const bft_entry_t* rd = fast.rootdir_fast ? fast.rootdir_fast : a->rootdir;
        for (size_t i = 0; i < (size_t)BFT_ROOTDIR_SIZE * S; i++) rkf[i * 4] = (uint64_t)rd[i / S].a | ((uint64_t)rd[i / S].b << 32);
        for (size_t i = 0; i < n; i++) {
            const bft_rkf_pos_t q = bft_rkf_pos(st[i].km, W, S);
            uint64_t* p = rkf + ((size_t)((uint32_t)st[i].km[0] & (BFT_ROOTDIR_SIZE - 1u)) * S + q.j) * 4;
            p[1] |= 1ULL << q.b1;
            p[2] |= 1ULL << q.b2;
            p[3] |= 1ULL << q.b3;
        }
        fast.rootkf = rkf;
        fast.rkf_sectors = S;
    }
    /* --- stored-k-mer filter (k_kf_insert), 6 bits per k-mer */
    const uint32_t kf_blocks = (uint32_t)(6.0 * (double)n / 256.0) + 1;
    uint64_t* kf = calloc((size_t)kf_blocks * 4 + 4, 8);
    for (size_t i = 0; i < n; i++) {
        const bft_kf_pos_t q = bft_kf_pos(st[i].km, W, a->k, kf_blocks);
        uint64_t* p = kf + (size_t)q.block * 4;
        p[0] |= 1ULL << q.b0; p[1] |= 1ULL << q.b1; p[2] |= 1ULL << q.b2; p[3] |= 1ULL << q.b3;
    }
    fast.kfilter = kf;
    fast.kf_blocks = kf_blocks;

    /* --- every stored k-mer through the accelerated view: found, with its own class (presence-only: found) */
    size_t stored_bad = 0;
    for (size_t i = 0; i < n; i++) {
        stored_bad += bft_lookup_w(&fast, st[i].km, W) != st[i].cls;
        stored_bad += bft_lookup_loc(&fast, st[i].km, W, BFT_LK_PRESENCE, NULL, NULL) == BFT_CLS_NONE;
        stored_bad += bft_lookup_loc(&fast, st[i].km, W, BFT_LK_NO_FILTER, NULL, NULL) != st[i].cls;
        stored_bad += bft_lookup_loc(&fast, st[i].km, W, BFT_LK_FILTER_FIRST, NULL, NULL) != st[i].cls;
    }
    /* --- the query file through both views */
    uint64_t* q; size_t nq;
    if (bft_read_kmer_file(argv[2], 1, a->k, a->W, &q, &nq)) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    size_t mismatch = 0, present = 0, rejected = 0;
    for (size_t i = 0; i < nq; i++) {
        const uint32_t want = bft_lookup_w(&plain, q + i * W, W);
        present += want != BFT_CLS_NONE;
        mismatch += bft_lookup_w(&fast, q + i * W, W) != want;
        mismatch += bft_lookup_loc(&fast, q + i * W, W, BFT_LK_NO_FILTER, NULL, NULL) != want;
        mismatch += bft_lookup_loc(&fast, q + i * W, W, BFT_LK_FILTER_FIRST, NULL, NULL) != want;
        mismatch += (bft_lookup_loc(&fast, q + i * W, W, BFT_LK_PRESENCE, NULL, NULL) != BFT_CLS_NONE) != (want != BFT_CLS_NONE);
        uint32_t stt[8] = {0};
        (void)bft_lookup_loc(&fast, q + i * W, W, 0, stt, NULL); /* statistics mode walks the structure and counts what the fast path does */
        rejected += stt[6];
        mismatch += (stt[2] != 0) != (want != BFT_CLS_NONE);
    }
    printf("stored=%zu deep_prefixes=%zu deep_buckets=%zu deep_kmers=%zu moved_by_probing=%zu rkf_sectors=%u stored_bad=%zu queries=%zu present=%zu "
           "filter_rejects=%zu mismatch=%zu\n", n, n_deep_prefixes, n_db, deep_kmers, probes_moved, S, stored_bad, nq, present, rejected, mismatch);
    free(q); free(kf); free(rkf); free(db); free(dcls); free(fastdir); free(cnt); free(st);
    bft_arena_free(a);
    return 0;
}
