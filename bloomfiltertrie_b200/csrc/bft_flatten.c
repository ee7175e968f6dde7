/* bft_flatten.c — serializer: .bft file -> flattened arena (layout in bft_arena.h).
 *
 * File format = what the reference writes in write_BFT_Root / write_Node / write_UC / write_CC
 * (reference src/write_to_disk.c:21-258) and reads back in read_BFT_Root_offset / read_Node / read_UC / read_CC
 * (:264-776). Bloom-filter bits, SkipFilter2 and SkipFilter3 are not in the file; the reference rebuilds them on
 * load (:578-581, 649-772). This serializer rebuilds the Bloom filters the same way (same hash, same prefixes) and
 * folds them into the per-Node first-CC table; rank/select helpers are replaced by exact prefix sums.
 * Only root->compressed == 0 files are accepted (the only kind the reference CLI produces, src/main.c:180,185).
 */
#define _GNU_SOURCE
#include "bft_flatten.h"
#include "bft_xxh64.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>

#define CEILDIV(a, b) (((a) + (b) - 1) / (b))

typedef struct {
    int nb_bits_skip2, nb_bits_skip3, nb_ucs_skp, nb_kmers_uc, level_min, modulo_hash, tresh_suf_pref;
} lvl_info_t;

typedef struct { uint64_t k[BFT_MAX_WORDS]; uint32_t cls; uint32_t idx; /* line index in the UC, i.e. the reference's stored order */ } line_tmp_t;

typedef struct {
    const uint8_t* buf;
    size_t len, pos;
    bft_arena_t* a;
    lvl_info_t lvl[16];
    uint64_t* hv1; /* raw XXH64 per idx14 */
    uint64_t* hv2;
    /* capacities */
    int64_t hp_off[16]; /* per level: offset of its bit-position table in hpos[], or -1 */
    size_t cap_hpos;
    size_t cap_nodes, cap_ccs, cap_firstcc, cap_csr, cap_filter3, cap_pref, cap_buckets, cap_ovf, cap_uc_lines, cap_cls_off, cap_cls_bytes;
    /* class hash map */
    /* colour-class dedup: open addressing; a slot carries the hash and (for annotations up to 16 bytes — nearly all
     * of them) the bytes themselves, so a look-up that hits costs one cache line instead of three dependent misses */
    struct cls_slot { uint64_t h; uint32_t id; uint32_t len; uint8_t inl[16]; }* map;
    size_t map_cap, map_used;
    struct cls_slot hot[4096]; /* recently seen classes (direct mapped): most lines of a UC share a few */
    /* scratch */
    line_tmp_t* tmp; size_t cap_tmp;
    uint8_t* scratch; size_t cap_scratch;   /* one annotation + extended byte */
    int16_t* extbuf; size_t cap_extbuf;     /* per line of the UC being read: extended byte or -1 */
    uint32_t* bkt; size_t cap_bkt;          /* bucketing scratch */
    void** live; size_t n_live, cap_live;   /* registered short-lived buffers */
    int depth;
    char* err; size_t errlen;
    jmp_buf jb;
} ctx_t;

static void fail(ctx_t* c, const char* fmt, ...) {
    if (c->err && c->errlen) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, c->errlen, fmt, ap);
        va_end(ap);
    }
    longjmp(c->jb, 1);
}

/* short-lived buffers of the recursive parse: registered so the failure path (longjmp) can release them */
static void* tmp_alloc(ctx_t* c, size_t n);
static void tmp_free(ctx_t* c, void* p);

static void* xrealloc(ctx_t* c, void* p, size_t n) {
    void* q = realloc(p, n ? n : 1);
    if (!q) fail(c, "bft_flatten: out of memory (%zu bytes)", n);
    return q;
}

static void* tmp_alloc(ctx_t* c, size_t n) {
    if (c->n_live == c->cap_live) {
        size_t ncap = c->cap_live ? c->cap_live * 2 : 64;
        void** t = (void**)realloc(c->live, ncap * sizeof(void*));
        if (!t) fail(c, "bft_flatten: out of memory");
        c->live = t;
        c->cap_live = ncap;
    }
    void* p = malloc(n ? n : 1);
    if (!p) fail(c, "bft_flatten: out of memory (%zu bytes)", n);
    c->live[c->n_live++] = p;
    return p;
}

static void tmp_free(ctx_t* c, void* p) {
    for (size_t i = c->n_live; i-- > 0;)
        if (c->live[i] == p) { c->live[i] = c->live[--c->n_live]; break; }
    free(p);
}

#define GROW(c, arr, cap, need, type)                                   \
    do {                                                                \
        if ((need) > (cap)) {                                           \
            size_t ncap_ = (cap) ? (cap) : 1024;                        \
            while (ncap_ < (need)) ncap_ += ncap_ / 2 + 1024;           \
            (arr) = (type*)xrealloc((c), (arr), ncap_ * sizeof(type));  \
            (cap) = ncap_;                                              \
        }                                                               \
    } while (0)

static const uint8_t* rd(ctx_t* c, size_t n) {
    if (n > c->len - c->pos) fail(c, "bft_flatten: truncated file (need %zu bytes at offset %zu of %zu)", n, c->pos, c->len);
    const uint8_t* p = c->buf + c->pos;
    c->pos += n;
    return p;
}
static uint16_t rd_u16(ctx_t* c) { uint16_t v; memcpy(&v, rd(c, 2), 2); return v; }
static uint32_t rd_u32(ctx_t* c) { uint32_t v; memcpy(&v, rd(c, 4), 4); return v; }
static int32_t rd_i32(ctx_t* c) { int32_t v; memcpy(&v, rd(c, 4), 4); return v; }
static int64_t rd_i64(ctx_t* c) { int64_t v; memcpy(&v, rd(c, 8), 8); return v; }

/* per-level geometry, reference create_info_per_level (src/CC.c:1883-1993) */
static int size_kmer_in_bytes(int sz) { return CEILDIV(sz * 2, 8); }
static int size_kmer_in_bytes_minus_1(int sz) { return sz > 9 ? CEILDIV((sz - 9) * 2, 8) : 0; }
static int exact_byte_level(int sz) { return sz == 45 || sz == 81 || sz == 117; }

/* ---- colour classes ------------------------------------------------------------------------------------ */
static uint32_t class_of(ctx_t* c, const uint8_t* s, size_t n) {
    bft_arena_t* a = c->a;
    if ((c->map_used + 1) * 2 > c->map_cap) {
        size_t ncap = c->map_cap ? c->map_cap * 2 : (1u << 16);
        struct cls_slot* nm = (struct cls_slot*)xrealloc(c, NULL, ncap * sizeof(struct cls_slot));
        memset(nm, 0xff, ncap * sizeof(struct cls_slot)); /* id 0xffffffff = empty */
        for (size_t i = 0; i < c->map_cap; i++) {
            if (c->map[i].id == 0xffffffffu) continue;
            size_t j = (size_t)c->map[i].h & (ncap - 1);
            while (nm[j].id != 0xffffffffu) j = (j + 1) & (ncap - 1);
            nm[j] = c->map[i];
        }
        free(c->map);
        c->map = nm;
        c->map_cap = ncap;
    }
    const uint64_t h = bft_xxh64(s, n, 0x5bd1e995);
    struct cls_slot* const hot = &c->hot[(size_t)(h >> 20) & 4095];
    if (hot->h == h && hot->len == (uint32_t)n && hot->id != 0xffffffffu &&
        (n <= sizeof hot->inl ? memcmp(hot->inl, s, n) == 0 : memcmp(a->cls_bytes + a->cls_off[hot->id], s, n) == 0))
        return hot->id;
    size_t j = (size_t)h & (c->map_cap - 1);
    for (;;) {
        const struct cls_slot* e = &c->map[j];
        if (e->id == 0xffffffffu) break;
        if (e->h == h && e->len == (uint32_t)n) {
            if (n <= sizeof e->inl ? memcmp(e->inl, s, n) == 0 : memcmp(a->cls_bytes + a->cls_off[e->id], s, n) == 0) {
                *hot = *e;
                return e->id;
            }
        }
        j = (j + 1) & (c->map_cap - 1);
    }
    if (a->n_classes >= 0xfffffff0u) fail(c, "bft_flatten: too many colour classes");
    uint32_t id = (uint32_t)a->n_classes;
    GROW(c, a->cls_off, c->cap_cls_off, a->n_classes + 2, uint32_t);
    GROW(c, a->cls_bytes, c->cap_cls_bytes, a->cls_bytes_len + n + 1, uint8_t);
    if (a->cls_bytes_len + n > 0xfffffff0u) fail(c, "bft_flatten: colour class bytes exceed 4 GiB");
    memcpy(a->cls_bytes + a->cls_bytes_len, s, n);
    a->cls_bytes_len += n;
    a->n_classes++;
    a->cls_off[a->n_classes] = (uint32_t)a->cls_bytes_len;
    if (n > a->max_cls_len) a->max_cls_len = n;
    c->map[j].h = h;
    c->map[j].id = id;
    c->map[j].len = (uint32_t)n;
    memset(c->map[j].inl, 0, sizeof c->map[j].inl);
    memcpy(c->map[j].inl, s, n < sizeof c->map[j].inl ? n : sizeof c->map[j].inl);
    c->map_used++;
    *hot = c->map[j];
    return id;
}

/* A UC body as stored in the file (reference write_UC, src/write_to_disk.c:119-207, uncompressed branch). */
typedef struct {
    const uint8_t* lines; /* n * (size_sub + size_annot) */
    const uint8_t* ext;   /* nb_ext * 3: (2-byte big-endian delta position, 1 byte) */
    int n, size_sub, size_annot, nb_ext;
    int16_t* ext_of_line; /* per line: extended byte or -1 (ctx scratch) */
} uc_body_t;

static void read_uc_body(ctx_t* c, uc_body_t* u, int size_sub, int n) {
    memset(u, 0, sizeof(*u));
    u->n = n;
    u->size_sub = size_sub;
    if (n == 0) return;
    u->nb_ext = rd_u16(c);
    u->size_annot = rd_i32(c);
    if (u->nb_ext == 0xffff) fail(c, "bft_flatten: compressed UC encountered (root->compressed files are not supported)");
    if (u->size_annot < 0 || u->size_annot > (1 << 24)) fail(c, "bft_flatten: implausible annotation size %d", u->size_annot);
    u->lines = rd(c, (size_t)n * (size_t)(size_sub + u->size_annot));
    u->ext = rd(c, (size_t)u->nb_ext * 3);
    GROW(c, c->extbuf, c->cap_extbuf, (size_t)n, int16_t);
    GROW(c, c->scratch, c->cap_scratch, (size_t)u->size_annot + 16, uint8_t);
    u->ext_of_line = c->extbuf; /* valid until the next read_uc_body: every UC is consumed before the next is read */
    for (int i = 0; i < n; i++) u->ext_of_line[i] = -1;
    /* get_extend_annot (src/UC.c:501-521): cumulative big-endian deltas, first entry reaching a position wins */
    int pos = 0;
    for (int i = 0; i < u->nb_ext; i++) {
        pos += (u->ext[i * 3] << 8) | u->ext[i * 3 + 1];
        if (pos < n && u->ext_of_line[pos] < 0) u->ext_of_line[pos] = u->ext[i * 3 + 2];
    }
}

/* class of line i: get_annot (src/UC.c:171-239) -> bytes handed to get_id_genomes_from_annot (src/bft.c:622-641) */
static uint32_t uc_line_class(ctx_t* c, const uc_body_t* u, int i) {
    uint8_t* scratch = c->scratch;
    if (u->size_annot == 0 || u->n == 0) return 0; /* class 0 = empty annotation */
    const uint8_t* annot = u->lines + (size_t)i * (size_t)(u->size_sub + u->size_annot) + u->size_sub;
    if (u->ext_of_line[i] >= 0) {
        memcpy(scratch, annot, (size_t)u->size_annot);
        scratch[u->size_annot] = (uint8_t)u->ext_of_line[i];
        return class_of(c, scratch, (size_t)u->size_annot + 1);
    }
    int n = u->size_annot; /* size_annot_sub: strip trailing zero bytes (include/annotation.h:183-191) */
    while (n > 0 && annot[n - 1] == 0) n--;
    return class_of(c, annot, (size_t)n);
}

static void line_key(ctx_t* c, const uc_body_t* u, int i, int key_bits, int strip_bit7, uint64_t* out) {
    uint8_t b[8 * BFT_MAX_WORDS];
    memset(b, 0, sizeof(b));
    if (u->size_sub > (int)sizeof(b)) fail(c, "bft_flatten: suffix of %d bytes exceeds the supported key width", u->size_sub);
    memcpy(b, u->lines + (size_t)i * (size_t)(u->size_sub + u->size_annot), (size_t)u->size_sub);
    if (strip_bit7 && u->size_sub > 0) b[u->size_sub - 1] &= 0x7f; /* in-band cluster flag, src/presenceNode.c:1901-1904 */
    for (int w = 0; w < BFT_MAX_WORDS; w++) memcpy(&out[w], b + 8 * w, 8);
    /* bits above key_bits must be zero: the reference compares whole bytes (memcmp) */
    for (int w = 0; w < BFT_MAX_WORDS; w++) {
        int lo = 64 * w;
        uint64_t allowed = key_bits >= lo + 64 ? ~0ULL : (key_bits <= lo ? 0ULL : ((1ULL << (key_bits - lo)) - 1));
        if (out[w] & ~allowed) fail(c, "bft_flatten: stray bits above %d in a stored suffix", key_bits);
    }
}

static int cmp_line_tmp(const void* pa, const void* pb) {
    const line_tmp_t* a = (const line_tmp_t*)pa;
    const line_tmp_t* b = (const line_tmp_t*)pb;
    for (int w = BFT_MAX_WORDS - 1; w >= 0; w--) {
        if (a->k[w] < b->k[w]) return -1;
        if (a->k[w] > b->k[w]) return 1;
    }
    return 0;
}

/* read lines [first, first+cnt) of a UC body into c->tmp (key + colour class), sorted by key */
static void load_block(ctx_t* c, const uc_body_t* u, int first, int cnt, int key_bits, int strip_bit7) {
    const int W = c->a->W;
    GROW(c, c->tmp, c->cap_tmp, (size_t)cnt, line_tmp_t);
    for (int i = 0; i < cnt; i++) {
        line_key(c, u, first + i, key_bits, strip_bit7, c->tmp[i].k);
        c->tmp[i].cls = uc_line_class(c, u, first + i);
        c->tmp[i].idx = (uint32_t)i;
        for (int w = W; w < BFT_MAX_WORDS; w++)
            if (c->tmp[i].k[w]) fail(c, "bft_flatten: key wider than W words");
    }
    if (cnt > 1) qsort(c->tmp, (size_t)cnt, sizeof(line_tmp_t), cmp_line_tmp);
    /* A well-formed BFT stores every k-mer once. The reference's own insertion breaks that for k > 63 (its
     * -extract_kmers then lists more records than distinct k-mers) and its answers depend on which copy a search
     * happens to meet; refuse such a file instead of answering differently. */
    for (int i = 1; i < cnt; i++)
        if (cmp_line_tmp(&c->tmp[i - 1], &c->tmp[i]) == 0)
            fail(c, "bft_flatten: this BFT stores the same k-mer twice (k=%d; the reference's insertion does that for k > 63): "
                    "its query results are not well defined, refusing to load it", c->a->k);
}

/* a Node's own UC -> uckeys/uccls (sorted, binary-searched); returns the index of its first line */
static uint32_t append_uc_lines(ctx_t* c, const uc_body_t* u, int cnt, int key_bits) {
    bft_arena_t* a = c->a;
    const int W = a->W;
    load_block(c, u, 0, cnt, key_bits, 0);
    if (a->n_uc_lines + (size_t)cnt >= 0xfffffff0u) fail(c, "bft_flatten: more than 2^32 Node-UC lines");
    const size_t need = a->n_uc_lines + (size_t)cnt;
    if (need > c->cap_uc_lines) {
        size_t ncap = c->cap_uc_lines ? c->cap_uc_lines : 4096;
        while (ncap < need) ncap += ncap / 2 + 4096;
        a->uckeys = (uint64_t*)xrealloc(c, a->uckeys, ncap * (size_t)W * sizeof(uint64_t));
        a->uccls = (uint32_t*)xrealloc(c, a->uccls, ncap * sizeof(uint32_t));
        a->uc_rank = (uint8_t*)xrealloc(c, a->uc_rank, ncap);
        c->cap_uc_lines = ncap;
    }
    const uint32_t start = (uint32_t)a->n_uc_lines;
    for (int i = 0; i < cnt; i++) {
        for (int w = 0; w < W; w++) a->uckeys[(a->n_uc_lines + (size_t)i) * W + w] = c->tmp[i].k[w];
        a->uccls[a->n_uc_lines + (size_t)i] = c->tmp[i].cls;
        a->uc_rank[a->n_uc_lines + (size_t)i] = (uint8_t)c->tmp[i].idx; /* a UC holds at most 255 lines */
    }
    a->n_uc_lines += (size_t)cnt;
    a->n_kmers += (size_t)cnt;
    return start;
}

/* one bucket of a block <- the lines idx[0..m): all of them if they fit, else BFT_BUCKET_KEYS - 1 and an overflow descriptor */
static void fill_bucket(ctx_t* c, uint64_t* bk, uint32_t* bc, size_t t, const uint32_t* idx, int m) {
    bft_arena_t* a = c->a;
    const int W = a->W, S = BFT_BUCKET_KEYS;
    const int in_bucket = m <= S ? m : S - 1;
    for (int q = 0; q < in_bucket; q++) {
        for (int w = 0; w < W; w++) bk[(t * S + q) * W + w] = c->tmp[idx[q]].k[w];
        bc[t * S + q] = c->tmp[idx[q]].cls;
    }
    if (m > S) { /* spill the rest; the last slot becomes the overflow descriptor */
        const int spill = m - (S - 1);
        if (a->n_ovf + (size_t)spill >= 0xfffffff0u) fail(c, "bft_flatten: overflow area exceeds 2^32 lines");
        if (a->n_ovf + (size_t)spill > c->cap_ovf) {
            size_t ncap = c->cap_ovf ? c->cap_ovf : 4096;
            while (ncap < a->n_ovf + (size_t)spill) ncap += ncap / 2 + 4096;
            a->ovf = (uint64_t*)xrealloc(c, a->ovf, ncap * (size_t)W * sizeof(uint64_t));
            a->ovfcls = (uint32_t*)xrealloc(c, a->ovfcls, ncap * sizeof(uint32_t));
            c->cap_ovf = ncap;
        }
        for (int q = 0; q < spill; q++) {
            for (int w = 0; w < W; w++) a->ovf[(a->n_ovf + (size_t)q) * W + w] = c->tmp[idx[S - 1 + q]].k[w];
            a->ovfcls[a->n_ovf + (size_t)q] = c->tmp[idx[S - 1 + q]].cls;
        }
        for (int w = 0; w < W - 1; w++) bk[(t * S + S - 1) * W + w] = 0;
        bk[(t * S + S - 1) * W + W - 1] = BFT_SLOT_SPECIAL | ((uint64_t)spill << 32) | (uint64_t)a->n_ovf;
        a->n_ovf += (size_t)spill;
    }
}

static void grow_buckets(ctx_t* c, size_t need) {
    bft_arena_t* a = c->a;
    const int W = a->W, S = BFT_BUCKET_KEYS;
    if (need >= 0xfffffff0u) fail(c, "bft_flatten: more than 2^32 suffix buckets");
    if (need > c->cap_buckets) {
        size_t ncap = c->cap_buckets ? c->cap_buckets : 4096;
        while (ncap < need) ncap += ncap / 2 + 4096;
        a->buckets = (uint64_t*)xrealloc(c, a->buckets, ncap * (size_t)(S * W) * sizeof(uint64_t));
        a->slotcls = (uint32_t*)xrealloc(c, a->slotcls, ncap * (size_t)S * sizeof(uint32_t));
        c->cap_buckets = ncap;
    }
}

/* a prefix's inline suffixes -> a block of B buckets of BFT_BUCKET_KEYS slots keyed by a hash of the suffix (layout in
 * bft_arena.h). Returns the INLINE entry (first bucket, B, count). */
static bft_entry_t append_inline_block(ctx_t* c, const uc_body_t* u, int first, int cnt, int key_bits, int strip_bit7) {
    bft_arena_t* a = c->a;
    const int W = a->W, S = BFT_BUCKET_KEYS;
    if (cnt > 255) fail(c, "bft_flatten: a prefix with %d inline suffixes (children_type counts are bytes: at most 255)", cnt);
    load_block(c, u, first, cnt, key_bits, strip_bit7);
    size_t B;
    if (BFT_PAIRED(W)) { /* load 0.6, an even number of 32-byte buckets so that they pair up in 64-byte lines */
        B = cnt <= S / 2 ? 1 : ((size_t)cnt * 5 + 11) / 12;
        if (B > 1 && (B & 1)) B++;
        if (B > 1 && (a->n_buckets & 1)) { /* blocks of pairs start on a 64-byte line: one unused bucket of padding */
            grow_buckets(c, a->n_buckets + 1);
            for (int i = 0; i < S * W; i++) a->buckets[a->n_buckets * (size_t)(S * W) + i] = BFT_SLOT_EMPTY;
            for (int i = 0; i < S; i++) a->slotcls[a->n_buckets * (size_t)S + i] = BFT_CLS_NONE;
            a->n_buckets++;
        }
    } else {
        uint32_t lb = 0;
        while (lb < BFT_MAX_LB && ((size_t)1 << lb) * (S / 2) < (size_t)cnt) lb++;
        B = (size_t)1 << lb;
    }
    grow_buckets(c, a->n_buckets + B);
    const uint32_t base = (uint32_t)a->n_buckets;
    uint64_t* bk = a->buckets + (size_t)base * (S * W);
    uint32_t* bc = a->slotcls + (size_t)base * S;
    for (size_t i = 0; i < B * (size_t)(S * W); i++) bk[i] = BFT_SLOT_EMPTY;
    for (size_t i = 0; i < B * (size_t)S; i++) bc[i] = BFT_CLS_NONE;
    /* bucket of every line, then a counting sort of the block by bucket */
    GROW(c, c->bkt, c->cap_bkt, (size_t)cnt * 2 + 2 * ((size_t)1 << BFT_MAX_LB) + 2, uint32_t);
    uint32_t* bkt = c->bkt;                 /* [cnt] bucket per line */
    uint32_t* order = c->bkt + cnt;         /* [cnt] lines grouped by bucket */
    uint32_t* head = c->bkt + 2 * (size_t)cnt; /* [B + 1] */
    for (size_t t = 0; t <= B; t++) head[t] = 0;
    for (int i = 0; i < cnt; i++) { bkt[i] = bft_bucket_idx(c->tmp[i].k, W, (uint32_t)B); head[bkt[i] + 1]++; }
    for (size_t t = 0; t < B; t++) head[t + 1] += head[t];
    uint32_t* fill = head + B + 1;          /* [B] running cursor */
    for (size_t t = 0; t < B; t++) fill[t] = head[t];
    for (int i = 0; i < cnt; i++) order[fill[bkt[i]]++] = (uint32_t)i;
    (void)key_bits;
    if (BFT_PAIRED(W) && B > 1) {
        for (size_t t = 0; t < B; t += 2) {
            const int m0 = (int)(head[t + 1] - head[t]), m1 = (int)(head[t + 2] - head[t + 1]);
            const uint32_t* i0 = order + head[t];
            const uint32_t* i1 = order + head[t + 1];
            if (m0 + m1 <= 2 * S) {
                /* the pair holds everything: each half takes up to S of its own lines, the rest sit in the other half's free slots */
                uint32_t half[2][BFT_BUCKET_KEYS];
                int n0 = 0, n1 = 0;
                for (int q = 0; q < m0 && q < S; q++) half[0][n0++] = i0[q];
                for (int q = 0; q < m1 && q < S; q++) half[1][n1++] = i1[q];
                for (int q = S; q < m0; q++) half[1][n1++] = i0[q];
                for (int q = S; q < m1; q++) half[0][n0++] = i1[q];
                fill_bucket(c, bk, bc, t, half[0], n0);
                fill_bucket(c, bk, bc, t + 1, half[1], n1);
            } else { /* too many for the line: each half on its own, with an overflow descriptor where it needs one */
                if (m0) fill_bucket(c, bk, bc, t, i0, m0);
                if (m1) fill_bucket(c, bk, bc, t + 1, i1, m1);
            }
        }
    } else {
        for (size_t t = 0; t < B; t++) {
            const int m = (int)(head[t + 1] - head[t]);
            if (m) fill_bucket(c, bk, bc, t, order + head[t], m);
        }
    }
    a->n_buckets += B;
    a->n_lines += (size_t)cnt;
    a->n_kmers += (size_t)cnt;
    return bft_mk_entry(BFT_KIND_INLINE, base, (uint32_t)cnt | ((uint32_t)B << BFT_NBK_SHIFT));
}

static uint32_t parse_node(ctx_t* c, int sz, int* cluster_flag);

static int get_nb_elts(const uint8_t* ct, int pos, int type8) { /* include/CC.h:358-366 */
    if (type8) return ct[pos];
    return (pos & 1) ? (ct[pos / 2] >> 4) : (ct[pos / 2] & 0xf);
}

/* parse one CC into descriptor slot cc_id; bf = this CC's regenerated Bloom filter (modulo_hash bits) */
static void parse_cc(ctx_t* c, int sz, uint32_t cc_id, uint8_t* bf) {
    bft_arena_t* a = c->a;
    const int lvl = sz / 9 - 1;
    const lvl_info_t* li = &c->lvl[lvl];
    const uint16_t type = rd_u16(c);
    const int nb_elem = rd_u16(c);
    const int nb_node_children = rd_u16(c);
    const int s = (type >> 1) & 0x1f;
    const int type8 = (type >> 6) & 1;
    if (s != 8 && s != 4) fail(c, "bft_flatten: CC with s=%d (expected 8 or 4)", s);
    const int p = BFT_PREFIX_BITS - s;
    const int n_pu = 1 << p;
    const int nb_skp = CEILDIV(nb_elem, li->nb_ucs_skp);

    const uint8_t* f2 = rd(c, (size_t)n_pu / 8);
    const size_t f3_bytes = s == 8 ? (size_t)nb_elem : (size_t)CEILDIV(nb_elem, 2);
    const uint8_t* f3 = rd(c, f3_bytes);
    const uint8_t* ef3 = li->level_min == 1 ? rd(c, (size_t)CEILDIV(nb_elem, 8)) : NULL;

    GROW(c, a->filter3, c->cap_filter3, a->filter3_bytes + f3_bytes + 8, uint8_t);
    if (a->filter3_bytes + f3_bytes > 0xfffffff0u) fail(c, "bft_flatten: filter3 arena exceeds 4 GiB");
    const uint32_t f3_off = (uint32_t)a->filter3_bytes;
    memcpy(a->filter3 + f3_off, f3, f3_bytes);
    a->filter3_bytes += f3_bytes;

    if (a->n_pref + (size_t)nb_elem > 0xfffffff0u) fail(c, "bft_flatten: more than 2^32 stored prefixes");
    const uint32_t pref_off = (uint32_t)a->n_pref;
    GROW(c, a->pref, c->cap_pref, a->n_pref + (size_t)nb_elem, bft_entry_t);
    a->n_pref += (size_t)nb_elem;

    uint8_t* flags = (uint8_t*)tmp_alloc(c, (size_t)nb_elem + 1);
    memset(flags, 0, (size_t)nb_elem + 1);
    if (ef3)
        for (int j = 0; j < nb_elem; j++) flags[j] = (ef3[j / 8] >> (j % 8)) & 1;

    int n_node_prefixes = 0;
    if (lvl > 0) {
        const uint8_t* ct = rd(c, type8 ? (size_t)nb_elem : (size_t)CEILDIV(nb_elem, 2));
        const int size_sub = size_kmer_in_bytes_minus_1(sz);
        const int key_bits = 2 * (sz - 9);
        const int strip = !exact_byte_level(sz);
        for (int b = 0; b < nb_skp; b++) {
            const int nlines = rd_u16(c);
            uc_body_t u;
            read_uc_body(c, &u, size_sub, nlines);
            const int j0 = b * li->nb_ucs_skp;
            const int j1 = j0 + li->nb_ucs_skp < nb_elem ? j0 + li->nb_ucs_skp : nb_elem;
            int off = 0;
            for (int j = j0; j < j1; j++) {
                const int ne = get_nb_elts(ct, j, type8);
                if (ne == 0) {
                    a->pref[pref_off + j] = bft_mk_entry(BFT_KIND_NODE, 0, 0);
                    n_node_prefixes++;
                    continue;
                }
                if (off + ne > nlines) fail(c, "bft_flatten: children_type overruns its UC bucket");
                if (li->level_min == 0) /* in-band cluster flag: bit 7 of the last suffix byte of the first line */
                    flags[j] = u.lines[(size_t)off * (size_t)(size_sub + u.size_annot) + size_sub - 1] >> 7;
                a->pref[pref_off + j] = append_inline_block(c, &u, off, ne, key_bits, strip);
                off += ne;
            }
            if (off != nlines) fail(c, "bft_flatten: UC bucket holds %d lines but children_type accounts for %d", nlines, off);
        }
        /* child Nodes follow, in prefix order (write_CC, src/write_to_disk.c:254-255) */
        if (n_node_prefixes != nb_node_children)
            fail(c, "bft_flatten: CC declares %d child nodes but children_type has %d", nb_node_children, n_node_prefixes);
        for (int j = 0; j < nb_elem; j++) {
            if ((a->pref[pref_off + j].b >> BFT_KIND_SHIFT) != BFT_KIND_NODE) continue;
            int flag = 0;
            uint32_t child = parse_node(c, sz - 9, &flag);
            a->pref[pref_off + j].a = child;
            if (li->level_min == 0) flags[j] = (uint8_t)flag; /* UC_array.nb_children & 1, src/presenceNode.c:1732 */
        }
    } else {
        /* leaf level: zero-length suffixes, one annotation per prefix (src/presenceNode.c:1453-1463) */
        if (nb_node_children) fail(c, "bft_flatten: leaf CC with child nodes");
        for (int b = 0; b < nb_skp; b++) {
            const int j0 = b * li->nb_ucs_skp;
            const int j1 = j0 + li->nb_ucs_skp < nb_elem ? j0 + li->nb_ucs_skp : nb_elem;
            uc_body_t u;
            read_uc_body(c, &u, 0, j1 - j0);
            for (int j = j0; j < j1; j++)
                a->pref[pref_off + j] = bft_mk_entry(BFT_KIND_LEAF, uc_line_class(c, &u, j - j0), 1);
        }
        a->n_kmers += (size_t)nb_elem;
        a->n_leaf_prefixes += (size_t)nb_elem;
    }

    /* cluster directory: rank over filter2 + select over the cluster-start flags (findCluster,
     * src/presenceNode.c:1578-1821) folded into exclusive prefix sums */
    int* starts = (int*)tmp_alloc(c, ((size_t)nb_elem + 2) * sizeof(int));
    int n_starts = 0;
    for (int j = 0; j < nb_elem; j++)
        if (flags[j]) starts[n_starts++] = j;
    starts[n_starts] = nb_elem;
    GROW(c, a->csr, c->cap_csr, a->n_csr + (size_t)n_pu + 1, uint16_t);
    if (a->n_csr + (size_t)n_pu + 1 > 0xfffffff0u) fail(c, "bft_flatten: cluster directory exceeds 2^32 entries");
    const uint32_t csr_off = (uint32_t)a->n_csr;
    uint16_t* csr = a->csr + csr_off;
    int rank = 0;
    for (int pu = 0; pu < n_pu; pu++) {
        csr[pu] = (uint16_t)(rank < n_starts ? starts[rank] : nb_elem); /* a p_u without flag: pos == nb_elem (:1394) */
        if ((f2[pu / 8] >> (pu % 8)) & 1) rank++;
    }
    csr[n_pu] = (uint16_t)(rank < n_starts ? starts[rank] : nb_elem);
    if (rank != n_starts) fail(c, "bft_flatten: filter2 has %d p_u but %d cluster starts (level of %d nt, %d prefixes, s=%d, level_min=%d, %d child nodes, file offset %zu)", rank, n_starts, sz, nb_elem, s, li->level_min, nb_node_children, c->pos);
    if (nb_elem && starts[0] != 0) fail(c, "bft_flatten: first stored prefix does not start a cluster");
    a->n_csr += (size_t)n_pu + 1;

    /* regenerate this CC's Bloom filter from its stored prefixes (read_CC, src/write_to_disk.c:656-683, 696-772) */
    const int nbf = CEILDIV(li->modulo_hash, 8);
    memset(bf, 0, (size_t)nbf);
    for (int pu = 0; pu < n_pu; pu++) {
        for (int j = csr[pu]; j < csr[pu + 1]; j++) {
            uint32_t pv = s == 8 ? f3[j] : ((j & 1) ? (uint32_t)(f3[j / 2] >> 4) : (uint32_t)(f3[j / 2] & 0xf));
            uint32_t rot = ((uint32_t)pu << s) | pv;
            uint32_t idx = (rot >> 4) & 0x3fffu;
            uint32_t h1 = (uint32_t)(c->hv1[idx] % (uint64_t)li->modulo_hash);
            uint32_t h2 = (uint32_t)(c->hv2[idx] % (uint64_t)li->modulo_hash);
            bf[h1 / 8] |= (uint8_t)(1u << (h1 % 8));
            bf[h2 / 8] |= (uint8_t)(1u << (h2 % 8));
        }
    }

    bft_cc_t* cc = &a->ccs[cc_id];
    cc->csr_off = csr_off;
    cc->f3_off = f3_off;
    cc->pref_off = pref_off;
    cc->nb_elem = (uint16_t)nb_elem;
    cc->s = (uint8_t)s;
    cc->pad = 0;
    tmp_free(c, starts);
    tmp_free(c, flags);
}

static uint32_t parse_node(ctx_t* c, int sz, int* cluster_flag) {
    bft_arena_t* a = c->a;
    if (sz < 9) fail(c, "bft_flatten: trie deeper than k/9 levels");
    const int lvl = sz / 9 - 1;
    const lvl_info_t* li = &c->lvl[lvl];
    c->depth++;
    if (c->depth > a->max_depth) a->max_depth = c->depth;

    if (a->n_nodes >= 0xfffffff0u) fail(c, "bft_flatten: more than 2^32 nodes");
    const uint32_t id = (uint32_t)a->n_nodes;
    GROW(c, a->nodes, c->cap_nodes, a->n_nodes + 1, bft_node_t);
    a->n_nodes++;
    memset(&a->nodes[id], 0, sizeof(bft_node_t));

    /* the Node's own UC (write_Node, src/write_to_disk.c:102-103): nb_children = count << 1 | cluster flag */
    const uint16_t raw = rd_u16(c);
    if (cluster_flag) *cluster_flag = raw & 1;
    const int n_uc = raw >> 1;
    uc_body_t u;
    read_uc_body(c, &u, size_kmer_in_bytes(sz), n_uc);
    uint32_t uc_begin = 0;
    if (n_uc) {
        uc_begin = append_uc_lines(c, &u, n_uc, 2 * sz);
    }
    const uint32_t n_cc = rd_u32(c);
    if (n_cc >= BFT_FIRSTCC_NONE) fail(c, "bft_flatten: node with %u CCs (first-CC table holds at most 254)", n_cc);
    if (a->n_ccs + n_cc > 0xfffffff0u) fail(c, "bft_flatten: more than 2^32 CCs");
    const uint32_t cc_begin = (uint32_t)a->n_ccs;
    GROW(c, a->ccs, c->cap_ccs, a->n_ccs + n_cc, bft_cc_t);
    a->n_ccs += n_cc;
    if ((int)n_cc > a->max_cc_per_node) a->max_cc_per_node = (int)n_cc;

    uint32_t fc_off = 0, bf_mode = 0, bf_stride = 0;
    if (n_cc) {
        const int nbf = CEILDIV(li->modulo_hash, 8);
        uint8_t* bfs = (uint8_t*)tmp_alloc(c, (size_t)n_cc * (size_t)nbf);
        for (uint32_t i = 0; i < n_cc; i++) parse_cc(c, sz, cc_begin + i, bfs + (size_t)i * nbf);
        if (n_cc <= BFT_BF_DIRECT_MAX) {
            /* few CCs: keep the Bloom filters themselves (bf_mode 1, see bft_arena.h) */
            if (c->hp_off[lvl] < 0) { /* this level's bit positions per idx14 (src/presenceNode.c:1341-1350) */
                GROW(c, a->hpos, c->cap_hpos, a->n_hpos + BFT_N_IDX14, uint32_t);
                for (uint32_t idx = 0; idx < BFT_N_IDX14; idx++) {
                    const uint32_t h1 = (uint32_t)(c->hv1[idx] % (uint64_t)li->modulo_hash);
                    const uint32_t h2 = (uint32_t)(c->hv2[idx] % (uint64_t)li->modulo_hash);
                    a->hpos[a->n_hpos + idx] = h1 | (h2 << 16);
                }
                c->hp_off[lvl] = (int64_t)a->n_hpos;
                a->n_hpos += BFT_N_IDX14;
            }
            const size_t stride = ((size_t)nbf + 3) & ~(size_t)3;
            if (a->firstcc_bytes + n_cc * stride > 0xfffffff0u) fail(c, "bft_flatten: Bloom filters exceed 4 GiB");
            GROW(c, a->firstcc, c->cap_firstcc, a->firstcc_bytes + n_cc * stride, uint8_t);
            fc_off = (uint32_t)a->firstcc_bytes;
            memset(a->firstcc + fc_off, 0, n_cc * stride);
            for (uint32_t i = 0; i < n_cc; i++) memcpy(a->firstcc + fc_off + i * stride, bfs + (size_t)i * nbf, (size_t)nbf);
            a->firstcc_bytes += n_cc * stride;
            bf_mode = 1;
            bf_stride = (uint32_t)stride;
        } else {
        /* first CC whose Bloom filter fires, per 14-bit hash index (src/presenceNode.c:1354-1362) */
        if (a->firstcc_bytes + BFT_N_IDX14 > 0xfffffff0u) fail(c, "bft_flatten: first-CC tables exceed 4 GiB");
        GROW(c, a->firstcc, c->cap_firstcc, a->firstcc_bytes + BFT_N_IDX14, uint8_t);
        fc_off = (uint32_t)a->firstcc_bytes;
        uint8_t* fc = a->firstcc + fc_off;
        for (uint32_t idx = 0; idx < BFT_N_IDX14; idx++) {
            uint32_t h1 = (uint32_t)(c->hv1[idx] % (uint64_t)li->modulo_hash);
            uint32_t h2 = (uint32_t)(c->hv2[idx] % (uint64_t)li->modulo_hash);
            uint8_t hit = BFT_FIRSTCC_NONE;
            for (uint32_t i = 0; i < n_cc; i++) {
                const uint8_t* bf = bfs + (size_t)i * nbf;
                if ((bf[h1 / 8] >> (h1 % 8)) & (bf[h2 / 8] >> (h2 % 8)) & 1) { hit = (uint8_t)i; break; }
            }
            fc[idx] = hit;
        }
        a->firstcc_bytes += BFT_N_IDX14;
        }
        tmp_free(c, bfs);
    }
    bft_node_t* nd = &a->nodes[id];
    nd->cc_begin = cc_begin;
    nd->n_cc = n_cc;
    nd->fc_off = fc_off;
    nd->uc_begin = uc_begin;
    nd->uc_n = (uint32_t)n_uc;
    nd->bf_mode = bf_mode;
    nd->hp_off = bf_mode ? (uint32_t)c->hp_off[lvl] : 0;
    nd->bf_stride = bf_stride;
    c->depth--;
    return id;
}

void bft_arena_view(const bft_arena_t* a, bft_view_t* v) {
    memset(v, 0, sizeof *v); /* no stored-k-mer filter on the host: it is built on the device */
    v->rootdir = a->rootdir;
    v->nodes = a->nodes;
    v->ccs = a->ccs;
    v->firstcc = a->firstcc;
    v->hpos = a->hpos;
    v->csr = a->csr;
    v->filter3 = a->filter3;
    v->pref = a->pref;
    v->buckets = a->buckets;
    v->ovf = a->ovf;
    v->slotcls = a->slotcls;
    v->ovfcls = a->ovfcls;
    v->uckeys = a->uckeys;
    v->uccls = a->uccls;
    v->uc_rank = a->uc_rank;
    v->pref_low18 = a->pref_low18;
    v->pref_node = a->pref_node;
    v->node_path = a->node_path;
    v->pref_out = a->pref_out;
    v->cls_shift = a->cls_shift;
    v->cls_mask = a->cls_mask;
    v->k = a->k;
    v->W = a->W;
    v->loc_ovf = (uint32_t)(a->n_buckets * BFT_BUCKET_KEYS);
    v->loc_uc = v->loc_ovf + (uint32_t)a->n_ovf;
    v->loc_leaf = v->loc_uc + (uint32_t)a->n_uc_lines;
}

void bft_arena_free(bft_arena_t* a) {
    if (!a) return;
    if (a->filenames) {
        for (int i = 0; i < a->n_genomes; i++) free(a->filenames[i]);
        free(a->filenames);
    }
    free(a->rootdir); free(a->nodes); free(a->ccs); free(a->firstcc); free(a->hpos); free(a->csr); free(a->filter3);
    free(a->pref_low18); free(a->pref_node); free(a->node_path); free(a->pref_out);
    free(a->pref); free(a->buckets); free(a->slotcls); free(a->ovf); free(a->ovfcls); free(a->uckeys); free(a->uccls); free(a->uc_rank); free(a->cls_off); free(a->cls_bytes);
    free(a->pool_last_index); free(a->pool_size_annot); free(a->pool_off); free(a->pool_bytes);
    free(a);
}

size_t bft_arena_bytes(const bft_arena_t* a) {
    return BFT_ROOTDIR_SIZE * sizeof(bft_entry_t) + a->n_nodes * sizeof(bft_node_t) + a->n_ccs * sizeof(bft_cc_t) +
           a->firstcc_bytes + a->n_hpos * 4 + a->n_csr * 2 + a->filter3_bytes + a->n_pref * sizeof(bft_entry_t) +
           (a->n_buckets * BFT_BUCKET_KEYS + a->n_ovf) * ((size_t)a->W * 8 + (a->cls_shift ? 0 : 4)) +
           a->n_uc_lines * ((size_t)a->W * 8 + 4) + (a->n_classes + 1) * 4 + a->cls_bytes_len + a->pool_bytes_len +
           a->n_pref * 16 + a->n_nodes * sizeof(bft_path_t);
}

bft_arena_t* bft_arena_from_memory(const uint8_t* buf, size_t len, char* err, size_t errlen) {
    ctx_t* c = (ctx_t*)calloc(1, sizeof(ctx_t));
    if (c) memset(c->hot, 0xff, sizeof c->hot); /* id 0xffffffff = empty */
    bft_arena_t* a = (bft_arena_t*)calloc(1, sizeof(bft_arena_t));
    if (!c || !a) { free(c); free(a); if (err) snprintf(err, errlen, "bft_flatten: out of memory"); return NULL; }
    c->buf = buf; c->len = len; c->a = a; c->err = err; c->errlen = errlen;
    for (int i = 0; i < 16; i++) c->hp_off[i] = -1;
    if (err && errlen) err[0] = 0;
    if (setjmp(c->jb)) {
        for (size_t i = 0; i < c->n_live; i++) free(c->live[i]);
        free(c->live);
        free(c->map); free(c->tmp); free(c->hv1); free(c->hv2); free(c->scratch); free(c->extbuf); free(c->bkt);
        free(c);
        bft_arena_free(a);
        return NULL;
    }
    /* header: comp_set_colors pools (write_BFT_Root, src/write_to_disk.c:34-61) */
    a->n_pools = rd_i32(c);
    if (a->n_pools < 0 || a->n_pools > (1 << 20)) fail(c, "bft_flatten: not a .bft file (pool count %d)", a->n_pools);
    a->pool_last_index = (int64_t*)xrealloc(c, NULL, (size_t)(a->n_pools + 1) * sizeof(int64_t));
    a->pool_size_annot = (int32_t*)xrealloc(c, NULL, (size_t)(a->n_pools + 1) * sizeof(int32_t));
    a->pool_off = (uint64_t*)xrealloc(c, NULL, (size_t)(a->n_pools + 1) * sizeof(uint64_t));
    size_t cap_pool = 0;
    for (int i = 0; i < a->n_pools; i++) {
        a->pool_last_index[i] = rd_i64(c);
        a->pool_size_annot[i] = rd_i32(c);
        int64_t cnt = i ? a->pool_last_index[i] - a->pool_last_index[i - 1] : a->pool_last_index[i] + 1;
        if (cnt < 0 || a->pool_size_annot[i] < 0) fail(c, "bft_flatten: corrupt colour pool %d", i);
        if ((uint64_t)cnt > c->len || (uint64_t)a->pool_size_annot[i] > c->len ||
            (uint64_t)cnt * (uint64_t)a->pool_size_annot[i] > c->len - c->pos)
            fail(c, "bft_flatten: colour pool %d is larger than the file", i);
        size_t nbytes = (size_t)cnt * (size_t)a->pool_size_annot[i];
        a->pool_off[i] = a->pool_bytes_len;
        GROW(c, a->pool_bytes, cap_pool, a->pool_bytes_len + nbytes + 1, uint8_t);
        memcpy(a->pool_bytes + a->pool_bytes_len, rd(c, nbytes), nbytes);
        a->pool_bytes_len += nbytes;
    }
    a->r1 = rd_i32(c);
    a->r2 = rd_i32(c);
    a->treshold_compression = rd_i32(c);
    a->n_genomes = rd_i32(c);
    a->k = rd_i32(c);
    a->compressed = *rd(c, 1);
    if (a->k <= 0 || a->k % 9 != 0 || a->k > 126) fail(c, "bft_flatten: not a .bft file (k=%d)", a->k);
    if (a->compressed) fail(c, "bft_flatten: root->compressed=%d files are not supported", a->compressed);
    if (a->n_genomes < 0 || a->n_genomes > 100000000 || (size_t)a->n_genomes * 2 > c->len - c->pos)
        fail(c, "bft_flatten: implausible genome count %d", a->n_genomes);
    a->W = a->k <= 27 ? 1 : (a->k <= 63 ? 2 : 4);
    a->n_levels = a->k / 9;
    a->filenames = (char**)xrealloc(c, NULL, (size_t)(a->n_genomes + 1) * sizeof(char*));
    memset(a->filenames, 0, (size_t)(a->n_genomes + 1) * sizeof(char*));
    for (int i = 0; i < a->n_genomes; i++) {
        uint16_t sl = rd_u16(c);
        const uint8_t* s = rd(c, sl);
        a->filenames[i] = (char*)xrealloc(c, NULL, (size_t)sl + 1);
        memcpy(a->filenames[i], s, sl);
        a->filenames[i][sl] = 0;
    }
    for (int i = 0; i < a->n_levels; i++) { /* src/write_to_disk.c:78-86 */
        c->lvl[i].nb_bits_skip2 = rd_i32(c);
        c->lvl[i].nb_bits_skip3 = rd_i32(c);
        c->lvl[i].nb_ucs_skp = rd_i32(c);
        c->lvl[i].nb_kmers_uc = rd_i32(c);
        c->lvl[i].level_min = rd_i32(c);
        c->lvl[i].modulo_hash = rd_i32(c);
        c->lvl[i].tresh_suf_pref = rd_i32(c);
        if (c->lvl[i].nb_ucs_skp <= 0 || c->lvl[i].modulo_hash <= 0 || c->lvl[i].modulo_hash > 65536)
            fail(c, "bft_flatten: corrupt level table");
    }
    /* hash_v restricted to the 16384 reachable entries (create_hash_v_array, include/Node.h:158-185;
     * use at src/presenceNode.c:1341-1343): bytes of i MSB-first over 18 bits */
    c->hv1 = (uint64_t*)xrealloc(c, NULL, BFT_N_IDX14 * sizeof(uint64_t));
    c->hv2 = (uint64_t*)xrealloc(c, NULL, BFT_N_IDX14 * sizeof(uint64_t));
    for (uint32_t i = 0; i < BFT_N_IDX14; i++) {
        uint8_t g[3] = {(uint8_t)((i >> 10) & 0xff), (uint8_t)((i >> 2) & 0xff), (uint8_t)((i << 6) & 0xff)};
        c->hv1[i] = bft_xxh64(g, 3, (uint64_t)(int64_t)a->r1);
        c->hv2[i] = bft_xxh64(g, 3, (uint64_t)(int64_t)a->r2);
    }
    /* class 0 = empty annotation */
    GROW(c, a->cls_off, c->cap_cls_off, 2, uint32_t);
    a->cls_off[0] = 0;
    uint8_t dummy = 0;
    class_of(c, &dummy, 0);

    if (c->pos < c->len) {
        parse_node(c, a->k, NULL);
    } else { /* write_root_only file: empty trie */
        GROW(c, a->nodes, c->cap_nodes, 1, bft_node_t);
        memset(&a->nodes[0], 0, sizeof(bft_node_t));
        a->n_nodes = 1;
    }
    if (c->pos != c->len) fail(c, "bft_flatten: %zu trailing bytes after the root node", c->len - c->pos);

    /* pad arrays the device reads with vector loads */
    GROW(c, a->filter3, c->cap_filter3, a->filter3_bytes + 16, uint8_t);
    memset(a->filter3 + a->filter3_bytes, 0, 16);
    GROW(c, a->csr, c->cap_csr, a->n_csr + 8, uint16_t);
    GROW(c, a->pref, c->cap_pref, a->n_pref + 1, bft_entry_t);
    GROW(c, a->ccs, c->cap_ccs, a->n_ccs + 1, bft_cc_t);
    GROW(c, a->firstcc, c->cap_firstcc, a->firstcc_bytes + 1, uint8_t);
    GROW(c, a->hpos, c->cap_hpos, a->n_hpos + 1, uint32_t);
    if (!a->buckets) {
        a->buckets = (uint64_t*)xrealloc(c, NULL, 8 * BFT_MAX_WORDS * BFT_BUCKET_KEYS);
        a->slotcls = (uint32_t*)xrealloc(c, NULL, 4 * BFT_BUCKET_KEYS);
    }
    if (!a->ovf) {
        a->ovf = (uint64_t*)xrealloc(c, NULL, 8 * BFT_MAX_WORDS);
        a->ovfcls = (uint32_t*)xrealloc(c, NULL, 8);
    }
    if (!a->uckeys) {
        a->uckeys = (uint64_t*)xrealloc(c, NULL, 8 * BFT_MAX_WORDS);
        a->uccls = (uint32_t*)xrealloc(c, NULL, 8);
        a->uc_rank = (uint8_t*)xrealloc(c, NULL, 8);
    }

    /* colour class ids into the spare top bits of the inline keys when they fit between the widest suffix and the
     * flag bit 63: k <= 27: 36 suffix bits + up to 27 class bits; k <= 63: 44 bits in the upper word + up to 19;
     * k <= 126: 42 bits in the top word + up to 21 */
    {
        int cls_bits = 1;
        while (((size_t)1 << cls_bits) < a->n_classes) cls_bits++;
        int top_bits = 0; /* suffix bits in the most significant key word, over all levels with inline blocks */
        for (int sz = a->k; sz >= 18; sz -= 9) {
            const int kb = 2 * (sz - 9) - 64 * (a->W - 1);
            if (kb > top_bits) top_bits = kb;
        }
        a->cls_shift = 0;
        a->cls_mask = 0;
        const char* no_embed = getenv("BFT_B200_NO_EMBED"); /* test hook: keep the separate class arrays */
        if (top_bits + cls_bits <= 63 && cls_bits < 32 && !(no_embed && no_embed[0] == '1')) {
            a->cls_shift = 63 - cls_bits;
            a->cls_mask = (uint32_t)(((uint64_t)1 << cls_bits) - 1);
            const size_t W = (size_t)a->W;
            for (size_t i = 0; i < a->n_buckets * BFT_BUCKET_KEYS; i++) {
                uint64_t* top = &a->buckets[i * W + W - 1];
                if (!(*top & BFT_SLOT_SPECIAL)) *top |= (uint64_t)a->slotcls[i] << a->cls_shift;
            }
            for (size_t i = 0; i < a->n_ovf; i++) a->ovf[i * W + W - 1] |= (uint64_t)a->ovfcls[i] << a->cls_shift;
            free(a->slotcls);
            free(a->ovfcls);
            a->slotcls = NULL;
            a->ovfcls = NULL;
        }
    }

    /* enumeration side tables: Nodes are numbered parent-before-child, so one forward sweep fixes every Node's
     * path. Enumeration order = stored prefixes in arena order (each followed by its inline suffixes in bucket
     * order), then the Nodes' own UC lines. */
    a->pref_low18 = (uint32_t*)xrealloc(c, NULL, (a->n_pref + 1) * sizeof(uint32_t));
    a->pref_node = (uint32_t*)xrealloc(c, NULL, (a->n_pref + 1) * sizeof(uint32_t));
    a->pref_out = (uint64_t*)xrealloc(c, NULL, (a->n_pref + 1) * sizeof(uint64_t));
    a->node_path = (bft_path_t*)xrealloc(c, NULL, (a->n_nodes + 1) * sizeof(bft_path_t));
    memset(a->node_path, 0, (a->n_nodes + 1) * sizeof(bft_path_t));
    for (size_t nid = 0; nid < a->n_nodes; nid++) {
        const bft_node_t* nd = &a->nodes[nid];
        const bft_path_t pp = a->node_path[nid];
        for (uint32_t ci = 0; ci < nd->n_cc; ci++) {
            const bft_cc_t* cc = &a->ccs[nd->cc_begin + ci];
            const uint16_t* csr = a->csr + cc->csr_off;
            const uint8_t* f3 = a->filter3 + cc->f3_off;
            const int n_pu = 1 << (BFT_PREFIX_BITS - cc->s);
            for (int pu = 0; pu < n_pu; pu++) {
                for (uint32_t j = csr[pu]; j < csr[pu + 1]; j++) {
                    const uint32_t pv = cc->s == 8 ? f3[j] : ((j & 1) ? (uint32_t)(f3[j / 2] >> 4) : (uint32_t)(f3[j / 2] & 0xf));
                    const uint32_t rot = ((uint32_t)pu << cc->s) | pv;              /* nuc1..nuc8,nuc0 */
                    const uint32_t r18 = (rot >> 2) | ((rot & 3u) << 16);           /* nuc0..nuc8, MSB first */
                    uint32_t low18 = 0;
                    for (int q = 0; q < 9; q++) low18 |= ((r18 >> (2 * (8 - q))) & 3u) << (2 * q);
                    const size_t pj = (size_t)cc->pref_off + j;
                    a->pref_low18[pj] = low18;
                    a->pref_node[pj] = (uint32_t)nid;
                    if ((a->pref[pj].b >> BFT_KIND_SHIFT) == BFT_KIND_NODE) {
                        bft_path_t* cp = &a->node_path[a->pref[pj].a];
                        *cp = pp;
                        const unsigned sh = BFT_PREFIX_BITS * pp.depth;
                        cp->acc[sh >> 6] |= (uint64_t)low18 << (sh & 63);
                        if ((sh & 63) > 64 - BFT_PREFIX_BITS && (sh >> 6) + 1 < BFT_MAX_WORDS) cp->acc[(sh >> 6) + 1] |= (uint64_t)low18 >> (64 - (sh & 63));
                        cp->depth = pp.depth + 1;
                    }
                }
            }
        }
    }
    {
        /* Enumeration order = the reference's iterate_over_kmers_from_node (src/extract_kmers.c:3-597): depth first —
         * the CCs of a Node in order, the stored prefixes of a CC in order, under a prefix either its inline suffix
         * lines (in the order the UC stores them) or the whole subtree of its child Node, and after the CCs the
         * Node's own UC lines. pref_out[j] / node_path[].uc_out = index of the first k-mer of prefix j / of the
         * Node's UC in that order; an explicit stack replaces the recursion (depth <= k/9 <= 14). */
        uint64_t run = 0;
        struct { uint32_t node, cc, j; } stk[16];
        int sp = 0;
        stk[0].node = 0; stk[0].cc = 0; stk[0].j = 0;
        memset(a->pref_out, 0, (a->n_pref + 1) * sizeof(uint64_t));
        while (sp >= 0) {
            const bft_node_t* nd = &a->nodes[stk[sp].node];
            int descended = 0;
            while (stk[sp].cc < nd->n_cc && !descended) {
                const bft_cc_t* cc = &a->ccs[nd->cc_begin + stk[sp].cc];
                while (stk[sp].j < cc->nb_elem) {
                    const size_t pj = (size_t)cc->pref_off + stk[sp].j++;
                    const uint32_t kind = a->pref[pj].b >> BFT_KIND_SHIFT;
                    a->pref_out[pj] = run;
                    if (kind == BFT_KIND_INLINE) run += BFT_INLINE_CNT(a->pref[pj]);
                    else if (kind == BFT_KIND_LEAF) run += 1;
                    else if (kind == BFT_KIND_NODE) {
                        if (sp + 1 >= 16) fail(c, "bft_flatten: trie deeper than 16 levels");
                        sp++;
                        stk[sp].node = a->pref[pj].a; stk[sp].cc = 0; stk[sp].j = 0;
                        descended = 1;
                        break;
                    }
                }
                if (!descended) { stk[sp].cc++; stk[sp].j = 0; }
            }
            if (descended) continue;
            a->node_path[stk[sp].node].uc_out_lo = (uint32_t)run;
            a->node_path[stk[sp].node].uc_out_hi = (uint32_t)(run >> 32);
            run += nd->uc_n;
            sp--;
        }
        a->pref_out[a->n_pref] = run;
        if (run != a->n_kmers) fail(c, "bft_flatten: enumeration order counts %llu k-mers, the arena %zu", (unsigned long long)run, a->n_kmers);
    }

    /* root directory: the root probe for every 9-nt prefix */
    a->rootdir = (bft_entry_t*)xrealloc(c, NULL, BFT_ROOTDIR_SIZE * sizeof(bft_entry_t));
    bft_view_t v;
    bft_arena_view(a, &v);
    for (uint32_t low18 = 0; low18 < BFT_ROOTDIR_SIZE; low18++) a->rootdir[low18] = bft_node_probe(&v, 0, low18, 0);

    free(c->live);
    free(c->map); free(c->tmp); free(c->hv1); free(c->hv2); free(c->scratch); free(c->extbuf); free(c->bkt);
    free(c);
    return a;
}

bft_arena_t* bft_arena_from_file(const char* path, char* err, size_t errlen) {
    FILE* f = fopen(path, "rb");
    if (!f) { if (err) snprintf(err, errlen, "bft_flatten: cannot open %s", path); return NULL; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* buf = (uint8_t*)malloc(n > 0 ? (size_t)n : 1);
    if (!buf || (n > 0 && fread(buf, 1, (size_t)n, f) != (size_t)n)) {
        if (err) snprintf(err, errlen, "bft_flatten: cannot read %s", path);
        free(buf); fclose(f);
        return NULL;
    }
    fclose(f);
    bft_arena_t* a = bft_arena_from_memory(buf, (size_t)n, err, errlen);
    free(buf);
    return a;
}
