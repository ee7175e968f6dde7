/* bft_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE. See bft_oracle.h.
 *
 * Restates, function by function, the reference's query path on the reference's own data layout. Citations are
 * paths in the GuillaumeHolley/BloomFilterTrie tree. Deliberate simplifications (none changes a result): the
 * SkipFilter2/SkipFilter3 acceleration arrays are not rebuilt — rank and select are plain scans — and
 * children_type prefix sums are computed once at load instead of per query (count_children / count_nodes).
 */
#define _GNU_SOURCE
#include "bft_oracle.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NB_CHAR_SUF_PREF 9            /* include/default_param.h:12 */
#define SIZE_BYTE_EXT_ANNOT 3         /* include/default_param.h:6 */
#define CEIL(a, b) (((a) / (b)) + (((a) % (b)) > 0 ? 1 : 0))

/* ---- XXH64 for short inputs, from the published xxHash64 specification (the reference vendors v0.6.2 and hashes
 * 3-byte prefixes: include/Node.h:158-185) ---- */
static uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static uint64_t xxh64_short(const uint8_t* p, size_t len, uint64_t seed) { /* len < 4 only */
    const uint64_t P1 = 0x9E3779B185EBCA87ULL, P2 = 0xC2B2AE3D27D4EB4FULL, P3 = 0x165667B19E3779F9ULL, P5 = 0x27D4EB2F165667C5ULL;
    uint64_t h = seed + P5 + (uint64_t)len;
    for (size_t i = 0; i < len; i++) {
        h ^= (uint64_t)p[i] * P5;
        h = rotl64(h, 11) * P1;
    }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

typedef struct {
    int nb_bits_skip2, nb_bits_skip3, nb_ucs_skp, nb_kmers_uc, level_min, modulo_hash, tresh_suf_pref;
    int size_kmer_in_bytes, size_kmer_in_bytes_minus_1, mask_shift_kmer, root;
} o_level; /* info_per_level, include/Node.h:37-51 */

typedef struct {
    uint8_t* suffixes; /* n lines of (size_sub + size_annot) bytes, then nb_ext * 3 bytes (include/UC.h:13-20) */
    int n, size_sub, size_annot, nb_ext, flag;
} o_uc;

struct o_node;
typedef struct {
    int type8, s, p, nb_elem, nb_nodes;
    uint8_t *bf, *f2, *f3, *ef3, *ct;
    o_uc* buckets;
    int nb_skp;
    struct o_node* nodes;
    int* first_line; /* per prefix: index of its first suffix line inside its bucket, -1 if it has a child Node */
    int* node_of;    /* per prefix: index of its child Node, -1 if inline */
} o_cc; /* CC, include/CC.h:34-67 */

typedef struct o_node {
    o_uc uc;
    int n_cc;
    o_cc* ccs;
} o_node; /* Node, include/Node.h:55-58 */

typedef struct { int64_t last_index; int size_annot; uint8_t* bytes; } o_pool; /* annotation_array_elem, include/annotation.h:51-55 */

struct o_bft {
    int k, n_genomes, n_levels, n_pools, r1, r2, compressed;
    char** names;
    o_pool* pools;
    o_level lvl[16];
    uint64_t* hv; /* hash_v for the 16384 reachable indices, 2 per index */
    o_node root;
    uint8_t rev[256]; /* rev[]: reverses the four 2-bit groups of a byte (src/popcnt.c) */
    /* parser state */
    const uint8_t* buf; size_t len, pos;
    int failed; char* err; size_t errlen;
};

static void o_fail(o_bft* b, const char* msg) {
    if (!b->failed && b->err) snprintf(b->err, b->errlen, "bft_oracle: %s (offset %zu)", msg, b->pos);
    b->failed = 1;
}
static const uint8_t* rd(o_bft* b, size_t n) {
    static const uint8_t zeros[64] = {0};
    if (b->failed || n > b->len - b->pos) { o_fail(b, "truncated file"); return n <= sizeof zeros ? zeros : NULL; }
    const uint8_t* p = b->buf + b->pos;
    b->pos += n;
    return p;
}
static int rd_i32(o_bft* b) { int v = 0; const uint8_t* p = rd(b, 4); if (p) memcpy(&v, p, 4); return v; }
static int rd_u16(o_bft* b) { uint16_t v = 0; const uint8_t* p = rd(b, 2); if (p) memcpy(&v, p, 2); return v; }
static uint8_t* dup_bytes(o_bft* b, size_t n) {
    const uint8_t* p = rd(b, n);
    uint8_t* q = (uint8_t*)calloc(n + 8, 1);
    if (p && q && !b->failed) memcpy(q, p, n);
    return q;
}

/* create_info_per_level, src/CC.c:1883-1993 */
static void make_level(o_level* l, int size, int size_max) {
    l->size_kmer_in_bytes = CEIL(size * 2, 8);
    l->size_kmer_in_bytes_minus_1 = size > 9 ? CEIL((size - 9) * 2, 8) : 0;
    l->root = size == size_max;
    static const int masks[4] = {0xff, 0x3, 0xf, 0x3f}; /* sizes 9,18,27,36 then repeating every 36 */
    l->mask_shift_kmer = masks[(size / 9 - 1) % 4];
}

static void read_uc(o_bft* b, o_uc* u, int size_sub, int n) { /* read_UC, src/write_to_disk.c:384-534 (compressed == 0) */
    memset(u, 0, sizeof *u);
    u->n = n;
    u->size_sub = size_sub;
    if (!n) return;
    u->nb_ext = rd_u16(b);
    u->size_annot = rd_i32(b);
    if (u->nb_ext == 0xffff || u->size_annot < 0 || u->size_annot > (1 << 24)) { o_fail(b, "unsupported UC encoding"); return; }
    u->suffixes = dup_bytes(b, (size_t)n * (size_t)(size_sub + u->size_annot) + (size_t)u->nb_ext * SIZE_BYTE_EXT_ANNOT);
}

static void read_node(o_bft* b, o_node* nd, int size);

static int get_nb_elts(const o_cc* cc, int pos) { /* getNbElts, include/CC.h:358-366 */
    if (cc->type8) return cc->ct[pos];
    return (pos & 1) ? cc->ct[pos / 2] >> 4 : cc->ct[pos / 2] & 0xf;
}

/* cluster-start flag of prefix j: extra_filter3 bit where the level stores it explicitly, else the in-band bit —
 * bit 7 of the last suffix byte of the prefix's first line, or bit 0 of the child Node's UC nb_children
 * (findCluster, src/presenceNode.c:1654-1689 vs 1690-1812) */
static int cluster_flag(const o_bft* b, const o_cc* cc, int size, int j) {
    const o_level* l = &b->lvl[size / 9 - 1];
    if (l->level_min == 1) return (cc->ef3[j / 8] >> (j % 8)) & 1;
    if (cc->node_of[j] >= 0) return cc->nodes[cc->node_of[j]].uc.flag;
    const o_uc* u = &cc->buckets[j / l->nb_ucs_skp];
    return u->suffixes[(size_t)cc->first_line[j] * (size_t)(u->size_sub + u->size_annot) + u->size_sub - 1] >> 7;
}

static void read_cc(o_bft* b, o_cc* cc, int size) { /* read_CC, src/write_to_disk.c:536-776 */
    const int lvl = size / 9 - 1;
    const o_level* l = &b->lvl[lvl];
    memset(cc, 0, sizeof *cc);
    const int type = rd_u16(b);
    cc->nb_elem = rd_u16(b);
    cc->nb_nodes = rd_u16(b);
    cc->type8 = (type >> 6) & 1;
    cc->s = (type >> 1) & 0x1f;
    cc->p = 2 * NB_CHAR_SUF_PREF - cc->s;
    if (cc->s != 8 && cc->s != 4) { o_fail(b, "unexpected p_v width"); return; }
    cc->nb_skp = CEIL(cc->nb_elem, l->nb_ucs_skp);
    cc->f2 = dup_bytes(b, (size_t)(1 << cc->p) / 8);
    cc->f3 = dup_bytes(b, cc->s == 8 ? (size_t)cc->nb_elem : (size_t)CEIL(cc->nb_elem, 2));
    if (l->level_min == 1) cc->ef3 = dup_bytes(b, (size_t)CEIL(cc->nb_elem, 8));
    cc->buckets = (o_uc*)calloc((size_t)cc->nb_skp + 1, sizeof(o_uc));
    cc->nodes = (o_node*)calloc((size_t)cc->nb_nodes + 1, sizeof(o_node));
    cc->first_line = (int*)calloc((size_t)cc->nb_elem + 1, sizeof(int));
    cc->node_of = (int*)calloc((size_t)cc->nb_elem + 1, sizeof(int));
    if (lvl) {
        cc->ct = dup_bytes(b, cc->type8 ? (size_t)cc->nb_elem : (size_t)CEIL(cc->nb_elem, 2));
        for (int i = 0; i < cc->nb_skp && !b->failed; i++) read_uc(b, &cc->buckets[i], l->size_kmer_in_bytes_minus_1, rd_u16(b));
        int node = 0, line = 0;
        for (int j = 0; j < cc->nb_elem && !b->failed; j++) {
            if (j % l->nb_ucs_skp == 0) line = 0;
            const int ne = get_nb_elts(cc, j);
            cc->first_line[j] = ne ? line : -1;
            cc->node_of[j] = ne ? -1 : node++;
            line += ne;
        }
    } else {
        for (int i = 0; i < cc->nb_skp && !b->failed; i++) {
            const int cnt = i != cc->nb_skp - 1 ? l->nb_ucs_skp : cc->nb_elem - i * l->nb_ucs_skp;
            read_uc(b, &cc->buckets[i], 0, cnt);
        }
        for (int j = 0; j < cc->nb_elem; j++) { cc->first_line[j] = j % l->nb_ucs_skp; cc->node_of[j] = -1; }
    }
    for (int i = 0; i < cc->nb_nodes && !b->failed; i++) read_node(b, &cc->nodes[i], size - NB_CHAR_SUF_PREF);
    if (b->failed) return;
    /* the Bloom filter is not stored: re-insert every stored prefix (src/write_to_disk.c:656-683, 696-772) */
    cc->bf = (uint8_t*)calloc((size_t)CEIL(l->modulo_hash, 8) + 1, 1);
    int j = 0;
    for (int pu = 0; pu < (1 << cc->p); pu++) {
        if (!((cc->f2[pu / 8] >> (pu % 8)) & 1)) continue;
        int first = 1;
        while (j < cc->nb_elem && (first || !cluster_flag(b, cc, size, j))) {
            const uint32_t pv = cc->s == 8 ? cc->f3[j] : ((j & 1) ? cc->f3[j / 2] >> 4 : cc->f3[j / 2] & 0xf);
            const uint32_t sp = ((((uint32_t)pu) << cc->s) | pv) >> 4; /* compressed <= 0: 14-bit index */
            const uint32_t h1 = (uint32_t)(b->hv[sp * 2] % (uint64_t)l->modulo_hash), h2 = (uint32_t)(b->hv[sp * 2 + 1] % (uint64_t)l->modulo_hash);
            cc->bf[h1 / 8] |= (uint8_t)(1 << (h1 % 8));
            cc->bf[h2 / 8] |= (uint8_t)(1 << (h2 % 8));
            j++;
            first = 0;
        }
    }
}

static void read_node(o_bft* b, o_node* nd, int size) { /* read_Node, src/write_to_disk.c:359-382 */
    memset(nd, 0, sizeof *nd);
    const int raw = rd_u16(b);
    read_uc(b, &nd->uc, b->lvl[size / 9 - 1].size_kmer_in_bytes, raw >> 1);
    nd->uc.flag = raw & 1;
    uint32_t n = 0;
    const uint8_t* p = rd(b, 4);
    if (p) memcpy(&n, p, 4);
    if (b->failed || n > 100000) { o_fail(b, "corrupt CC count"); return; }
    nd->n_cc = (int)n;
    nd->ccs = (o_cc*)calloc((size_t)n + 1, sizeof(o_cc));
    for (uint32_t i = 0; i < n && !b->failed; i++) read_cc(b, &nd->ccs[i], size);
}

static void free_node(o_node* nd) {
    free(nd->uc.suffixes);
    for (int i = 0; i < nd->n_cc; i++) {
        o_cc* cc = &nd->ccs[i];
        free(cc->bf); free(cc->f2); free(cc->f3); free(cc->ef3); free(cc->ct); free(cc->first_line); free(cc->node_of);
        if (cc->buckets) for (int u = 0; u < cc->nb_skp; u++) free(cc->buckets[u].suffixes);
        free(cc->buckets);
        if (cc->nodes) for (int n = 0; n < cc->nb_nodes; n++) free_node(&cc->nodes[n]);
        free(cc->nodes);
    }
    free(nd->ccs);
}

void o_free(o_bft* b) {
    if (!b) return;
    free_node(&b->root);
    if (b->names) for (int i = 0; i < b->n_genomes; i++) free(b->names[i]);
    free(b->names);
    if (b->pools) for (int i = 0; i < b->n_pools; i++) free(b->pools[i].bytes);
    free(b->pools);
    free(b->hv);
    free(b);
}

int o_k(const o_bft* b) { return b->k; }
int o_n_genomes(const o_bft* b) { return b->n_genomes; }

o_bft* o_load(const char* path, char* err, size_t errlen) {
    FILE* f = fopen(path, "rb");
    if (!f) { if (err) snprintf(err, errlen, "bft_oracle: cannot open %s", path); return NULL; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* buf = (uint8_t*)malloc((size_t)n + 1);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(buf); if (err) snprintf(err, errlen, "bft_oracle: read error"); return NULL; }
    fclose(f);
    o_bft* b = (o_bft*)calloc(1, sizeof *b);
    b->buf = buf; b->len = (size_t)n; b->err = err; b->errlen = errlen;
    for (int i = 0; i < 256; i++) b->rev[i] = (uint8_t)(((i & 3) << 6) | ((i & 0xc) << 2) | ((i & 0x30) >> 2) | ((i & 0xc0) >> 6));
    /* read_BFT_Root_offset, src/write_to_disk.c:281-350 */
    b->n_pools = rd_i32(b);
    if (b->n_pools < 0 || b->n_pools > (1 << 20)) o_fail(b, "not a .bft file");
    b->pools = (o_pool*)calloc((size_t)(b->failed ? 0 : b->n_pools) + 1, sizeof(o_pool));
    for (int i = 0; i < b->n_pools && !b->failed; i++) {
        const uint8_t* p = rd(b, 8);
        if (p) memcpy(&b->pools[i].last_index, p, 8);
        b->pools[i].size_annot = rd_i32(b);
        int64_t cnt = i ? b->pools[i].last_index - b->pools[i - 1].last_index : b->pools[i].last_index + 1;
        if (cnt < 0 || b->pools[i].size_annot < 0) { o_fail(b, "corrupt colour pool"); break; }
        b->pools[i].bytes = dup_bytes(b, (size_t)cnt * (size_t)b->pools[i].size_annot);
    }
    b->r1 = rd_i32(b); b->r2 = rd_i32(b);
    rd_i32(b); /* treshold_compression */
    b->n_genomes = rd_i32(b);
    b->k = rd_i32(b);
    const uint8_t* pc = rd(b, 1);
    b->compressed = pc ? *pc : 0;
    if (!b->failed && (b->k <= 0 || b->k % 9 || b->k > 126 || b->compressed || b->n_genomes < 0)) o_fail(b, "unsupported .bft header");
    if (!b->failed) {
        b->names = (char**)calloc((size_t)b->n_genomes + 1, sizeof(char*));
        for (int i = 0; i < b->n_genomes && !b->failed; i++) {
            int sl = rd_u16(b);
            b->names[i] = (char*)dup_bytes(b, (size_t)sl);
        }
        b->n_levels = b->k / 9;
        for (int i = 0; i < b->n_levels; i++) {
            o_level* l = &b->lvl[i];
            l->nb_bits_skip2 = rd_i32(b); l->nb_bits_skip3 = rd_i32(b); l->nb_ucs_skp = rd_i32(b); l->nb_kmers_uc = rd_i32(b);
            l->level_min = rd_i32(b); l->modulo_hash = rd_i32(b); l->tresh_suf_pref = rd_i32(b);
            make_level(l, 9 * (i + 1), b->k);
            if (!b->failed && (l->nb_ucs_skp <= 0 || l->modulo_hash <= 0)) o_fail(b, "corrupt level table");
        }
        /* create_hash_v_array, include/Node.h:158-185, restricted to the indices the uncompressed mode can reach */
        b->hv = (uint64_t*)malloc(16384 * 2 * sizeof(uint64_t));
        for (uint32_t i = 0; i < 16384; i++) {
            uint8_t g[3] = {(uint8_t)(i >> 10), (uint8_t)(i >> 2), (uint8_t)(i << 6)};
            b->hv[2 * i] = xxh64_short(g, 3, (uint64_t)(int64_t)b->r1);
            b->hv[2 * i + 1] = xxh64_short(g, 3, (uint64_t)(int64_t)b->r2);
        }
        if (!b->failed && b->pos < b->len) read_node(b, &b->root, b->k);
    }
    free(buf);
    b->buf = NULL;
    if (b->failed) { o_free(b); return NULL; }
    return b;
}

/* ---- suffix-line search ------------------------------------------------------------------------------------ */
/* binary_search_UC, src/UC.c:81-124 */
static int binary_search_uc(const o_uc* u, int pos_start, int pos_end, const uint8_t* suf, int nbytes, uint8_t mask) {
    if (!u->suffixes) return 0;
    int imin = pos_start, imax = pos_end;
    const int size_line = nbytes + u->size_annot;
    while (imin < imax) {
        const int imid = imin + (imax - imin) / 2;
        const uint8_t* line = &u->suffixes[(size_t)imid * size_line];
        if (mask == 0xff) {
            if (memcmp(line, suf, (size_t)nbytes) < 0) imin = imid + 1; else imax = imid;
        } else {
            const int cmp = memcmp(line, suf, (size_t)nbytes - 1);
            if (cmp < 0 || (cmp == 0 && (line[nbytes - 1] & mask) < suf[nbytes - 1])) imin = imid + 1; else imax = imid;
        }
    }
    return imin;
}

typedef struct { const o_uc* uc; int line; int size_sub; } o_hit; /* where the annotation of a found k-mer lives */

typedef struct { int kind; const o_cc* cc; int pos; const o_node* child; } o_probe; /* kind: 0 absent, 1 search node UC, 2 prefix at pos */

/* the prefix part of presenceKmer (src/presenceNode.c:1326-1410, 1470-1488): Bloom chain, filter2, cluster, filter3.
 * sub = the 9 leading nucleotides MSB-first: sub[0], sub[1], sub[2] & 0xc0 */
static o_probe probe_prefix(const o_bft* b, const o_node* nd, int size, const uint8_t sub_in[3]) {
    o_probe r = {1, NULL, 0, NULL};
    const o_level* l = &b->lvl[size / 9 - 1];
    const uint32_t sp = ((uint32_t)sub_in[0] << 8) | sub_in[1];
    const uint32_t idx = sp & 0x3fff; /* nuc1..nuc7 */
    const uint32_t h1 = (uint32_t)(b->hv[idx * 2] % (uint64_t)l->modulo_hash), h2 = (uint32_t)(b->hv[idx * 2 + 1] % (uint64_t)l->modulo_hash);
    for (int i = 0; i < nd->n_cc; i++) {
        const o_cc* cc = &nd->ccs[i];
        if (!((cc->bf[h1 / 8] >> (h1 % 8)) & 1) || !((cc->bf[h2 / 8] >> (h2 % 8)) & 1)) continue;
        /* first CC whose filter fires decides; prefix rotated to nuc1..nuc8,nuc0 (:1367-1371) */
        uint8_t sub[3];
        sub[0] = (uint8_t)(sp >> 6);
        sub[1] = (uint8_t)((sp << 2) | (sub_in[2] >> 6));
        sub[2] = (uint8_t)((sp >> 8) & 0xc0);
        const int pu = cc->p == 10 ? (sub[0] << 2) | (sub[1] >> 6) : (sub[0] << 6) | (sub[1] >> 2);
        r.kind = 0;
        r.cc = cc;
        if (!((cc->f2[pu / 8] >> (pu % 8)) & 1)) return r;
        int rank = 0; /* findCluster: Hamming weight of filter2[0..pu] */
        for (int q = 0; q <= pu; q++) rank += (cc->f2[q / 8] >> (q % 8)) & 1;
        int pos = 0, seen = 0; /* select: the rank-th cluster start */
        for (; pos < cc->nb_elem; pos++) {
            seen += cluster_flag(b, cc, size, pos);
            if (seen == rank) break;
        }
        if (pos >= cc->nb_elem) return r;
        int hw0 = 0;
        while (pos + hw0 + 1 < cc->nb_elem && !cluster_flag(b, cc, size, pos + hw0 + 1)) hw0++;
        const int pv = cc->s == 8 ? (uint8_t)((sub[1] << 2) | (sub[2] >> 6)) : ((sub[1] & 3) << 2) | (sub[2] >> 6);
        int imin = pos, imax = pos + hw0;
        while (imin < imax) {
            const int imid = (imin + imax) / 2;
            const int t = cc->s == 8 ? cc->f3[imid] : ((imid & 1) ? cc->f3[imid / 2] >> 4 : cc->f3[imid / 2] & 0xf);
            if (t < pv) imin = imid + 1; else imax = imid;
        }
        const int t = cc->s == 8 ? cc->f3[imin] : ((imin & 1) ? cc->f3[imin / 2] >> 4 : cc->f3[imin / 2] & 0xf);
        if (t != pv) return r;
        r.kind = 2;
        r.pos = imin;
        r.child = cc->node_of[imin] >= 0 ? &cc->nodes[cc->node_of[imin]] : NULL;
        return r;
    }
    return r;
}

static void msb_first_prefix(const o_bft* b, const uint8_t* kmer, uint8_t sub[3]) { /* :1327-1329 */
    sub[0] = b->rev[kmer[0]];
    sub[1] = b->rev[kmer[1]];
    sub[2] = b->rev[kmer[2]] & 0xc0;
}

/* drop the 9 leading nucleotides (isKmerPresent, src/presenceNode.c:1853-1861) */
static void shift_level(const o_bft* b, uint8_t* kmer, int size) {
    const o_level* l = &b->lvl[size / 9 - 1];
    const int nb_cell = l->size_kmer_in_bytes;
    const int del = 2 + (size == 45 || size == 81 || size == 117);
    int j;
    for (j = 0; j < nb_cell - del; j++) {
        kmer[j] = kmer[j + 2] >> 2;
        if (j + 3 < nb_cell) kmer[j] |= (uint8_t)(kmer[j + 3] << 6);
    }
    kmer[j - 1] &= (uint8_t)l->mask_shift_kmer;
}

/* isKmerPresent, src/presenceNode.c:1823-1921 */
static int find_kmer(const o_bft* b, const o_node* nd, const uint8_t* kmer_in, int size, o_hit* hit) {
    const o_level* l = &b->lvl[size / 9 - 1];
    uint8_t kmer[40];
    memcpy(kmer, kmer_in, (size_t)l->size_kmer_in_bytes);
    uint8_t sub[3];
    msb_first_prefix(b, kmer, sub);
    const o_probe pr = probe_prefix(b, nd, size, sub);
    if (pr.kind == 0) return 0;
    if (pr.kind == 1) { /* node UC, whole remainder (:1554-1573) */
        const o_uc* u = &nd->uc;
        if (!u->suffixes) return 0;
        const int nb = l->size_kmer_in_bytes, pos = binary_search_uc(u, 0, u->n - 1, kmer, nb, 0xff);
        if (memcmp(&u->suffixes[(size_t)pos * (nb + u->size_annot)], kmer, (size_t)nb)) return 0;
        hit->uc = u; hit->line = pos; hit->size_sub = nb;
        return 1;
    }
    if (size == NB_CHAR_SUF_PREF) { /* leaf: one annotation per prefix (:1453-1463) */
        hit->uc = &pr.cc->buckets[pr.pos / l->nb_ucs_skp];
        hit->line = pr.pos % l->nb_ucs_skp;
        hit->size_sub = 0;
        return 1;
    }
    shift_level(b, kmer, size);
    if (pr.child) return find_kmer(b, pr.child, kmer, size - NB_CHAR_SUF_PREF, hit);
    const o_uc* u = &pr.cc->buckets[pr.pos / l->nb_ucs_skp];
    const int nb = l->size_kmer_in_bytes_minus_1, first = pr.cc->first_line[pr.pos], ne = get_nb_elts(pr.cc, pr.pos);
    const uint8_t mask = (size == 45 || size == 81 || size == 117) ? 0xff : 0x7f; /* :1887-1914 */
    const int j = binary_search_uc(u, first, first + ne - 1, kmer, nb, mask);
    const uint8_t* line = &u->suffixes[(size_t)j * (nb + u->size_annot)];
    if (memcmp(line, kmer, (size_t)nb - 1) || (line[nb - 1] & mask) != kmer[nb - 1]) return 0;
    hit->uc = u; hit->line = j; hit->size_sub = nb;
    return 1;
}

/* ---- enumeration: iterate_over_kmers_from_node (src/extract_kmers.c:3-597) --------------------------------------
 * Order of the reference: the CCs of a Node in order; inside a CC the stored prefixes in order (clusters by ascending
 * p_u, inside a cluster ascending p_v); for a prefix its inline suffix lines in order, or the k-mers of its child
 * Node; after the CCs the Node's own UC lines. K-mers are rebuilt as ASCII: the path above the Node, the prefix
 * un-rotated from (p_u, p_v) (:84-85), the suffix nucleotides of the line. */
typedef struct { char* out; size_t n, cap; int k; } o_emit;

static void emit_kmer(o_emit* e, const char* acc, int acc_len, const uint8_t* packed, int n_nuc) {
    if (e->n < e->cap) {
        char* d = e->out + e->n * (size_t)e->k;
        memcpy(d, acc, (size_t)acc_len);
        for (int j = 0; j < n_nuc; j++) d[acc_len + j] = "ACGT"[(packed[j >> 2] >> (2 * (j & 3))) & 3];
    }
    e->n++;
}

static void extract_node(const o_bft* b, const o_node* nd, int size, char* acc, int acc_len, o_emit* e) {
    const o_level* l = &b->lvl[size / 9 - 1];
    for (int i = 0; i < nd->n_cc; i++) {
        const o_cc* cc = &nd->ccs[i];
        int pu = -1;
        for (int pos = 0; pos < cc->nb_elem; pos++) {
            if (cluster_flag(b, cc, size, pos)) { /* first prefix of the next cluster: the next set bit of filter2 */
                do pu++; while (pu < (1 << cc->p) && !((cc->f2[pu / 8] >> (pu % 8)) & 1));
            }
            const uint32_t pv = cc->s == 8 ? cc->f3[pos] : ((pos & 1) ? cc->f3[pos / 2] >> 4 : cc->f3[pos / 2] & 0xf);
            const uint32_t rot = ((uint32_t)pu << cc->s) | pv;             /* nuc1..nuc8, nuc0 */
            const uint32_t r18 = (rot >> 2) | ((rot & 3u) << 16);          /* nuc0..nuc8, MSB first */
            for (int j = 0; j < 9; j++) acc[acc_len + j] = "ACGT"[(r18 >> (2 * (8 - j))) & 3];
            if (size == NB_CHAR_SUF_PREF) { emit_kmer(e, acc, acc_len + 9, NULL, 0); continue; }
            if (cc->node_of[pos] >= 0) { extract_node(b, &cc->nodes[cc->node_of[pos]], size - 9, acc, acc_len + 9, e); continue; }
            const o_uc* u = &cc->buckets[pos / l->nb_ucs_skp];
            const int nb = l->size_kmer_in_bytes_minus_1, first = cc->first_line[pos], ne = get_nb_elts(cc, pos);
            for (int q = first; q < first + ne; q++)
                emit_kmer(e, acc, acc_len + 9, &u->suffixes[(size_t)q * (nb + u->size_annot)], size - 9);
        }
    }
    const o_uc* u = &nd->uc;
    for (int q = 0; q < u->n; q++)
        emit_kmer(e, acc, acc_len, &u->suffixes[(size_t)q * (l->size_kmer_in_bytes + u->size_annot)], size);
}

size_t o_extract_kmers(o_bft* b, char* out, size_t cap) {
    char acc[160];
    o_emit e = {out, 0, out ? cap : 0, b->k};
    extract_node(b, &b->root, b->k, acc, 0, &e);
    return e.n;
}

/* ---- annotation -------------------------------------------------------------------------------------------- */
/* get_annot + get_extend_annot, src/UC.c:171-239, 501-521 */
static void get_annot(const o_hit* h, const uint8_t** annot, int* size_annot, const uint8_t** ext) {
    const o_uc* u = h->uc;
    const int size_line = h->size_sub + u->size_annot;
    *ext = NULL;
    if (u->nb_ext) {
        const uint8_t* e = &u->suffixes[(size_t)u->n * size_line];
        int pos = (e[0] << 8) | e[1], i = SIZE_BYTE_EXT_ANNOT;
        while (pos < h->line && i < u->nb_ext * SIZE_BYTE_EXT_ANNOT) {
            pos += (e[i] << 8) | e[i + 1];
            i += SIZE_BYTE_EXT_ANNOT;
        }
        if (pos == h->line) *ext = &e[i - 1];
    }
    if (u->size_annot == 0) { *annot = NULL; *size_annot = 0; return; }
    *annot = &u->suffixes[(size_t)h->line * size_line + h->size_sub];
    if (*ext) { *size_annot = u->size_annot; return; }
    int n = u->size_annot; /* size_annot_sub, include/annotation.h:183-191 */
    while (n > 0 && (*annot)[n - 1] == 0) n--;
    *size_annot = n;
}

/* get_id_genomes_from_annot + decomp_annotation, src/annotation.c:2086-2250, 1840-1922 -> ids[0]=count, ids[1..] */
static int decode_ids(const o_bft* b, const uint8_t* annot, int size_annot, const uint8_t* ext, uint32_t* ids) {
    uint8_t tmp[1 << 12];
    uint8_t* a = tmp;
    uint8_t* heap = NULL;
    uint32_t n = 0;
    ids[0] = 0;
    if (size_annot == 0) return 0;
    int size = size_annot + (ext ? 1 : 0), delta = 0;
    if (size > (int)sizeof tmp) a = heap = (uint8_t*)malloc((size_t)size);
    memcpy(a, annot, (size_t)size_annot);
    if (ext) a[size_annot] = *ext;
    if ((a[0] & 3) == 3) { /* position into the colour pools (:2097-2112) */
        uint32_t position = a[0] >> 2;
        int i = 1;
        while (i < size && (a[i] & 1)) { position |= ((uint32_t)(a[i] >> 1)) << (6 + (i - 1) * 7); i++; }
        int p = 0; /* extract_from_annotation_array_elem, include/annotation.h:309-323 */
        while (p < b->n_pools && (int64_t)position > b->pools[p].last_index) p++;
        if (p >= b->n_pools) { free(heap); return -1; }
        const int64_t first = p ? b->pools[p - 1].last_index + 1 : 0;
        size = b->pools[p].size_annot;
        const uint8_t* src = b->pools[p].bytes + (size_t)((int64_t)position - first) * (size_t)size;
        free(heap);
        heap = NULL;
        a = tmp;
        if (size > (int)sizeof tmp) a = heap = (uint8_t*)malloc((size_t)size);
        memcpy(a, src, (size_t)size);
        delta = 1;
        if (size && (a[0] & 3) == 3) { free(heap); return -1; } /* ERROR "mode 3, should not happen" (:2247) */
    }
    const int mode = size ? a[0] & 3 : 0;
    if (size == 0) {
    } else if (mode == 0) {
        for (int i = 2; i < size * 8; i++)
            if (a[i / 8] & (1 << (i % 8))) ids[++n] = (uint32_t)(i - 2);
    } else {
        const uint8_t f1 = mode == 1 ? 1 : 2, f2 = mode == 1 ? 2 : 1;
        uint32_t raw[4096];
        uint32_t* v = raw;
        uint32_t* vheap = NULL;
        if (size > 4096) v = vheap = (uint32_t*)malloc((size_t)size * sizeof(uint32_t));
        int i = 0, m = 0;
        while (i < size && (a[i] & f1)) {
            uint32_t x = a[i] >> 2;
            i++;
            while (i < size && (a[i] & f2)) { x = (x << 6) | (a[i] >> 2); i++; }
            v[m++] = x;
        }
        if (delta) for (int q = 1; q < m; q++) v[q] += v[q - 1]; /* decomp_annotation prefix sums (:1877-1916) */
        if (mode == 2) {
            for (int q = 0; q < m; q++) ids[++n] = v[q];
        } else { /* (start, stop) pairs, inclusive (:2150-2176, 2193-2222) */
            for (int q = 0; q < m; q++) {
                if (q % 2 == 0) ids[++n] = v[q];
                else for (uint32_t z = v[q - 1] + 1; z <= v[q]; z++) ids[++n] = z;
            }
        }
        free(vheap);
    }
    free(heap);
    ids[0] = n;
    return 0;
}

int o_query_kmer(o_bft* b, const uint8_t* kmer, uint32_t* ids) {
    o_hit h;
    ids[0] = 0;
    if (!find_kmer(b, &b->root, kmer, b->k, &h)) return 0;
    const uint8_t *annot, *ext;
    int size;
    get_annot(&h, &annot, &size, &ext);
    decode_ids(b, annot, size, ext, ids);
    return 1;
}

/* ---- branching --------------------------------------------------------------------------------------------- */
/* isBranchingRight below the root shift (src/branchingNode.c:43-104) with presenceNeighborsRight
 * (src/presenceNode.c:676-1211): kmer = the successor with its last nucleotide set to 0. */
static int branching_right_rec(const o_bft* b, const o_node* nd, const uint8_t* kmer_in, int size) {
    const o_level* l = &b->lvl[size / 9 - 1];
    uint8_t kmer[40];
    memcpy(kmer, kmer_in, (size_t)l->size_kmer_in_bytes);
    uint8_t sub[3];
    msb_first_prefix(b, kmer, sub);
    int count = 0;
    if (size == NB_CHAR_SUF_PREF) {
        sub[1] &= 0xfc; /* :721 — the reference clears nucleotide 7 here */
        const uint32_t sp = ((uint32_t)sub[0] << 8) | sub[1];
        const uint32_t idx = sp & 0x3fff;
        const uint32_t h1 = (uint32_t)(b->hv[idx * 2] % (uint64_t)l->modulo_hash), h2 = (uint32_t)(b->hv[idx * 2 + 1] % (uint64_t)l->modulo_hash);
        for (int i = 0; i < nd->n_cc; i++) {
            const o_cc* cc = &nd->ccs[i];
            if (!((cc->bf[h1 / 8] >> (h1 % 8)) & 1) || !((cc->bf[h2 / 8] >> (h2 % 8)) & 1)) continue;
            /* the four candidates differ in nuc8 = bits 2-3 of p_v; probe each */
            for (int c = 0; c < 4; c++) {
                uint8_t s3[3] = {sub[0], sub[1], (uint8_t)(c << 6)};
                /* probe_prefix would redo the chain with the same idx (nuc8 is not hashed): same CC */
                const o_probe pr = probe_prefix(b, nd, size, s3);
                if (pr.kind == 2) count++;
            }
            return count;
        }
        /* node UC: lines equal on nuc0..nuc7 (the unmodified k-mer), any nuc8 (:1164-1208) */
        const o_uc* u = &nd->uc;
        if (u->suffixes)
            for (int q = 0; q < u->n; q++) {
                const uint8_t* line = &u->suffixes[(size_t)q * (l->size_kmer_in_bytes + u->size_annot)];
                if (!memcmp(line, kmer, (size_t)l->size_kmer_in_bytes - 1) && kmer[l->size_kmer_in_bytes - 1] == 0) count++;
            }
        return count > 4 ? 4 : count;
    }
    const o_probe pr = probe_prefix(b, nd, size, sub);
    if (pr.kind == 0) return 0;
    if (pr.kind == 1) { /* node UC scan, last nucleotide masked out (:1164-1208) */
        const o_uc* u = &nd->uc;
        if (!u->suffixes) return 0;
        const int nb = l->size_kmer_in_bytes;
        uint8_t mask = (uint8_t)l->mask_shift_kmer;
        if (mask == 0xff) mask = 0;
        for (int q = 0; q < u->n; q++) {
            const uint8_t* line = &u->suffixes[(size_t)q * (nb + u->size_annot)];
            if (!memcmp(line, kmer, (size_t)nb - 1) && (line[nb - 1] & mask) == kmer[nb - 1]) count++;
        }
        return count > 4 ? 4 : count;
    }
    shift_level(b, kmer, size);
    if (pr.child) return branching_right_rec(b, pr.child, kmer, size - NB_CHAR_SUF_PREF);
    const o_uc* u = &pr.cc->buckets[pr.pos / l->nb_ucs_skp];
    const int nb = l->size_kmer_in_bytes_minus_1, first = pr.cc->first_line[pr.pos], ne = get_nb_elts(pr.cc, pr.pos);
    uint8_t mask = (uint8_t)b->lvl[size / 9 - 2].mask_shift_kmer; /* info_per_lvl[lvl_node-1] (src/branchingNode.c:89-90) */
    if (mask == 0xff) mask = 0;
    for (int q = first; q < first + ne; q++) {
        const uint8_t* line = &u->suffixes[(size_t)q * (nb + u->size_annot)];
        if (!memcmp(line, kmer, (size_t)nb - 1) && (line[nb - 1] & mask) == kmer[nb - 1]) count++;
    }
    return count > 4 ? 4 : count;
}

int o_branching_right(o_bft* b, const uint8_t* kmer) {
    const int nb = b->lvl[b->n_levels - 1].size_kmer_in_bytes;
    uint8_t sh[40];
    for (int i = 0; i < nb; i++) { /* drop nucleotide 0 (src/branchingNode.c:43-48) */
        sh[i] = kmer[i] >> 2;
        if (i + 1 < nb) sh[i] |= (uint8_t)(kmer[i + 1] << 6);
    }
    return branching_right_rec(b, &b->root, sh, b->k);
}

/* isBranchingLeft (src/branchingNode.c:240-413) with presenceNeighborsLeft (src/presenceNode.c:15-662): at the
 * root the four predecessors share the hash index and p_u (the prefix is stored rotated, nuc0 last) and each one
 * found is followed like an ordinary lookup of the remainder. */
int o_branching_left(o_bft* b, const uint8_t* kmer) {
    const int nb = b->lvl[b->n_levels - 1].size_kmer_in_bytes;
    uint8_t sh[40];
    const uint8_t shifting = (uint8_t)(0xff >> (nb * 8 - b->k * 2));
    for (int i = nb - 1; i >= 0; i--) { /* :262-268 */
        sh[i] = (uint8_t)(kmer[i] << 2);
        if (i > 0) sh[i] |= kmer[i - 1] >> 6;
    }
    sh[nb - 1] &= shifting;
    int count = 0;
    for (int c = 0; c < 4; c++) {
        uint8_t cand[40];
        memcpy(cand, sh, (size_t)nb);
        cand[0] = (uint8_t)((cand[0] & 0xfc) | c);
        o_hit h;
        count += find_kmer(b, &b->root, cand, b->k, &h);
    }
    return count;
}

/* ---- sequences --------------------------------------------------------------------------------------------- */
static int parse_kmer(const char* s, int k, uint8_t* out) { /* parseKmerCount, src/fasta.c:3-53 */
    memset(out, 0, (size_t)CEIL(k * 2, 8));
    for (int i = 0; i < k; i++) {
        switch (s[i]) {
            case 'a': case 'A': break;
            case 'c': case 'C': out[i / 4] |= (uint8_t)(1 << (2 * (i % 4))); break;
            case 'g': case 'G': out[i / 4] |= (uint8_t)(2 << (2 * (i % 4))); break;
            case 'u': case 'U': case 't': case 'T': out[i / 4] |= (uint8_t)(3 << (2 * (i % 4))); break;
            default: return 0;
        }
    }
    return 1;
}

static int reverse_complement(const char* s1, char* s2, int length) { /* src/fasta.c:387-440 */
    static const char from[] = "aAcCgGuUtTmMrRwWsSyYkKvVhHdDbBnN";
    static const char to[] = "tTgGcCaAaAkKyYwWsSrRmMbBdDhHvVnN";
    for (int pos = length - 1, i = 0; pos >= 0; pos--, i++) {
        const char* p = s1[pos] ? strchr(from, s1[pos]) : NULL;
        if (!p) return 0;
        s2[i] = to[p - from];
    }
    s2[length] = 0;
    return 1;
}

int o_query_sequence(o_bft* b, const char* seq, double threshold, int canonical, uint32_t* ids) { /* src/bft.c:1241-1351 */
    const int k = b->k, G = b->n_genomes;
    uint32_t* count = (uint32_t*)calloc((size_t)G + 1, sizeof(uint32_t));
    uint32_t* kid = (uint32_t*)malloc(((size_t)G + 2) * sizeof(uint32_t));
    char fwd[128], rc[128];
    uint8_t packed[40];
    const int64_t n = (int64_t)strlen(seq) - k + 1;
    const int64_t need = (int64_t)ceil((double)n * threshold);
    int64_t found = 0;
    int rcode = 0;
    fwd[k] = 0;
    for (int64_t it = 0; it < n; it++) {
        memcpy(fwd, seq + it, (size_t)k);
        const char* kmer = fwd;
        if (canonical) {
            if (!reverse_complement(fwd, rc, k)) { rcode = -1; break; }
            if (strcmp(fwd, rc) >= 0) kmer = rc;
        }
        if (!strpbrk(kmer, "rRyYsSwWkKmMbBdDhHvVnN.-")) { /* is_substring_IUPAC, src/fasta.c:357-363 */
            if (!parse_kmer(kmer, k, packed)) { rcode = -1; break; }
            if (o_query_kmer(b, packed, kid)) {
                found++;
                for (uint32_t q = 1; q <= kid[0]; q++) count[kid[q]]++;
            }
        }
        if (found + n - it < need) break;
    }
    uint32_t m = 0;
    for (int g = 0; g < G; g++)
        if (count[g] && (int64_t)count[g] >= need) ids[++m] = (uint32_t)g;
    ids[0] = m;
    free(count);
    free(kid);
    return rcode;
}
