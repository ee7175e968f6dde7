/* arena_enum_check — TEST TOOL (host, no GPU): enumerates every k-mer of a flattened arena exactly the way the
 * extraction kernels do (k_extract_prefix_kmers / k_extract_uc_kmers in bft_kernels.cuh: slot of a line = pref_out of
 * its prefix + the line's rank in memcmp order of the suffix bytes; Node-UC lines at uc_out + uc_rank) and writes the
 * k-mers as ASCII, concatenated, to a file. tests/test_host.py compares that with the reference's iterate_over_kmers
 * order (golden digests + the oracle), so the serializer's enumeration tables are checked without a GPU. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bft_flatten.h"

static int W;
typedef struct { uint64_t be[BFT_MAX_WORDS]; uint64_t key[BFT_MAX_WORDS]; } line_t;

static int cmp_line(const void* pa, const void* pb) {
    const line_t* a = (const line_t*)pa;
    const line_t* b = (const line_t*)pb;
    for (int w = 0; w < W; w++) {
        if (a->be[w] < b->be[w]) return -1;
        if (a->be[w] > b->be[w]) return 1;
    }
    return 0;
}

static void or_shl(uint64_t* dst, const uint64_t* src, int sh) { /* dst |= src << sh over BFT_MAX_WORDS words */
    const int ws = sh >> 6, bs = sh & 63;
    for (int w = BFT_MAX_WORDS - 1; w >= 0; w--) {
        uint64_t x = 0;
        if (w - ws >= 0) x = src[w - ws] << bs;
        if (bs && w - ws - 1 >= 0) x |= src[w - ws - 1] >> (64 - bs);
        dst[w] |= x;
    }
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s file.bft out.txt\n", argv[0]); return 2; }
    char err[256];
    bft_arena_t* a = bft_arena_from_file(argv[1], err, sizeof err);
    if (!a) { fprintf(stderr, "%s\n", err); return 1; }
    W = a->W;
    const int k = a->k;
    uint64_t* out = calloc((a->n_kmers + 1) * BFT_MAX_WORDS, 8);
    uint8_t* seen = calloc(a->n_kmers + 1, 1);
    const uint64_t top_mask = a->cls_shift ? ((1ULL << a->cls_shift) - 1ULL) : ~BFT_SLOT_SPECIAL;
    size_t dup = 0;
    line_t lines[1024];
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
            const uint64_t o = a->pref_out[j];
            dup += seen[o]++;
            memcpy(out + o * BFT_MAX_WORDS, base, sizeof base);
            continue;
        }
        const uint32_t n_slots = BFT_BUCKET_KEYS * BFT_INLINE_NBK(e);
        uint32_t n = 0;
        for (uint32_t s = 0; s < n_slots; s++) {
            const uint64_t* p = a->buckets + ((size_t)e.a * BFT_BUCKET_KEYS + s) * W;
            const uint64_t top = p[W - 1];
            if (!(top & BFT_SLOT_SPECIAL)) {
                memset(&lines[n], 0, sizeof lines[n]);
                for (int w = 0; w < W; w++) lines[n].key[w] = p[w];
                lines[n].key[W - 1] &= top_mask;
                n++;
            } else if (top != BFT_SLOT_EMPTY) {
                const uint32_t cnt = (uint32_t)(top >> 32) & 0x7fffffffu, start = (uint32_t)top;
                for (uint32_t i = 0; i < cnt; i++) {
                    memset(&lines[n], 0, sizeof lines[n]);
                    for (int w = 0; w < W; w++) lines[n].key[w] = a->ovf[((size_t)start + i) * W + w];
                    lines[n].key[W - 1] &= top_mask;
                    n++;
                }
            }
        }
        if (n != BFT_INLINE_CNT(e)) { fprintf(stderr, "prefix %zu: %u lines gathered, %u declared\n", j, n, BFT_INLINE_CNT(e)); return 1; }
        for (uint32_t i = 0; i < n; i++)
            for (int w = 0; w < W; w++) lines[i].be[w] = __builtin_bswap64(lines[i].key[w]);
        qsort(lines, n, sizeof(line_t), cmp_line);
        for (uint32_t i = 0; i < n; i++) {
            uint64_t km[BFT_MAX_WORDS];
            memcpy(km, base, sizeof km);
            or_shl(km, lines[i].key, BFT_PREFIX_BITS * (int)(path->depth + 1));
            const uint64_t o = a->pref_out[j] + i;
            dup += seen[o]++;
            memcpy(out + o * BFT_MAX_WORDS, km, sizeof km);
        }
    }
    for (size_t nid = 0; nid < a->n_nodes; nid++) {
        const bft_node_t* nd = &a->nodes[nid];
        const bft_path_t* path = &a->node_path[nid];
        const uint64_t o0 = ((uint64_t)path->uc_out_hi << 32) | path->uc_out_lo;
        for (uint32_t i = 0; i < nd->uc_n; i++) {
            uint64_t km[BFT_MAX_WORDS], key[BFT_MAX_WORDS] = {0};
            memcpy(km, path->acc, sizeof km);
            for (int w = 0; w < W; w++) key[w] = a->uckeys[((size_t)nd->uc_begin + i) * W + w];
            or_shl(km, key, BFT_PREFIX_BITS * (int)path->depth);
            const uint64_t o = o0 + a->uc_rank[nd->uc_begin + i];
            dup += seen[o]++;
            memcpy(out + o * BFT_MAX_WORDS, km, sizeof km);
        }
    }
    size_t missing = 0;
    for (size_t i = 0; i < a->n_kmers; i++) missing += !seen[i];
    FILE* f = fopen(argv[2], "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    char line[160];
    for (size_t i = 0; i < a->n_kmers; i++) {
        for (int c = 0; c < k; c++) line[c] = "ACGT"[(out[i * BFT_MAX_WORDS + (size_t)(c >> 5)] >> (2 * (c & 31))) & 3];
        fwrite(line, 1, (size_t)k, f);
    }
    fclose(f);
    printf("kmers=%zu duplicates=%zu missing=%zu\n", a->n_kmers, dup, missing);
    free(out); free(seen);
    bft_arena_free(a);
    return 0;
}
