/* arena_loc_check — TEST TOOL (host build of the arena walk, no GPU): looks every k-mer of a kmers_comp file up with
 * bft_lookup_loc and reports how many were found, how many distinct storage locations they ended in and the largest
 * one. The graph traversals key their vertex table on these locations (bft_arena.h, bft_graph.cuh): the stored k-mers
 * of a BFT must map to pairwise distinct locations below n_loc. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bft_flatten.h"
#include "bft_io.h"

static int cmp_u32(const void* a, const void* b) {
    const uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return x < y ? -1 : x > y;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s file.bft kmers.kc\n", argv[0]); return 2; }
    char err[256];
    bft_arena_t* a = bft_arena_from_file(argv[1], err, sizeof err);
    if (!a) { fprintf(stderr, "%s\n", err); return 1; }
    bft_view_t v;
    bft_arena_view(a, &v);
    uint64_t* q;
    size_t n;
    if (bft_read_kmer_file(argv[2], 1, a->k, a->W, &q, &n)) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    uint32_t* locs = malloc((n + 1) * sizeof(uint32_t));
    size_t found = 0, mismatch = 0;
    uint32_t max_loc = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t loc = 0xffffffffu;
        const uint32_t cls = bft_lookup_loc(&v, q + i * a->W, a->W, 0, NULL, &loc);
        if (cls != bft_lookup(&v, q + i * a->W)) mismatch++; /* the location output must not change the answer */
        if (cls == BFT_CLS_NONE) continue;
        if (loc > max_loc) max_loc = loc;
        locs[found++] = loc;
    }
    qsort(locs, found, sizeof(uint32_t), cmp_u32);
    size_t distinct = 0;
    for (size_t i = 0; i < found; i++) distinct += i == 0 || locs[i] != locs[i - 1];
    printf("queries=%zu found=%zu distinct_locs=%zu max_loc=%u n_loc=%llu stored=%zu answer_mismatch=%zu\n", n, found, distinct, max_loc,
           (unsigned long long)v.loc_leaf + a->n_pref, a->n_kmers, mismatch);
    free(locs); free(q);
    bft_arena_free(a);
    return 0;
}
