/* arena_host_query — TEST TOOL (not shipped in the product library).
 * Runs the flattened-arena walk (bft_arena.h) and the colour decoder (bft_colour.h) on the HOST so the serializer
 * and the walk logic can be checked against the reference CLI without a GPU:
 *   arena_host_query file.bft {kmers|kmers_comp} queries out.csv
 * writes the same CSV bytes as `bft load file.bft -query_kmers ...` (reference src/file_io.c:651-895). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bft_flatten.h"
#include "bft_colour.h"
#include "bft_io.h"

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s file.bft {kmers|kmers_comp} queries out.csv\n", argv[0]); return 2; }
    char err[256];
    bft_arena_t* a = bft_arena_from_file(argv[1], err, sizeof err);
    if (!a) { fprintf(stderr, "%s\n", err); return 1; }
    fprintf(stderr, "k=%d W=%d genomes=%d nodes=%zu ccs=%zu (max/node %d) depth=%d lines=%zu buckets=%zu overflow=%zu uc_lines=%zu leaf_prefixes=%zu kmers=%zu classes=%zu cls_shift=%d pools=%d arena=%.1f MB\n",
            a->k, a->W, a->n_genomes, a->n_nodes, a->n_ccs, a->max_cc_per_node, a->max_depth, a->n_lines, a->n_buckets, a->n_ovf, a->n_uc_lines, a->n_leaf_prefixes,
            a->n_kmers, a->n_classes, a->cls_shift, a->n_pools, bft_arena_bytes(a) / 1e6);
    bft_view_t v;
    bft_arena_view(a, &v);
    const int rw = (a->n_genomes + 31) / 32;
    uint32_t* rows = calloc((a->n_classes + 1) * (size_t)rw, 4);
    bft_pools_t pools = {a->n_pools, a->pool_last_index, a->pool_size_annot, a->pool_off, a->pool_bytes};
    for (size_t c = 0; c < a->n_classes; c++)
        if (bft_decode_annotation(a->cls_bytes + a->cls_off[c], (int)(a->cls_off[c + 1] - a->cls_off[c]), &pools, rows + c * rw, rw))
            fprintf(stderr, "class %zu: malformed annotation\n", c);
    uint64_t* q; size_t n;
    if (bft_read_kmer_file(argv[3], strcmp(argv[2], "kmers_comp") == 0, a->k, a->W, &q, &n)) { fprintf(stderr, "cannot read %s\n", argv[3]); return 1; }
    uint32_t* out = calloc((n + 1) * (size_t)rw, 4);
    size_t present = 0;
    for (size_t i = 0; i < n; i++) {
        uint32_t cls = bft_lookup(&v, q + i * a->W);
        if (cls != BFT_CLS_NONE) { present++; memcpy(out + i * rw, rows + (size_t)cls * rw, (size_t)rw * 4); }
    }
    FILE* f = fopen(argv[4], "w");
    bft_csv_write_header(f, a->filenames, a->n_genomes);
    bft_csv_write_rows(f, out, n, a->n_genomes, rw);
    bft_csv_finish(f);
    fclose(f);
    printf("Nb k-mers present = %zu\n", present);
    free(q); free(out); free(rows);
    bft_arena_free(a);
    return 0;
}
