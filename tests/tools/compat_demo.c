/* compat_demo — TEST TOOL: a caller written against the reference's public API names (include/bft.h) but compiled
 * with include/bft_compat.h and linked with libbft_b200.so. Prints, for every k-mer of a text file:
 *   presence, the ascending genome ids (get_annotation + get_list_id_genomes), get_count_id_genomes,
 *   presence_genome of genome 0, and the presence flags of get_neighbors (4 predecessors, 4 successors);
 * then for every line of a sequence file the ids query_sequence returns. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bft_compat.h"

static size_t count_and_check(BFT_kmer* km, BFT* g, va_list args) { /* a BFT_func_ptr, as in the reference's snippets */
    size_t* n = va_arg(args, size_t*);
    size_t* colours = va_arg(args, size_t*);
    BFT_annotation* a = get_annotation(km);
    *colours += get_count_id_genomes(a, g);
    free_BFT_annotation(a);
    (*n)++;
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    BFT* g = load_BFT(argv[1]);
    double thr = atof(argv[4]);
    int canonical = atoi(argv[5]);
    printf("k=%d genomes=%d first=%s\n", g->k, g->nb_genomes, g->nb_genomes ? g->filenames[0] : "");
    char line[1 << 16];
    FILE* f = fopen(argv[2], "r");
    set_neighbors_traversal(g);
    while (fgets(line, sizeof line, f)) {
        line[strcspn(line, "\r\n")] = 0;
        BFT_kmer* km = get_kmer(line, g);
        if (!is_kmer_in_cdbg(km)) {
            printf("K 0\n");
        } else {
            BFT_annotation* a = get_annotation(km);
            uint32_t* ids = get_list_id_genomes(a, g);
            printf("K 1 n=%u c=%u g0=%d ids", ids[0], get_count_id_genomes(a, g), (int)presence_genome(0, a, g));
            for (uint32_t i = 1; i <= ids[0]; i++) printf(" %u", ids[i]);
            BFT_kmer* nb = get_neighbors(km, g);
            printf(" nb");
            for (int i = 0; i < 8; i++) printf(" %d", (int)is_kmer_in_cdbg(&nb[i]));
            BFT_kmer* pr = get_predecessors(km, g);
            BFT_kmer* su = get_successors(km, g);
            int same = 1;
            for (int i = 0; i < 4; i++) same &= is_kmer_in_cdbg(&pr[i]) == is_kmer_in_cdbg(&nb[i]) && is_kmer_in_cdbg(&su[i]) == is_kmer_in_cdbg(&nb[4 + i]);
            printf(" consistent=%d first_pred=%s\n", same, nb[0].kmer);
            free_BFT_kmer(nb, 8); free_BFT_kmer(pr, 4); free_BFT_kmer(su, 4);
            free(ids);
            free_BFT_annotation(a);
        }
        free_BFT_kmer(km, 1);
    }
    unset_neighbors_traversal(g);
    fclose(f);
    f = fopen(argv[3], "r");
    while (fgets(line, sizeof line, f)) {
        line[strcspn(line, "\r\n")] = 0;
        uint32_t* ids = query_sequence(g, line, thr, canonical);
        printf("S n=%u ids", ids[0]);
        for (uint32_t i = 1; i <= ids[0]; i++) printf(" %u", ids[i]);
        printf("\n");
        free(ids);
    }
    fclose(f);
    size_t n_iter = 0, n_colours = 0;
    iterate_over_kmers(g, count_and_check, &n_iter, &n_colours);
    printf("I kmers=%zu colours=%zu\n", n_iter, n_colours);
    /* prefix matching + set algebra on the first two k-mers of the query file that are present */
    {
        FILE* fq = fopen(argv[2], "r");
        BFT_annotation* found[2] = {NULL, NULL};
        int nf = 0;
        char pfx[16];
        while (nf < 2 && fgets(line, sizeof line, fq)) {
            line[strcspn(line, "\r\n")] = 0;
            BFT_kmer* km = get_kmer(line, g);
            if (is_kmer_in_cdbg(km)) {
                if (nf == 0) { memcpy(pfx, line, 5); pfx[5] = 0; }
                found[nf++] = get_annotation(km);
            }
            free_BFT_kmer(km, 1);
        }
        fclose(fq);
        if (nf == 2) {
            size_t pn = 0, pc = 0;
            bool any = prefix_matching(g, pfx, count_and_check, &pn, &pc);
            BFT_annotation* ai = intersection_annotations(g, 2, found[0], found[1]);
            BFT_annotation* au = union_annotations(g, 2, found[0], found[1]);
            BFT_annotation* ax = sym_difference_annotations(g, 2, found[0], found[1]);
            uint32_t* la = get_list_id_genomes(found[0], g);
            uint32_t* lb = get_list_id_genomes(found[1], g);
            uint32_t* li = intersection_list_id_genomes(la, lb);
            printf("P prefix=%s any=%d n=%zu | A inter=%u union=%u xor=%u a=%u b=%u listinter=%u\n", pfx, (int)any, pn,
                   get_count_id_genomes(ai, g), get_count_id_genomes(au, g), get_count_id_genomes(ax, g), la[0], lb[0], li[0]);
            free(la); free(lb); free(li);
            free_BFT_annotation(ai); free_BFT_annotation(au); free_BFT_annotation(ax);
            free_BFT_annotation(found[0]); free_BFT_annotation(found[1]);
        }
    }
    free_cdbg(g);
    return 0;
}
