/* snippets_demo — TEST TOOL: one caller of the reference's traversal / marking API (include/bft.h, include/snippets.h),
 * compiled twice: against the unmodified reference (oracle/Makefile -> oracle/_ref/snippets_demo_ref) and, with
 * -DUSE_COMPAT, against include/bft_compat.h + libbft_b200.so. tests/test_gpu_graph.py compares what the two print and
 * write.
 *   snippets_demo file.bft queries.txt outdir [core_ratio]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef USE_COMPAT
#include "bft_compat.h"
#else
#include "bft.h"
#include "snippets.h"
#endif

static uint8_t flag_of(const char* kmer, int k) { /* a mark that depends on the k-mer only */
    unsigned s = 0;
    for (int i = 0; i < k; i++) s = s * 5u + (unsigned char)kmer[i];
    return (uint8_t)(s & 3u);
}

static size_t mark_all(BFT_kmer* km, BFT* g, va_list args) {
    size_t* n = va_arg(args, size_t*);
    set_flag_kmer(flag_of(km->kmer, g->k), km, g);
    (*n)++;
    return 1;
}

static size_t check_all(BFT_kmer* km, BFT* g, va_list args) {
    size_t* bad = va_arg(args, size_t*);
    size_t* hist = va_arg(args, size_t*);
    const uint8_t f = get_flag_kmer(km, g);
    if (f != flag_of(km->kmer, g->k)) (*bad)++;
    hist[f & 3]++;
    return 1;
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    BFT* g = load_BFT(argv[1]);
    const double ratio = argc > 4 ? atof(argv[4]) : 0.5;
    char path[4096];

    int n_bfs = 0, n_dfs = 0;
    get_nb_connected_component(g, &n_bfs, BFS);
    get_nb_connected_component(g, &n_dfs, DFS);
    printf("C bfs=%d dfs=%d\n", n_bfs, n_dfs);
    cdbg_traversal(g, BFS);
    printf("T done\n");

    snprintf(path, sizeof path, "%s/core.txt", argv[3]);
    extract_pangenome_kmers_to_disk(g, path, extract_core_kmers);
    snprintf(path, sizeof path, "%s/dispensable.txt", argv[3]);
    extract_pangenome_kmers_to_disk(g, path, extract_dispensable_kmers);
    snprintf(path, sizeof path, "%s/singleton.txt", argv[3]);
    extract_pangenome_kmers_to_disk(g, path, extract_singleton_kmers);

    snprintf(path, sizeof path, "%s/paths_core0.txt", argv[3]);
    extract_simple_core_paths_to_disk(g, 0.0, path);
    if (ratio > 0) {
        snprintf(path, sizeof path, "%s/paths_core.txt", argv[3]);
        extract_simple_core_paths_to_disk(g, ratio, path);
    }

    /* marking: flag every k-mer, read the flags back through the iterator and through get_kmer */
    size_t n_set = 0, bad = 0, hist[4] = {0, 0, 0, 0}, q_ok = 0, q_present = 0;
    set_marking(g);
    iterate_over_kmers(g, mark_all, &n_set);
    iterate_over_kmers(g, check_all, &bad, hist);
    FILE* f = fopen(argv[2], "r");
    char line[1024];
    while (f && fgets(line, sizeof line, f)) {
        line[strcspn(line, "\r\n")] = 0;
        BFT_kmer* km = get_kmer(line, g);
        if (is_kmer_in_cdbg(km)) {
            q_present++;
            q_ok += get_flag_kmer(km, g) == flag_of(line, g->k);
            set_flag_kmer(0, km, g);
            q_ok += get_flag_kmer(km, g) == 0;
        }
        free_BFT_kmer(km, 1);
    }
    if (f) fclose(f);
    unset_marking(g);
    printf("M set=%zu bad=%zu hist=%zu,%zu,%zu,%zu q_present=%zu q_ok=%zu\n", n_set, bad, hist[0], hist[1], hist[2], hist[3], q_present, q_ok);
    free_cdbg(g);
    return 0;
}
