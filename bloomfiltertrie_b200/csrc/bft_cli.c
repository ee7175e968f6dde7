/* bft_cli.c — `bft_b200`: the query half of the reference CLI (src/main.c:204-316) on the GPU engine.
 *   bft_b200 load file_bft [-query_kmers {kmers|kmers_comp} list] [-query_sequences thr {canonical|non_canonical} list]
 *                          [-query_branching {kmers|kmers_comp} list] [-extract_kmers {kmers|kmers_comp} file]
 *                          [-connected_components] [-simple_paths core_ratio file]
 * (the last two expose the library's src/snippets.c traversals, which the reference binary has no option for)
 * Output files are named and placed as the reference does (basename with extension replaced by .csv in the cwd,
 * src/main.c:258-264). `build` / -add_genomes stay with the reference binary. */
#define _GNU_SOURCE
#include <libgen.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bft_compat.h"

static char* csv_name(char* path) {
    char* b = basename(path);
    char* out = (char*)malloc(strlen(b) + 5);
    strcpy(out, b);
    char* dot = strrchr(out, '.');
    if (dot) strcpy(dot, ".csv"); else strcat(out, ".csv");
    return out;
}

int main(int argc, char** argv) {
    if (argc == 2 && strcmp(argv[1], "--version") == 0) { printf("0.8-b200\n"); return 0; }
    if (argc < 3 || strcmp(argv[1], "load") != 0) {
        fprintf(stderr, "Usage:\nbft_b200 load file_bft [-query_sequences threshold {canonical|non_canonical} list_sequence_files]\n"
                        "                       [-query_kmers {kmers|kmers_comp} list_kmer_files]\n"
                        "                       [-query_branching {kmers|kmers_comp} list_kmer_files]\n"
                        "                       [-extract_kmers {kmers|kmers_comp} kmers_file]\n"
                        "                       [-connected_components] [-simple_paths core_ratio paths_file]\n"
                        "Graph construction (build, -add_genomes) is served by the reference `bft` binary.\n");
        return EXIT_FAILURE;
    }
    BFT* bft = load_BFT(argv[2]);
    char buffer[2048];
    for (int i = 3; i < argc;) {
        if (strcmp(argv[i], "-query_kmers") == 0 && i + 2 < argc) {
            int binary = strcmp(argv[i + 1], "kmers_comp") == 0;
            FILE* fl = fopen(argv[i + 2], "r");
            if (!fl) { fprintf(stderr, "Invalid k-mer queries files list.\n"); return EXIT_FAILURE; }
            while (fgets(buffer, sizeof buffer, fl)) {
                buffer[strcspn(buffer, "\r\n")] = 0;
                char* out = csv_name(buffer);
                printf("\nNb k-mers present = %d\n", queryBFT_kmerPresences_from_KmerFiles(bft, buffer, binary, out));
                free(out);
            }
            fclose(fl);
            i += 3;
        } else if (strcmp(argv[i], "-query_sequences") == 0 && i + 3 < argc) {
            double thr = atof(argv[i + 1]);
            if (thr == 0) { fprintf(stderr, "Could not parse threshold for command -query_sequences.\n"); return EXIT_FAILURE; }
            if (strcmp(argv[i + 2], "canonical") && strcmp(argv[i + 2], "non_canonical")) {
                fprintf(stderr, "Unrecognized type of k-mers to search for %s.\n", argv[i]);
                return EXIT_FAILURE;
            }
            bool canonical = strcmp(argv[i + 2], "canonical") == 0;
            FILE* fl = fopen(argv[i + 3], "r");
            if (!fl) { fprintf(stderr, "Invalid sequence query file list.\n"); return EXIT_FAILURE; }
            while (fgets(buffer, sizeof buffer, fl)) {
                buffer[strcspn(buffer, "\r\n")] = 0;
                char* out = csv_name(buffer);
                query_sequences_outputCSV(bft, buffer, out, thr, canonical);
                free(out);
            }
            fclose(fl);
            i += 4;
        } else if (strcmp(argv[i], "-query_branching") == 0 && i + 2 < argc) {
            int binary = strcmp(argv[i + 1], "kmers_comp") == 0;
            FILE* fl = fopen(argv[i + 2], "r");
            if (!fl) { fprintf(stderr, "Invalid branching k-mer queries files list.\n"); return EXIT_FAILURE; }
            while (fgets(buffer, sizeof buffer, fl)) {
                buffer[strcspn(buffer, "\r\n")] = 0;
                printf("\nNb branching k-mers = %d\n", queryBFT_kmerBranching_from_KmerFiles(bft, buffer, binary));
            }
            fclose(fl);
            i += 3;
        } else if (strcmp(argv[i], "-extract_kmers") == 0 && i + 2 < argc) { /* src/main.c:317, write_kmers_2disk */
            printf("\nExtraction of k-mers from the BFT to file %s\n\n", argv[i + 2]);
            extract_kmers_to_disk(bft, argv[i + 2], strcmp(argv[i + 1], "kmers_comp") == 0);
            i += 3;
        } else if (strcmp(argv[i], "-connected_components") == 0) { /* get_nb_connected_component(graph, &n, BFS) */
            int n = 0;
            get_nb_connected_component(bft, &n, BFS);
            printf("\nNb connected components = %d\n", n);
            i += 1;
        } else if (strcmp(argv[i], "-simple_paths") == 0 && i + 2 < argc) { /* extract_simple[_core]_paths_to_disk */
            const double ratio = atof(argv[i + 1]);
            if (ratio > 0) extract_simple_core_paths_to_disk(bft, ratio, argv[i + 2]);
            else extract_simple_paths_to_disk(bft, argv[i + 2]);
            i += 3;
        } else {
            fprintf(stderr, "Unrecognized command %s.\n", argv[i]);
            return EXIT_FAILURE;
        }
    }
    free_cdbg(bft);
    return EXIT_SUCCESS;
}
