/* ref_harness — TEST / BASELINE INFRASTRUCTURE (not product code).
 *
 * Drives the UNMODIFIED reference library (compiled from /root/reference by oracle/Makefile into oracle/_ref/)
 * through its own public API, in memory, so that (a) parity tests get the reference's answers as arrays instead of
 * CSV text and (b) bench.py can time the reference's CPU query path on all host cores. Parallelism follows the
 * reference's own idiom for concurrent readers: one copy_BFT_Root per thread (include/CC.h:292-305) under OpenMP.
 *
 *   ref_harness kmers     file.bft queries.kc out.bin [threads] [repeat]
 *       per query: isKmerPresent + get_annotation + get_list_id_genomes (src/file_io.c:732-752)
 *       out.bin: n * (1 + 4*RW) bytes: present u8[n] then rows u32[n][RW]
 *   ref_harness branching file.bft queries.kc out.bin [threads] [repeat]
 *       per query: isBranchingRight, isBranchingLeft (src/file_io.c:943-946); out.bin: succ u8[n], pred u8[n]
 *   ref_harness sequences file.bft seqs.txt threshold {canonical|non_canonical} out.bin [threads] [repeat]
 *       per line: query_sequence (src/bft.c:1241); out.bin: rows u32[n][RW]
 * Prints "REF_PASS i seconds=..." per pass and one line "REF mode=... n=... threads=... seconds=... per_sec=..."
 * (best of `repeat` passes).
 */
#define _GNU_SOURCE
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "bft.h"

static double now_s(void) {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + tv.tv_usec * 1e-6;
}

static void ids_to_row(const uint32_t* ids, uint32_t* row) {
    for (uint32_t i = 1; i <= ids[0]; i++) row[ids[i] >> 5] |= 1u << (ids[i] & 31);
}

int main(int argc, char** argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: see header of oracle/ref_harness.c\n");
        return 2;
    }
    const char* mode = argv[1];
    BFT* bft = load_BFT(argv[2]);
    const int k = bft->k, G = bft->nb_genomes, rw = (G + 31) / 32 > 0 ? (G + 31) / 32 : 1;
    const int nb = CEIL(k * 2, SIZE_BITS_UINT_8T), lvl_root = k / NB_CHAR_SUF_PREF - 1;
    int is_seq = strcmp(mode, "sequences") == 0;
    int argi = is_seq ? 7 : 5;
    const char* out_path = argv[is_seq ? 6 : 4];
    int threads = argc > argi ? atoi(argv[argi]) : omp_get_max_threads();
    int repeat = argc > argi + 1 ? atoi(argv[argi + 1]) : 1;
    if (threads < 1) threads = 1;
    omp_set_num_threads(threads);

    BFT** copies = malloc(sizeof(BFT*) * threads);
    if (strcmp(mode, "branching") == 0) bft->skip_sp = build_skip_nodes(&(bft->node));
    for (int t = 0; t < threads; t++) copies[t] = copy_BFT_Root(bft);

    size_t n = 0;
    double best = 1e300;
    FILE* fo = fopen(out_path, "wb");
    if (!fo) { fprintf(stderr, "cannot write %s\n", out_path); return 1; }

    if (!is_seq) {
        FILE* f = fopen(argv[3], "rb");
        if (!f) { fprintf(stderr, "cannot read %s\n", argv[3]); return 1; }
        char line[128];
        if (!fgets(line, 100, f) || !fgets(line, 100, f)) return 1;
        long start = ftell(f);
        fseek(f, 0, SEEK_END);
        n = (size_t)(ftell(f) - start) / (size_t)nb;
        fseek(f, start, SEEK_SET);
        uint8_t* q = malloc(n * (size_t)nb + 1);
        if (fread(q, (size_t)nb, n, f) != n) return 1;
        fclose(f);
        if (strcmp(mode, "kmers") == 0) {
            uint8_t* present = calloc(n + 1, 1);
            uint32_t* rows = calloc((n + 1) * (size_t)rw, 4);
            for (int rep = 0; rep < repeat; rep++) {
                memset(rows, 0, n * (size_t)rw * 4);
                double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 4096)
                for (size_t i = 0; i < n; i++) {
                    BFT* b = copies[omp_get_thread_num()];
                    BFT_kmer km;
                    km.kmer = NULL;
                    km.kmer_comp = &q[i * (size_t)nb];
                    km.res = isKmerPresent(&(b->node), b, lvl_root, km.kmer_comp, k);
                    present[i] = is_kmer_in_cdbg(&km);
                    if (present[i]) {
                        BFT_annotation* a = get_annotation(&km);
                        uint32_t* ids = get_list_id_genomes(a, b);
                        free_BFT_annotation(a);
                        ids_to_row(ids, rows + i * (size_t)rw);
                        free(ids);
                    }
                    free(km.res);
                }
                double dt = now_s() - t0;
                printf("REF_PASS %d seconds=%.6f\n", rep, dt);
                if (dt < best) best = dt;
            }
            fwrite(present, 1, n, fo);
            fwrite(rows, 4, n * (size_t)rw, fo);
        } else if (strcmp(mode, "branching") == 0) {
            uint8_t* succ = calloc(n + 1, 1);
            uint8_t* pred = calloc(n + 1, 1);
            for (int rep = 0; rep < repeat; rep++) {
                double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 4096)
                for (size_t i = 0; i < n; i++) {
                    BFT* b = copies[omp_get_thread_num()];
                    succ[i] = (uint8_t)isBranchingRight(&(b->node), b, lvl_root, &q[i * (size_t)nb], k);
                    pred[i] = (uint8_t)isBranchingLeft(&(b->node), b, lvl_root, &q[i * (size_t)nb], k);
                }
                double dt = now_s() - t0;
                printf("REF_PASS %d seconds=%.6f\n", rep, dt);
                if (dt < best) best = dt;
            }
            fwrite(succ, 1, n, fo);
            fwrite(pred, 1, n, fo);
        } else {
            fprintf(stderr, "unknown mode %s\n", mode);
            return 2;
        }
    } else {
        double thr = atof(argv[4]);
        bool canonical = strcmp(argv[5], "canonical") == 0;
        FILE* f = fopen(argv[3], "r");
        if (!f) { fprintf(stderr, "cannot read %s\n", argv[3]); return 1; }
        size_t cap = 1024;
        char** seqs = malloc(cap * sizeof(char*));
        char* line = NULL;
        size_t lcap = 0;
        while (getline(&line, &lcap, f) != -1) {
            line[strcspn(line, "\r\n")] = '\0';
            if (n == cap) { cap *= 2; seqs = realloc(seqs, cap * sizeof(char*)); }
            seqs[n++] = strdup(line);
        }
        fclose(f);
        uint32_t* rows = calloc((n + 1) * (size_t)rw, 4);
        for (int rep = 0; rep < repeat; rep++) {
            memset(rows, 0, n * (size_t)rw * 4);
            double t0 = now_s();
#pragma omp parallel for schedule(dynamic, 256)
            for (size_t i = 0; i < n; i++) {
                BFT* b = copies[omp_get_thread_num()];
                uint32_t* ids = query_sequence(b, seqs[i], thr, canonical);
                ids_to_row(ids, rows + i * (size_t)rw);
                free(ids);
            }
            double dt = now_s() - t0;
            printf("REF_PASS %d seconds=%.6f\n", rep, dt);
            if (dt < best) best = dt;
        }
        fwrite(rows, 4, n * (size_t)rw, fo);
    }
    fclose(fo);
    printf("REF mode=%s n=%zu k=%d genomes=%d threads=%d seconds=%.6f per_sec=%.1f\n", mode, n, k, G, threads, best, (double)n / best);
    return 0;
}
