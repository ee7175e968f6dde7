/* oracle_cli — TEST INFRASTRUCTURE. Command-line front end of the plain-C restatement (bft_oracle.c); argument
 * and output conventions are those of oracle/ref_harness.c so tests can swap one for the other:
 *   oracle_cli kmers     file.bft queries.kc out.bin        -> present u8[n], rows u32[n][RW]
 *   oracle_cli branching file.bft queries.kc out.bin        -> succ u8[n], pred u8[n]
 *   oracle_cli sequences file.bft seqs.txt thr {canonical|non_canonical} out.bin -> rows u32[n][RW] */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "bft_oracle.h"

static double now_s(void) { struct timeval tv; gettimeofday(&tv, NULL); return tv.tv_sec + tv.tv_usec * 1e-6; }

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: see oracle/oracle_cli.c\n"); return 2; }
    char err[256];
    o_bft* b = o_load(argv[2], err, sizeof err);
    if (!b) { fprintf(stderr, "%s\n", err); return 1; }
    const int k = o_k(b), G = o_n_genomes(b), rw = (G + 31) / 32 > 0 ? (G + 31) / 32 : 1, nb = (2 * k + 7) / 8;
    uint32_t* ids = malloc(((size_t)G + 2) * sizeof(uint32_t));
    const int is_seq = strcmp(argv[1], "sequences") == 0;
    FILE* fo = fopen(argv[is_seq ? 6 : 4], "wb");
    if (!fo) { fprintf(stderr, "cannot write output\n"); return 1; }
    size_t n = 0;
    double t0 = now_s();
    if (!is_seq) {
        FILE* f = fopen(argv[3], "rb");
        char line[128];
        if (!f || !fgets(line, 100, f) || !fgets(line, 100, f)) { fprintf(stderr, "cannot read %s\n", argv[3]); return 1; }
        long start = ftell(f);
        fseek(f, 0, SEEK_END);
        n = (size_t)(ftell(f) - start) / (size_t)nb;
        fseek(f, start, SEEK_SET);
        uint8_t* q = malloc(n * (size_t)nb + 1);
        if (fread(q, (size_t)nb, n, f) != n) return 1;
        fclose(f);
        uint8_t* a = calloc(n + 1, 1);
        uint8_t* c = calloc(n + 1, 1);
        uint32_t* rows = calloc((n + 1) * (size_t)rw, 4);
        t0 = now_s();
        if (strcmp(argv[1], "kmers") == 0) {
            for (size_t i = 0; i < n; i++) {
                a[i] = (uint8_t)o_query_kmer(b, q + i * (size_t)nb, ids);
                for (uint32_t j = 1; j <= ids[0]; j++) rows[i * (size_t)rw + (ids[j] >> 5)] |= 1u << (ids[j] & 31);
            }
            fwrite(a, 1, n, fo);
            fwrite(rows, 4, n * (size_t)rw, fo);
        } else {
            for (size_t i = 0; i < n; i++) {
                a[i] = (uint8_t)o_branching_right(b, q + i * (size_t)nb);
                c[i] = (uint8_t)o_branching_left(b, q + i * (size_t)nb);
            }
            fwrite(a, 1, n, fo);
            fwrite(c, 1, n, fo);
        }
    } else {
        const double thr = atof(argv[4]);
        const int canonical = strcmp(argv[5], "canonical") == 0;
        FILE* f = fopen(argv[3], "r");
        if (!f) { fprintf(stderr, "cannot read %s\n", argv[3]); return 1; }
        char* line = NULL;
        size_t cap = 0;
        uint32_t* row = malloc((size_t)rw * 4);
        while (getline(&line, &cap, f) != -1) {
            line[strcspn(line, "\r\n")] = 0;
            memset(row, 0, (size_t)rw * 4);
            if (o_query_sequence(b, line, thr, canonical, ids)) { fprintf(stderr, "bad character in sequence %zu\n", n); return 3; }
            for (uint32_t j = 1; j <= ids[0]; j++) row[ids[j] >> 5] |= 1u << (ids[j] & 31);
            fwrite(row, 4, (size_t)rw, fo);
            n++;
        }
        fclose(f);
    }
    fclose(fo);
    double dt = now_s() - t0;
    printf("REF_PASS 0 seconds=%.6f\nORACLE mode=%s n=%zu k=%d genomes=%d threads=1 seconds=%.6f per_sec=%.1f\n", dt, argv[1], n, k, G, dt, n / dt);
    o_free(b);
    return 0;
}
