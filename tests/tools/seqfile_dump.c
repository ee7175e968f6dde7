/* seqfile_dump — TEST TOOL: what bft_read_sequence_file / bft_read_kmer_text_file (csrc/bft_io.c) make of a file.
 *   seqfile_dump seq FILE      -> "<n>\n" then every sequence on its own line
 *   seqfile_dump kmers K FILE  -> "<n>\n" then every accepted k-mer line (first K characters) on its own line */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bft_io.h"

int main(int argc, char** argv) {
    if (argc == 3 && !strcmp(argv[1], "seq")) {
        char* chars; uint64_t* offs; size_t n;
        if (bft_read_sequence_file(argv[2], &chars, &offs, &n)) return 1;
        printf("%zu\n", n);
        for (size_t i = 0; i < n; i++) { fwrite(chars + offs[i], 1, (size_t)(offs[i + 1] - offs[i]), stdout); fputc('\n', stdout); }
        free(chars); free(offs);
        return 0;
    }
    if (argc == 4 && !strcmp(argv[1], "kmers")) {
        char* a; size_t n; const int k = atoi(argv[2]);
        if (bft_read_kmer_text_file(argv[3], k, &a, &n)) return 1;
        printf("%zu\n", n);
        for (size_t i = 0; i < n; i++) { fwrite(a + i * (size_t)k, 1, (size_t)k, stdout); fputc('\n', stdout); }
        free(a);
        return 0;
    }
    return 2;
}
