/* bft_io.h — query-file readers and CSV writers of the file-level drivers
 * (reference src/file_io.c:651-895, 897-1020, 1464-1574). Pure host I/O; no query logic lives here. */
#ifndef BFT_IO_H
#define BFT_IO_H

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ASCII -> 2-bit packed k-mer, the reference's parseKmerCount (src/fasta.c:3-53): A/a=0 C/c=1 G/g=2 T/t/U/u=3,
 * nucleotide i in byte i/4 at bits 2*(i%4). Returns 1, or 0 if one of the first k characters is anything else. */
int bft_parse_kmer(const char* s, int k, uint64_t* words, int W);

/* Read a k-mer query file into packed words (n * W uint64, malloc'd).
 * binary != 0: "kmers_comp" layout — two text header lines, then ceil(2k/8)-byte records (src/file_io.c:721-727).
 * binary == 0: one ASCII k-mer per line; lines that fail bft_parse_kmer are dropped, as the reference drops them
 * (src/file_io.c:786-862). Returns 0 on success. */
int bft_read_kmer_file(const char* path, int binary, int k, int W, uint64_t** words, size_t* n);

/* Read a sequence file into concatenated characters + offsets (n+1 entries). Three layouts, told apart by the first
 * character of the file:
 *   - anything but '>' ';' '@': one sequence per line, CR/LF stripped — the only layout the reference reads
 *     (src/file_io.c:1519-1521);
 *   - '>' (or a ';' comment): FASTA — header lines dropped, the lines of a record joined into one sequence, so a FASTA
 *     file is answered exactly like its line-per-sequence flattening (one CSV row per record, empty records included);
 *   - '@': FASTQ — second line of every 4-line record.
 * A valid line-per-sequence file never starts with one of these (the reference exit(1)s on such a line). Returns 0 on success. */
int bft_read_sequence_file(const char* path, char** chars, uint64_t** offs, size_t* n);

/* Text k-mer file ("kmers", one ASCII k-mer per line): the first k characters of every line that has at least k before
 * its CR/LF/NUL, back to back (n * k chars, malloc'd) — the input of bft_b200_query_kmers_ascii, which does
 * parseKmerCount on the GPU and flags the k-mers it rejects. Returns 0 on success. */
int bft_read_kmer_text_file(const char* path, int k, char** ascii, size_t* n);

/* CSV output shared by -query_kmers and -query_sequences: header = genome names joined by ',' + '\n'
 * (src/file_io.c:706-719); one row per query of "0,1,...\n" (2*G bytes); finish() overwrites the last byte
 * written with '\0' (src/file_io.c:873-876, 1554-1557). rows = n * row_words bitmap words. */
int bft_csv_write_header(FILE* f, char* const* names, int n_genomes);
int bft_csv_write_rows(FILE* f, const uint32_t* rows, size_t n, int n_genomes, int row_words);
int bft_csv_finish(FILE* f);

#ifdef __cplusplus
}
#endif
#endif
