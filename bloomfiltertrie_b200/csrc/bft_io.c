/* bft_io.c — see bft_io.h. */
#define _GNU_SOURCE
#include "bft_io.h"

#include <stdlib.h>
#include <string.h>
#include <sys/types.h>

int bft_parse_kmer(const char* s, int k, uint64_t* words, int W) {
    for (int w = 0; w < W; w++) words[w] = 0;
    for (int i = 0; i < k; i++) {
        uint64_t code;
        switch (s[i]) {
            case 'a': case 'A': code = 0; break;
            case 'c': case 'C': code = 1; break;
            case 'g': case 'G': code = 2; break;
            case 't': case 'T': case 'u': case 'U': code = 3; break;
            default: return 0;
        }
        words[i >> 5] |= code << (2 * (i & 31));
    }
    return 1;
}

int bft_read_kmer_file(const char* path, int binary, int k, int W, uint64_t** words, size_t* n) {
    *words = NULL;
    *n = 0;
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    size_t cap = 1 << 16, cnt = 0;
    uint64_t* out = (uint64_t*)malloc(cap * (size_t)W * sizeof(uint64_t));
    if (!out) { fclose(f); return -1; }
    if (binary) {
        char line[128];
        if (!fgets(line, 100, f) || !fgets(line, 100, f)) { free(out); fclose(f); return -2; }
        const size_t nb = (size_t)(2 * k + 7) / 8;
        uint8_t rec[32];
        while (fread(rec, 1, nb, f) == nb) {
            if (cnt == cap) {
                cap *= 2;
                uint64_t* t = (uint64_t*)realloc(out, cap * (size_t)W * sizeof(uint64_t));
                if (!t) { free(out); fclose(f); return -1; }
                out = t;
            }
            uint8_t b[32];
            memset(b, 0, sizeof(b));
            memcpy(b, rec, nb);
            if ((2 * k) & 7) b[nb - 1] &= (uint8_t)((1u << ((2 * k) & 7)) - 1u); /* pad bits of the last byte carry no nucleotide */
            memcpy(out + cnt * (size_t)W, b, (size_t)W * 8);
            cnt++;
        }
    } else {
        char* line = NULL;
        size_t lcap = 0;
        ssize_t len;
        while ((len = getline(&line, &lcap, f)) != -1) {
            line[strcspn(line, "\r\n")] = '\0';
            if ((int)strlen(line) < k) continue; /* parseKmerCount hits '\0' -> line dropped */
            if (cnt == cap) {
                cap *= 2;
                uint64_t* t = (uint64_t*)realloc(out, cap * (size_t)W * sizeof(uint64_t));
                if (!t) { free(out); free(line); fclose(f); return -1; }
                out = t;
            }
            if (bft_parse_kmer(line, k, out + cnt * (size_t)W, W)) cnt++;
        }
        free(line);
    }
    fclose(f);
    *words = out;
    *n = cnt;
    return 0;
}

/* whole file into memory (+1 NUL) */
static char* slurp(const char* path, size_t* len) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    size_t cap = 1 << 20, n = 0;
    char* b = (char*)malloc(cap + 1);
    while (b) {
        const size_t got = fread(b + n, 1, cap - n, f);
        n += got;
        if (got == 0) break;
        if (n == cap) {
            cap *= 2;
            char* t = (char*)realloc(b, cap + 1);
            if (!t) { free(b); b = NULL; }
            else b = t;
        }
    }
    fclose(f);
    if (!b) return NULL;
    b[n] = 0;
    *len = n;
    return b;
}

int bft_read_sequence_file(const char* path, char** chars, uint64_t** offs, size_t* n) {
    *chars = NULL;
    *offs = NULL;
    *n = 0;
    size_t len = 0;
    char* buf = slurp(path, &len);
    if (!buf) return -1;
    /* format: FASTA if the first line starts with '>' (or ';'), FASTQ if it starts with '@', else one sequence per line */
    const int fasta = len && (buf[0] == '>' || buf[0] == ';');
    const int fastq = len && buf[0] == '@';
    size_t ocap = 1 << 12, cnt = 0, w = 0; /* the sequences are compacted in place: w <= read position always */
    uint64_t* ob = (uint64_t*)malloc(ocap * sizeof(uint64_t));
    if (!ob) { free(buf); return -1; }
    ob[0] = 0;
    int open_record = 0;   /* FASTA: a header was seen and its sequence is being collected */
    size_t line_no = 0;    /* FASTQ: position inside the 4-line record */
    for (size_t pos = 0; pos < len;) {
        const char* nl = (const char*)memchr(buf + pos, '\n', len - pos);
        const size_t end = nl ? (size_t)(nl - buf) : len;
        size_t l = end - pos;
        { /* the reference cuts a line at the first CR or LF (strcspn, src/file_io.c:1519-1521) */
            const char* cr = (const char*)memchr(buf + pos, '\r', l);
            if (cr) l = (size_t)(cr - (buf + pos));
        }
        int close_record = 0, take = 0;
        if (fasta) {
            if (l && (buf[pos] == '>' || buf[pos] == ';')) {
                if (buf[pos] == '>') { close_record = open_record; open_record = 1; }
            } else take = open_record;
        } else if (fastq) {
            take = (line_no & 3) == 1;
            close_record = (line_no & 3) == 3;
            line_no++;
        } else {
            take = 1;
            close_record = 1;
        }
        if (close_record && !take) { /* FASTA header closing the previous record / FASTQ quality line */
            cnt++;
            ob[cnt] = w;
        }
        if (take) {
            memmove(buf + w, buf + pos, l);
            w += l;
            if (close_record) { cnt++; ob[cnt] = w; }
        }
        if (cnt + 2 > ocap) {
            ocap *= 2;
            uint64_t* t = (uint64_t*)realloc(ob, ocap * sizeof(uint64_t));
            if (!t) { free(buf); free(ob); return -1; }
            ob = t;
        }
        pos = end + 1;
    }
    if (fasta && open_record) { cnt++; ob[cnt] = w; }
    *chars = buf;
    *offs = ob;
    *n = cnt;
    return 0;
}

int bft_read_kmer_text_file(const char* path, int k, char** ascii, size_t* n) {
    *ascii = NULL;
    *n = 0;
    size_t len = 0;
    char* buf = slurp(path, &len);
    if (!buf) return -1;
    size_t w = 0, cnt = 0;
    for (size_t pos = 0; pos < len;) {
        const char* nl = (const char*)memchr(buf + pos, '\n', len - pos);
        const size_t end = nl ? (size_t)(nl - buf) : len;
        size_t l = end - pos;
        const char* cr = (const char*)memchr(buf + pos, '\r', l);
        if (cr) l = (size_t)(cr - (buf + pos));
        const char* z = (const char*)memchr(buf + pos, 0, l); /* parseKmerCount stops at a NUL */
        if (z) l = (size_t)(z - (buf + pos));
        if (l >= (size_t)k) { /* a shorter line fails parseKmerCount and is dropped (src/file_io.c:786-862) */
            memmove(buf + w, buf + pos, (size_t)k);
            w += (size_t)k;
            cnt++;
        }
        pos = end + 1;
    }
    *ascii = buf;
    *n = cnt;
    return 0;
}

int bft_csv_write_header(FILE* f, char* const* names, int n_genomes) {
    for (int i = 0; i < n_genomes; i++) {
        if (fwrite(names[i], 1, strlen(names[i]), f) != strlen(names[i])) return -1;
        if (fputc(i + 1 < n_genomes ? ',' : '\n', f) == EOF) return -1;
    }
    return 0;
}

int bft_csv_write_rows(FILE* f, const uint32_t* rows, size_t n, int n_genomes, int row_words) {
    const size_t rowlen = (size_t)n_genomes * 2;
    const size_t chunk = 4096;
    char* buf = (char*)malloc(rowlen * chunk + 1);
    if (!buf) return -1;
    for (size_t base = 0; base < n; base += chunk) {
        const size_t m = n - base < chunk ? n - base : chunk;
        char* p = buf;
        for (size_t i = 0; i < m; i++) {
            const uint32_t* row = rows + (base + i) * (size_t)row_words;
            for (int g = 0; g < n_genomes; g++) {
                *p++ = (char)('0' + ((row[g >> 5] >> (g & 31)) & 1u));
                *p++ = g + 1 < n_genomes ? ',' : '\n';
            }
        }
        if (fwrite(buf, 1, rowlen * m, f) != rowlen * m) { free(buf); return -1; }
    }
    free(buf);
    return 0;
}

int bft_csv_finish(FILE* f) {
    const char eol = '\0';
    if (fseek(f, -1L, SEEK_CUR) != 0) return -1;
    return fwrite(&eol, 1, 1, f) == 1 ? 0 : -1;
}
