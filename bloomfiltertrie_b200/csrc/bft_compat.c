/* bft_compat.c — reference-named entry points over the GPU engine (see include/bft_compat.h). */
#define _GNU_SOURCE
#include "bft_compat.h"
#include "bft_b200.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define DIE(...) do { fprintf(stderr, __VA_ARGS__); exit(EXIT_FAILURE); } while (0)
#define ENGINE_OK(call, who) do { if ((call) != BFT_B200_OK) DIE("%s: %s\n", who, bft_b200_last_error()); } while (0)
#define NOT_NULL(p, who) do { if ((p) == NULL) DIE("%s: NULL pointer\n", who); } while (0)

static int nb_bytes(int k) { return (2 * k + 7) / 8; }

/* parseKmerCount (src/fasta.c:3-53) */
static int parse_kmer_bytes(const char* s, int k, uint8_t* out) {
    for (int i = 0; i < k; i++) {
        uint8_t code;
        switch (s[i]) {
            case 'a': case 'A': code = 0; break;
            case 'c': case 'C': code = 1; break;
            case 'g': case 'G': code = 2; break;
            case 't': case 'T': case 'u': case 'U': code = 3; break;
            default: return 0;
        }
        out[i / 4] |= (uint8_t)(code << (2 * (i % 4)));
    }
    return 1;
}

BFT* load_BFT(char* filename) {
    NOT_NULL(filename, "load_BFT()");
    BFT* bft = (BFT*)calloc(1, sizeof(BFT));
    NOT_NULL(bft, "load_BFT()");
    ENGINE_OK(bft_b200_open(filename, 0, &bft->engine), "load_BFT()");
    bft->k = bft_b200_k(bft->engine);
    bft->nb_genomes = bft_b200_n_genomes(bft->engine);
    bft->filenames = (char**)calloc((size_t)bft->nb_genomes + 1, sizeof(char*));
    for (int i = 0; i < bft->nb_genomes; i++) bft->filenames[i] = strdup(bft_b200_genome_name(bft->engine, i));
    return bft;
}

void free_cdbg(BFT* bft) {
    NOT_NULL(bft, "free_cdbg()");
    for (int i = 0; i < bft->nb_genomes; i++) free(bft->filenames[i]);
    free(bft->filenames);
    free(bft->marks);
    bft_b200_close(bft->engine);
    free(bft);
}

BFT_kmer* create_empty_kmer(void) {
    BFT_kmer* km = (BFT_kmer*)calloc(1, sizeof(BFT_kmer));
    NOT_NULL(km, "create_empty_kmer()");
    return km;
}

BFT_kmer* create_kmer(const char* kmer, int k) { /* src/bft.c:122-150 */
    NOT_NULL(kmer, "create_kmer()");
    BFT_kmer* km = create_empty_kmer();
    km->kmer = (char*)malloc((size_t)k + 1);
    km->kmer_comp = (uint8_t*)calloc((size_t)nb_bytes(k), 1);
    memcpy(km->kmer, kmer, (size_t)k);
    km->kmer[k] = '\0';
    if (!parse_kmer_bytes(km->kmer, k, km->kmer_comp)) DIE("create_kmer(): Unexpected character encountered in k-mer.\n");
    km->res = (resultPresence*)calloc(1, sizeof(resultPresence));
    km->res->class_id = 0xffffffffu;
    return km;
}

void free_BFT_kmer_content(BFT_kmer* km, int n) {
    NOT_NULL(km, "free_BFT_kmer_content()");
    for (int i = 0; i < n; i++) {
        free(km[i].kmer);
        free(km[i].kmer_comp);
        free(km[i].res);
    }
}

void free_BFT_kmer(BFT_kmer* km, int n) {
    free_BFT_kmer_content(km, n);
    free(km);
}

static void lookup_into(BFT* bft, BFT_kmer* kms, int n) {
    const int W = bft_b200_kmer_words(bft->engine), nb = nb_bytes(bft->k);
    uint64_t words[8 * 4];
    uint8_t present[8];
    uint32_t cls[8];
    memset(words, 0, sizeof words);
    for (int i = 0; i < n; i++) memcpy(&words[i * W], kms[i].kmer_comp, (size_t)nb);
    ENGINE_OK(bft_b200_query_kmers(bft->engine, words, (size_t)n, present, NULL, cls), "get_kmer()");
    for (int i = 0; i < n; i++) {
        kms[i].res = (resultPresence*)calloc(1, sizeof(resultPresence));
        kms[i].res->present = present[i];
        kms[i].res->class_id = cls[i];
        kms[i].res->bft = bft;
    }
}

BFT_kmer* get_kmer(const char* kmer, BFT* bft) { /* src/bft.c:216-240 */
    NOT_NULL(kmer, "get_kmer()");
    NOT_NULL(bft, "get_kmer()");
    BFT_kmer* km = (BFT_kmer*)calloc(1, sizeof(BFT_kmer));
    km->kmer = (char*)malloc((size_t)bft->k + 1);
    km->kmer_comp = (uint8_t*)calloc((size_t)nb_bytes(bft->k), 1);
    memcpy(km->kmer, kmer, (size_t)bft->k);
    km->kmer[bft->k] = '\0';
    if (!parse_kmer_bytes(km->kmer, bft->k, km->kmer_comp)) DIE("get_kmer(): Unexpected character encountered in k-mer.\n");
    lookup_into(bft, km, 1);
    return km;
}

bool is_kmer_in_cdbg(BFT_kmer* km) { return km->res != NULL && km->res->present != 0; }

BFT_annotation* create_BFT_annotation(void) {
    BFT_annotation* a = (BFT_annotation*)calloc(1, sizeof(BFT_annotation));
    NOT_NULL(a, "create_BFT_annotation()");
    a->class_id = 0xffffffffu;
    return a;
}

void free_BFT_annotation(BFT_annotation* a) {
    NOT_NULL(a, "free_BFT_annotation()");
    if (!a->from_BFT) free(a->row);
    free(a);
}

BFT_annotation* get_annotation(BFT_kmer* km) { /* src/bft.c:363-387 */
    NOT_NULL(km, "get_annotation()");
    if (!is_kmer_in_cdbg(km)) DIE("get_annotation(): k-mer is not present in the graph.\n");
    BFT_annotation* a = create_BFT_annotation();
    a->class_id = km->res->class_id;
    a->from_BFT = 1;
    return a;
}

static const uint32_t* class_row(BFT* bft, uint32_t cls, int* rw);

/* the colour bitmap behind an annotation: a class row of the engine, or the owned row of a computed annotation */
static const uint32_t* annot_row(BFT* bft, const BFT_annotation* a, int* rw) {
    if (!a->from_BFT && a->row) { *rw = bft_b200_row_words(bft->engine); return a->row; }
    return class_row(bft, a->class_id, rw);
}

static BFT_annotation* combine_annotations(BFT* bft, uint32_t nb, va_list args, int op, const char* who) { /* src/bft.c:421-613 */
    if (nb == 0) DIE("%s: no annotations given as parameters.\n", who);
    const int rw = bft_b200_row_words(bft->engine);
    BFT_annotation* out = create_BFT_annotation();
    out->row = (uint32_t*)calloc((size_t)rw, sizeof(uint32_t));
    NOT_NULL(out->row, who);
    for (uint32_t i = 0; i < nb; i++) {
        BFT_annotation* a = va_arg(args, BFT_annotation*);
        NOT_NULL(a, who);
        int rw2;
        const uint32_t* r = annot_row(bft, a, &rw2);
        for (int w = 0; w < rw; w++) {
            if (i == 0) out->row[w] = r[w];
            else if (op == 0) out->row[w] &= r[w];
            else if (op == 1) out->row[w] |= r[w];
            else out->row[w] ^= r[w];
        }
    }
    return out;
}

BFT_annotation* intersection_annotations(BFT* bft, uint32_t nb_annotations, ...) {
    va_list args;
    va_start(args, nb_annotations);
    BFT_annotation* r = combine_annotations(bft, nb_annotations, args, 0, "intersection_annotations()");
    va_end(args);
    return r;
}
BFT_annotation* union_annotations(BFT* bft, uint32_t nb_annotations, ...) {
    va_list args;
    va_start(args, nb_annotations);
    BFT_annotation* r = combine_annotations(bft, nb_annotations, args, 1, "union_annotations()");
    va_end(args);
    return r;
}
BFT_annotation* sym_difference_annotations(BFT* bft, uint32_t nb_annotations, ...) {
    va_list args;
    va_start(args, nb_annotations);
    BFT_annotation* r = combine_annotations(bft, nb_annotations, args, 2, "sym_difference_annotations()");
    va_end(args);
    return r;
}

uint32_t* intersection_list_id_genomes(uint32_t* list_a, uint32_t* list_b) { /* src/bft.c:656-690: sorted [count, ids...] lists */
    NOT_NULL(list_a, "intersection_list_id_genomes()");
    NOT_NULL(list_b, "intersection_list_id_genomes()");
    const uint32_t na = list_a[0], nb = list_b[0];
    uint32_t* out = (uint32_t*)malloc(((size_t)(na < nb ? na : nb) + 1) * sizeof(uint32_t));
    NOT_NULL(out, "intersection_list_id_genomes()");
    uint32_t i = 1, j = 1, n = 0;
    while (i <= na && j <= nb) {
        if (list_a[i] < list_b[j]) i++;
        else if (list_a[i] > list_b[j]) j++;
        else { out[++n] = list_a[i]; i++; j++; }
    }
    out[0] = n;
    return out;
}

static const uint32_t* class_row(BFT* bft, uint32_t cls, int* rw) {
    const uint32_t* rows;
    uint64_t n;
    ENGINE_OK(bft_b200_class_rows(bft->engine, &rows, &n), "get_list_id_genomes()");
    *rw = bft_b200_row_words(bft->engine);
    if (cls >= n) DIE("get_list_id_genomes(): annotation does not belong to this BFT\n");
    return rows + (size_t)cls * (size_t)*rw;
}

uint32_t* get_list_id_genomes(BFT_annotation* a, BFT* bft) { /* src/bft.c:622-641: [count, ids ascending] */
    NOT_NULL(a, "get_list_id_genomes()");
    NOT_NULL(bft, "get_list_id_genomes()");
    int rw;
    const uint32_t* row = annot_row(bft, a, &rw);
    uint32_t cnt = 0;
    for (int w = 0; w < rw; w++) cnt += (uint32_t)__builtin_popcount(row[w]);
    uint32_t* ids = (uint32_t*)malloc(((size_t)cnt + 1) * sizeof(uint32_t));
    NOT_NULL(ids, "get_list_id_genomes()");
    ids[0] = cnt;
    uint32_t j = 1;
    for (int g = 0; g < bft->nb_genomes; g++)
        if ((row[g >> 5] >> (g & 31)) & 1u) ids[j++] = (uint32_t)g;
    return ids;
}

uint32_t get_count_id_genomes(BFT_annotation* a, BFT* bft) { /* src/bft.c:648 */
    NOT_NULL(a, "get_count_id_genomes()");
    NOT_NULL(bft, "get_count_id_genomes()");
    if (!a->from_BFT && a->row) {
        uint32_t cnt = 0;
        for (int w = 0; w < bft_b200_row_words(bft->engine); w++) cnt += (uint32_t)__builtin_popcount(a->row[w]);
        return cnt;
    }
    const uint32_t* counts;
    uint64_t n;
    ENGINE_OK(bft_b200_class_counts(bft->engine, &counts, &n), "get_count_id_genomes()");
    if (a->class_id >= n) DIE("get_count_id_genomes(): annotation does not belong to this BFT\n");
    return counts[a->class_id];
}

bool presence_genome(uint32_t id_genome, BFT_annotation* a, BFT* bft) { /* src/bft.c:395-413 */
    NOT_NULL(a, "is_genome_present()");
    NOT_NULL(bft, "is_genome_present()");
    if (id_genome >= (uint32_t)bft->nb_genomes) return false;
    int rw;
    const uint32_t* row = annot_row(bft, a, &rw);
    return (row[id_genome >> 5] >> (id_genome & 31)) & 1u;
}

uint32_t* query_sequence(BFT* bft, char* sequence, double threshold, bool canonical_search) { /* src/bft.c:1241-1351 */
    NOT_NULL(bft, "query_sequence()");
    NOT_NULL(sequence, "query_sequence()");
    if (threshold <= 0) DIE("query_sequence(): the threshold must be superior to 0.\n");
    if (threshold > 1) DIE("query_sequence(): the threshold must be inferior or equal to 1.\n");
    const int rw = bft_b200_row_words(bft->engine);
    uint64_t offs[2] = {0, strlen(sequence)};
    uint32_t* row = (uint32_t*)calloc((size_t)rw, sizeof(uint32_t));
    uint8_t status = 0;
    if ((long long)offs[1] - bft->k + 1 < 0)
        printf("query_sequence(): query %s is too small and must be at least of length k.\n", sequence);
    ENGINE_OK(bft_b200_query_sequences(bft->engine, sequence, offs, 1, threshold, canonical_search, row, &status), "query_sequence()");
    if (status == BFT_B200_SEQ_BAD_CHAR) DIE("get_kmer(): Unexpected character encountered in k-mer.\n");
    uint32_t cnt = 0;
    for (int w = 0; w < rw; w++) cnt += (uint32_t)__builtin_popcount(row[w]);
    uint32_t* ids = (uint32_t*)malloc(((size_t)cnt + 1) * sizeof(uint32_t));
    ids[0] = cnt;
    uint32_t j = 1;
    for (int g = 0; g < bft->nb_genomes; g++)
        if ((row[g >> 5] >> (g & 31)) & 1u) ids[j++] = (uint32_t)g;
    free(row);
    return ids;
}

/* The reference builds a node-rank directory here (build_skip_nodes, src/CC.c:2297-2331); the flattened arena
 * already carries exact prefix sums, so locking is a no-op. */
void set_neighbors_traversal(BFT* bft) { NOT_NULL(bft, "set_neighbors_traversal()"); }
void unset_neighbors_traversal(BFT* bft) { NOT_NULL(bft, "unset_neighbors_traversal()"); }

static BFT_kmer* neighbours(BFT_kmer* km, BFT* bft, int first, int count, const char* who) { /* src/bft.c:804-1003 */
    NOT_NULL(km, who);
    NOT_NULL(bft, who);
    if (!is_kmer_in_cdbg(km)) DIE("%s: k-mer is not present in the graph.\n", who);
    static const char nuc[4] = {'A', 'C', 'G', 'T'};
    const int k = bft->k;
    BFT_kmer* out = (BFT_kmer*)calloc((size_t)count, sizeof(BFT_kmer));
    for (int i = 0; i < count; i++) {
        const int slot = first + i; /* 0-3 predecessors, 4-7 successors */
        out[i].kmer = (char*)malloc((size_t)k + 1);
        out[i].kmer_comp = (uint8_t*)calloc((size_t)nb_bytes(k), 1);
        if (slot < 4) {
            out[i].kmer[0] = nuc[slot];
            memcpy(out[i].kmer + 1, km->kmer, (size_t)k - 1);
        } else {
            memcpy(out[i].kmer, km->kmer + 1, (size_t)k - 1);
            out[i].kmer[k - 1] = nuc[slot - 4];
        }
        out[i].kmer[k] = '\0';
        parse_kmer_bytes(out[i].kmer, k, out[i].kmer_comp);
    }
    /* presence + class of the 8 neighbours in the reference's own order (and with its leaf-level successor rule) */
    uint64_t words[4] = {0, 0, 0, 0};
    uint32_t cls[8];
    memcpy(words, km->kmer_comp, (size_t)nb_bytes(k));
    ENGINE_OK(bft_b200_query_neighbors(bft->engine, words, 1, cls), who);
    for (int i = 0; i < count; i++) {
        out[i].res = (resultPresence*)calloc(1, sizeof(resultPresence));
        out[i].res->present = cls[first + i] != 0xffffffffu;
        out[i].res->class_id = cls[first + i];
        out[i].res->bft = bft;
    }
    return out;
}
BFT_kmer* get_neighbors(BFT_kmer* km, BFT* bft) { return neighbours(km, bft, 0, 8, "get_neighbors()"); }
BFT_kmer* get_predecessors(BFT_kmer* km, BFT* bft) { return neighbours(km, bft, 0, 4, "get_predecessors()"); }
BFT_kmer* get_successors(BFT_kmer* km, BFT* bft) { return neighbours(km, bft, 4, 4, "get_successors()"); }

/* shared by iteration and prefix matching: prefix == NULL iterates everything; returns the number of k-mers visited */
static size_t iterate_filtered(BFT* bft, const char* prefix, BFT_func_ptr f, va_list args);

void v_iterate_over_kmers(BFT* bft, BFT_func_ptr f, va_list args) { /* src/bft.c:1014-1034 */
    NOT_NULL(bft, "v_iterate_over_kmers()");
    iterate_filtered(bft, NULL, f, args);
}

static size_t iterate_filtered(BFT* bft, const char* prefix, BFT_func_ptr f, va_list args) {
    size_t matched = 0;
    uint64_t pmask[4] = {0, 0, 0, 0}, pval[4] = {0, 0, 0, 0};
    if (prefix) { /* compare the leading nucleotides in packed form */
        const int len = (int)strlen(prefix);
        uint8_t tmp[40];
        memset(tmp, 0, sizeof tmp);
        parse_kmer_bytes(prefix, len, tmp);
        memcpy(pval, tmp, 32);
        for (int j = 0; j < len; j++) pmask[j >> 5] |= 3ULL << (2 * (j & 31));
    }
    bft_b200_stats st;
    ENGINE_OK(bft_b200_get_stats(bft->engine, &st), "v_iterate_over_kmers()");
    const size_t n = (size_t)st.n_kmers, W = (size_t)bft_b200_kmer_words(bft->engine);
    const int k = bft->k, nb = nb_bytes(k);
    uint64_t* km = (uint64_t*)malloc((n + 1) * W * sizeof(uint64_t));
    uint32_t* cls = (uint32_t*)malloc((n + 1) * sizeof(uint32_t));
    NOT_NULL(km, "v_iterate_over_kmers()");
    NOT_NULL(cls, "v_iterate_over_kmers()");
    ENGINE_OK(bft_b200_extract_kmers(bft->engine, km, cls, NULL, n, NULL), "v_iterate_over_kmers()");
    BFT_kmer cur;
    resultPresence res;
    cur.kmer = (char*)malloc((size_t)k + 2);
    cur.kmer_comp = (uint8_t*)calloc((size_t)nb + 8, 1);
    cur.res = &res;
    for (size_t i = 0; i < n; i++) {
        if (prefix) {
            int match = 1;
            for (size_t w = 0; w < W; w++) match &= (km[i * W + w] & pmask[w]) == pval[w];
            if (!match) continue;
        }
        matched++;
        for (int j = 0; j < k; j++) cur.kmer[j] = "ACGT"[(km[i * W + (size_t)(j >> 5)] >> (2 * (j & 31))) & 3];
        cur.kmer[k] = '\0';
        memcpy(cur.kmer_comp, km + i * W, (size_t)nb);
        res.present = 1;
        res.class_id = cls[i];
        res.bft = bft;
        res.vertex1 = (uint32_t)i + 1;
        va_list copy;
        va_copy(copy, args);
        const size_t go_on = f(&cur, bft, copy);
        va_end(copy);
        if (go_on == 0) break;
    }
    free(cur.kmer);
    free(cur.kmer_comp);
    free(km);
    free(cls);
    return matched;
}

bool prefix_matching(BFT* bft, char* prefix, BFT_func_ptr f, ...) { /* src/bft.c:1093-1141 */
    NOT_NULL(bft, "prefix_matching()");
    NOT_NULL(prefix, "prefix_matching()");
    const int len = (int)strlen(prefix);
    if (len > bft->k) DIE("prefix_matching(): Prefix length is larger than k-mer length.\n");
    if (len == 0) DIE("prefix_matching(): Prefix length is 0.\n");
    uint8_t tmp[40];
    memset(tmp, 0, sizeof tmp);
    if (!parse_kmer_bytes(prefix, len, tmp)) DIE("prefix_matching(): Non-ACGT char. encountered in prefix.\n");
    va_list args;
    va_start(args, f);
    const size_t matched = iterate_filtered(bft, prefix, f, args);
    va_end(args);
    return matched > 0;
}

void iterate_over_kmers(BFT* bft, BFT_func_ptr f, ...) { /* src/bft.c:1043-1075 */
    va_list args;
    va_start(args, f);
    v_iterate_over_kmers(bft, f, args);
    va_end(args);
}

size_t write_kmer_ascii_to_disk(BFT_kmer* bft_kmer, BFT* bft, va_list args) { /* src/bft.c:291-300 */
    FILE* file = va_arg(args, FILE*);
    bft_kmer->kmer[bft->k] = '\n';
    fwrite(bft_kmer->kmer, sizeof(char), (size_t)bft->k + 1, file);
    return 1;
}

size_t write_kmer_comp_to_disk(BFT_kmer* bft_kmer, BFT* bft, va_list args) { /* src/bft.c:308-318 */
    (void)bft;
    int nb_bytes_kmer_comp = va_arg(args, int);
    FILE* file = va_arg(args, FILE*);
    fwrite(bft_kmer->kmer_comp, sizeof(uint8_t), (size_t)nb_bytes_kmer_comp, file);
    return 1;
}

void extract_kmers_to_disk(BFT* bft, char* filename_output, bool compressed_output) { /* src/bft.c:255-283 */
    NOT_NULL(bft, "extract_kmers_to_disk()");
    NOT_NULL(filename_output, "extract_kmers_to_disk()");
    ENGINE_OK(bft_b200_extract_kmers_file(bft->engine, filename_output, compressed_output), "extract_kmers_to_disk()");
}

int queryBFT_kmerPresences_from_KmerFiles(BFT_Root* root, char* query_filename, int binary_file, char* output_filename) {
    NOT_NULL(root, "queryBFT_kmerPresences_from_KmerFiles()");
    NOT_NULL(query_filename, "queryBFT_kmerPresences_from_KmerFiles()");
    uint64_t n = 0;
    printf("\nQuerying BFT for k-mers in %s\n\n", query_filename);
    ENGINE_OK(bft_b200_query_kmers_file(root->engine, query_filename, binary_file, output_filename, &n), "queryBFT_kmerPresences_from_KmerFiles()");
    return (int)n;
}

int queryBFT_kmerBranching_from_KmerFiles(BFT_Root* root, char* query_filename, int binary_file) {
    NOT_NULL(root, "queryBFT_kmerBranching_from_KmerFiles()");
    NOT_NULL(query_filename, "queryBFT_kmerBranching_from_KmerFiles()");
    uint64_t n = 0;
    printf("\nQuerying BFT for branching k-mers in %s\n\n", query_filename);
    ENGINE_OK(bft_b200_query_branching_file(root->engine, query_filename, binary_file, &n), "queryBFT_kmerBranching_from_KmerFiles()");
    return (int)n;
}

void query_sequences_outputCSV(BFT_Root* root, char* query_filename, char* output_filename, double threshold, bool canonical_search) {
    NOT_NULL(root, "query_sequences_outputCSV()");
    NOT_NULL(query_filename, "query_sequences_outputCSV()");
    NOT_NULL(output_filename, "query_sequences_outputCSV()");
    if (threshold <= 0) DIE("query_sequences_outputCSV(): the threshold must be superior to 0.\n");
    if (threshold > 1) DIE("query_sequences_outputCSV(): the threshold must be inferior or equal to 1.\n");
    ENGINE_OK(bft_b200_query_sequences_file(root->engine, query_filename, output_filename, threshold, canonical_search), "query_sequences_outputCSV()");
    printf("\nFile %s has been processed.\n", query_filename);
}

/* ---- marking (include/bft.h:143-146) -------------------------------------------------------------------------------
 * The reference appends 2-bit marks to the UCs of the trie (src/marking.c); here a k-mer's vertex index keys a host
 * array, so marks cost one byte per stored k-mer and nothing in device memory. */
void set_marking(BFT* bft) { /* src/bft.c:700-712 */
    NOT_NULL(bft, "set_marking()");
    bft_b200_stats st;
    ENGINE_OK(bft_b200_get_stats(bft->engine, &st), "set_marking()");
    free(bft->marks);
    bft->marks = (uint8_t*)calloc((size_t)st.n_kmers + 1, 1);
    NOT_NULL(bft->marks, "set_marking()");
}

void unset_marking(BFT* bft) { /* src/bft.c:717-729 */
    NOT_NULL(bft, "unset_marking()");
    free(bft->marks);
    bft->marks = NULL;
}

static uint32_t vertex_of(BFT_kmer* km, BFT* bft, const char* who) {
    NOT_NULL(km, who);
    NOT_NULL(bft, who);
    if (!bft->marks) DIE("%s: graph is not locked for marking (set_marking() must be called first).\n", who);
    if (!is_kmer_in_cdbg(km)) DIE("%s: k-mer is not present in the graph.\n", who);
    if (!km->res->vertex1) {
        uint64_t words[4] = {0, 0, 0, 0};
        uint32_t vid = 0xffffffffu;
        memcpy(words, km->kmer_comp, (size_t)nb_bytes(bft->k));
        ENGINE_OK(bft_b200_query_vertex_ids(bft->engine, words, 1, &vid), who);
        if (vid == 0xffffffffu) DIE("%s: k-mer is not present in the graph.\n", who);
        km->res->vertex1 = vid + 1;
    }
    return km->res->vertex1 - 1;
}

void set_flag_kmer(uint8_t flag, BFT_kmer* km, BFT* bft) { /* src/bft.c:737-763 */
    bft->marks[vertex_of(km, bft, "set_flag_kmer()")] = flag & 3u;
}

uint8_t get_flag_kmer(BFT_kmer* km, BFT* bft) { /* src/bft.c:770-796 */
    return bft->marks[vertex_of(km, bft, "get_flag_kmer()")];
}

/* ---- src/snippets.c ------------------------------------------------------------------------------------------------ */
static size_t write_if_count(BFT_kmer* kmer, BFT* graph, va_list args, int want) { /* want: 0 core, 1 dispensable, 2 singleton */
    FILE* file = va_arg(args, FILE*);
    int* nb_kmers = va_arg(args, int*);
    BFT_annotation* a = get_annotation(kmer);
    const uint32_t cnt = get_count_id_genomes(a, graph);
    free_BFT_annotation(a);
    const uint32_t all = (uint32_t)graph->nb_genomes;
    if ((want == 0 && cnt == all) || (want == 1 && cnt < all) || (want == 2 && cnt == 1)) {
        fwrite(kmer->kmer, sizeof(char), strlen(kmer->kmer) + 1, file); /* the k-mer and its NUL, as the reference writes it */
        *nb_kmers += 1;
    }
    return 1;
}
size_t extract_core_kmers(BFT_kmer* kmer, BFT* graph, va_list args) { return write_if_count(kmer, graph, args, 0); }        /* src/snippets.c:10-27 */
size_t extract_dispensable_kmers(BFT_kmer* kmer, BFT* graph, va_list args) { return write_if_count(kmer, graph, args, 1); } /* :35-52 */
size_t extract_singleton_kmers(BFT_kmer* kmer, BFT* graph, va_list args) { return write_if_count(kmer, graph, args, 2); }   /* :60-76 */

void extract_pangenome_kmers_to_disk(BFT* graph, char* filename_output, BFT_func_ptr f) { /* src/snippets.c:86-107 */
    NOT_NULL(graph, "extract_pangenome_kmers_to_disk()");
    NOT_NULL(filename_output, "extract_pangenome_kmers_to_disk()");
    FILE* file = fopen(filename_output, "w");
    if (!file) DIE("extract_pangenome_kmers_to_disk(): failed to create/open output file.\n");
    int nb_kmers = 0;
    iterate_over_kmers(graph, f, file, &nb_kmers);
    fclose(file);
    printf("Number of extracted k-mers is %d.\n", nb_kmers);
}

void extract_simple_paths_to_disk(BFT* graph, char* filename_output) { /* src/snippets.c:310-338 */
    NOT_NULL(graph, "extract_simple_paths_to_disk()");
    NOT_NULL(filename_output, "extract_simple_paths_to_disk()");
    uint64_t longest = 0;
    ENGINE_OK(bft_b200_simple_paths_file(graph->engine, 0.0, filename_output, NULL, &longest), "extract_simple_paths_to_disk()");
    printf("Longest simple path has %d nuc.\n", (int)longest);
}

void extract_simple_core_paths_to_disk(BFT* graph, double core_ratio, char* filename_output) { /* src/snippets.c:572-597 */
    NOT_NULL(graph, "extract_simple_core_paths_to_disk()");
    NOT_NULL(filename_output, "extract_simple_core_paths_to_disk()");
    uint64_t longest = 0;
    ENGINE_OK(bft_b200_simple_paths_file(graph->engine, core_ratio, filename_output, NULL, &longest), "extract_simple_core_paths_to_disk()");
    printf("Longest simple core path has %d nuc.\n", (int)longest);
}

/* The four traversal callbacks only name the traversal wanted; the traversal itself is a device kernel. */
static size_t traversal_tag(const char* who) {
    DIE("%s: per-k-mer traversal callbacks are not run on the host; pass this function to cdbg_traversal() or get_nb_connected_component().\n", who);
    return 0;
}
size_t BFS(BFT_kmer* kmer, BFT* graph, va_list args) { (void)kmer; (void)graph; (void)args; return traversal_tag("BFS()"); }
size_t DFS(BFT_kmer* kmer, BFT* graph, va_list args) { (void)kmer; (void)graph; (void)args; return traversal_tag("DFS()"); }
size_t BFS_subgraph(BFT_kmer* kmer, BFT* graph, va_list args) { (void)kmer; (void)graph; (void)args; return traversal_tag("BFS_subgraph()"); }
size_t DFS_subgraph(BFT_kmer* kmer, BFT* graph, va_list args) { (void)kmer; (void)graph; (void)args; return traversal_tag("DFS_subgraph()"); }

/* counts the components the traversal `f` (with its genome-id arguments still in args) would start */
static uint64_t run_traversal(BFT* graph, BFT_func_ptr f, va_list args, const char* who) {
    uint64_t n = 0;
    if (f == BFS || f == DFS) {
        ENGINE_OK(bft_b200_connected_components(graph->engine, NULL, 0, &n, NULL), who);
    } else if (f == BFS_subgraph || f == DFS_subgraph) {
        const int nb_id_genomes = va_arg(args, int);
        if (nb_id_genomes <= 0) return 0; /* is_in_subgraph() is false for every k-mer (src/snippets.c:848) */
        uint32_t* ids = (uint32_t*)malloc((size_t)nb_id_genomes * sizeof(uint32_t));
        NOT_NULL(ids, who);
        for (int i = 0; i < nb_id_genomes; i++) ids[i] = va_arg(args, uint32_t);
        ENGINE_OK(bft_b200_connected_components(graph->engine, ids, nb_id_genomes, &n, NULL), who);
        free(ids);
    } else {
        DIE("%s: only BFS, DFS, BFS_subgraph and DFS_subgraph run on the device; use iterate_over_kmers() for other callbacks.\n", who);
    }
    return n;
}

void cdbg_traversal(BFT* graph, BFT_func_ptr f, ...) { /* src/snippets.c:883-907: visits every k-mer; nothing is returned */
    NOT_NULL(graph, "cdbg_traversal()");
    va_list args;
    va_start(args, f);
    (void)run_traversal(graph, f, args, "cdbg_traversal()");
    va_end(args);
}

void get_nb_connected_component(BFT* graph, ...) { /* src/snippets.c:915-958: args = int* count, traversal, its arguments */
    NOT_NULL(graph, "get_nb_connected_component()");
    va_list args;
    va_start(args, graph);
    int* nb_connected_comp = va_arg(args, int*);
    BFT_func_ptr f = va_arg(args, BFT_func_ptr);
    NOT_NULL(nb_connected_comp, "get_nb_connected_component()");
    *nb_connected_comp += (int)run_traversal(graph, f, args, "get_nb_connected_component()");
    va_end(args);
}
