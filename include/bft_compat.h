/* bft_compat.h — source-level drop-in for the QUERY subset of the reference's public API (include/bft.h).
 *
 * Same names, argument meaning, ownership and error behaviour as the reference (fprintf(stderr)+exit(1) on the
 * conditions where the reference calls ERROR(), include/useful_macros.h:33-43), implemented as batch-of-one /
 * batch-of-eight calls into the GPU engine (include/bft_b200.h). Graph construction and mutation
 * (create_cdbg, insert_*, marking, iteration, set algebra, write_BFT) stay on the reference host library and
 * are not declared here. Do not include the reference's bft.h in the same translation unit.
 *
 * Reference declarations replaced (include/bft.h): load_BFT :176, free_cdbg :63, get_kmer :125,
 * is_kmer_in_cdbg :126, query_sequence :127, get_annotation :97, presence_genome :98, get_list_id_genomes :115,
 * get_count_id_genomes :116, free_BFT_annotation :96, create_kmer/free_BFT_kmer/free_BFT_kmer_content :78-82,
 * set/unset_neighbors_traversal :154-155, get_neighbors/get_predecessors/get_successors :156-158,
 * intersection/union/sym_difference_annotations :99-113, prefix_matching :137, iterate_over_kmers/v_iterate_over_kmers :164-165, extract_kmers_to_disk + write_kmer_{ascii,comp}_to_disk :88-90; and the
 * file-level drivers of include/file_io.h (queryBFT_kmerPresences_from_KmerFiles, queryBFT_kmerBranching_from_KmerFiles,
 * query_sequences_outputCSV); set_marking/unset_marking/set_flag_kmer/get_flag_kmer :143-146; and include/snippets.h:
 * extract_{core,dispensable,singleton}_kmers + extract_pangenome_kmers_to_disk :39-42, extract_simple_paths_to_disk :50,
 * extract_simple_core_paths_to_disk :52, BFS/DFS/BFS_subgraph/DFS_subgraph + cdbg_traversal +
 * get_nb_connected_component :60-67.
 */
#ifndef BFT_COMPAT_H
#define BFT_COMPAT_H

#include <stdarg.h>
#include <stdbool.h>
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

struct bft_b200_ctx;

/* BFT_Root (include/Node.h:96-122): the documented public fields, then the engine handle. */
typedef struct {
    char** filenames; /* inserted genome file names */
    int k;            /* size of k-mers */
    int nb_genomes;   /* number of genomes inserted */
    struct bft_b200_ctx* engine;
    uint8_t* marks;   /* set_marking(): one flag per stored k-mer */
} BFT_Root;
typedef BFT_Root BFT;

/* resultPresence (include/Node.h:61-89) holds raw pointers into the host trie in the reference; here it is the
 * device-side locator of the k-mer: presence + colour class. Treat as opaque, as reference users do. */
typedef struct {
    uint32_t present;
    uint32_t class_id;
    BFT* bft;
    uint32_t vertex1; /* 1 + index of the k-mer among the stored k-mers once marking looked it up; 0 = not yet */
} resultPresence;

typedef struct {
    char* kmer;          /* ASCII null-terminated k-mer */
    uint8_t* kmer_comp;  /* 2 bits encoded form */
    resultPresence* res;
} BFT_kmer;

typedef struct {
    uint8_t* annot;      /* not materialised on the host (the colour class below identifies the set) */
    uint8_t* annot_ext;
    uint8_t* annot_cplx;
    int size_annot;
    int size_annot_cplx;
    uint8_t from_BFT;
    uint32_t class_id;   /* from_BFT: the colour class of the k-mer */
    uint32_t* row;       /* !from_BFT: an owned colour bitmap (result of the set operations below) */
} BFT_annotation;

BFT* load_BFT(char* filename);
void free_cdbg(BFT* bft);

BFT_kmer* create_kmer(const char* kmer, int k);
BFT_kmer* create_empty_kmer(void);
void free_BFT_kmer(BFT_kmer* bft_kmer, int nb_bft_kmer);
void free_BFT_kmer_content(BFT_kmer* bft_kmer, int nb_bft_kmer);

BFT_kmer* get_kmer(const char* kmer, BFT* bft);
bool is_kmer_in_cdbg(BFT_kmer* bft_kmer);
uint32_t* query_sequence(BFT* bft, char* sequence, double threshold, bool canonical_search);

BFT_annotation* create_BFT_annotation(void);
void free_BFT_annotation(BFT_annotation* bft_annot);
BFT_annotation* get_annotation(BFT_kmer* bft_kmer);
bool presence_genome(uint32_t id_genome, BFT_annotation* bft_annot, BFT* bft);
uint32_t* get_list_id_genomes(BFT_annotation* bft_annot, BFT* bft);
uint32_t get_count_id_genomes(BFT_annotation* bft_annot, BFT* bft);

/* Set algebra over colour sets (include/bft.h:99-113, src/bft.c:421-613): bitmap operations on the decoded class rows.
 * nb_annotations BFT_annotation* follow; the result is a new annotation to free with free_BFT_annotation(). */
BFT_annotation* intersection_annotations(BFT* bft, uint32_t nb_annotations, ...);
BFT_annotation* union_annotations(BFT* bft, uint32_t nb_annotations, ...);
BFT_annotation* sym_difference_annotations(BFT* bft, uint32_t nb_annotations, ...);
uint32_t* intersection_list_id_genomes(uint32_t* list_a, uint32_t* list_b);

void set_neighbors_traversal(BFT* bft);
void unset_neighbors_traversal(BFT* bft);
BFT_kmer* get_neighbors(BFT_kmer* bft_kmer, BFT* bft);
BFT_kmer* get_predecessors(BFT_kmer* bft_kmer, BFT* bft);
BFT_kmer* get_successors(BFT_kmer* bft_kmer, BFT* bft);

/* Iteration (src/bft.c:1014-1075, src/extract_kmers.c): f is called once per stored k-mer with its ASCII and 2-bit
 * forms and a locator usable with get_annotation(); iteration stops when f returns 0. The k-mers are produced by the
 * device enumeration (bft_b200_extract_kmers) in arena order, not the reference's trie order. */
typedef size_t (*BFT_func_ptr)(BFT_kmer* bft_kmer, BFT* bft, va_list args);
void iterate_over_kmers(BFT* bft, BFT_func_ptr f, ...);
/* prefix_matching (include/bft.h:137, src/bft.c:1093-1141): f on every stored k-mer that starts with `prefix`
 * (1 <= strlen(prefix) <= k); returns whether at least one k-mer matched. */
bool prefix_matching(BFT* bft, char* prefix, BFT_func_ptr f, ...);
void v_iterate_over_kmers(BFT* bft, BFT_func_ptr f, va_list args);
void extract_kmers_to_disk(BFT* bft, char* filename_output, bool compressed_output);
size_t write_kmer_ascii_to_disk(BFT_kmer* bft_kmer, BFT* bft, va_list args);
size_t write_kmer_comp_to_disk(BFT_kmer* bft_kmer, BFT* bft, va_list args);

int queryBFT_kmerPresences_from_KmerFiles(BFT_Root* root, char* query_filename, int binary_file, char* output_filename);
int queryBFT_kmerBranching_from_KmerFiles(BFT_Root* root, char* query_filename, int binary_file);
void query_sequences_outputCSV(BFT_Root* root, char* query_filename, char* output_filename, double threshold, bool canonical_search);

/* ---- marking (include/bft.h:143-146, src/marking.c): one 2-bit flag per stored k-mer, kept on the host and keyed by
 * the k-mer's vertex index (bft_b200_query_vertex_ids) */
void set_marking(BFT* bft);
void unset_marking(BFT* bft);
void set_flag_kmer(uint8_t flag, BFT_kmer* bft_kmer, BFT* bft);
uint8_t get_flag_kmer(BFT_kmer* bft_kmer, BFT* bft);

/* ---- src/snippets.c (include/snippets.h): pan-genome k-mer classes, simple paths, traversals ------------------------
 * BFS / DFS / BFS_subgraph / DFS_subgraph are accepted by cdbg_traversal and get_nb_connected_component as the
 * reference's callers pass them; the traversal itself runs on the device (bft_b200_connected_components,
 * bft_b200_simple_paths), so calling them directly on one k-mer is not supported (exit(1), like any reference ERROR). */
#define V_NOT_VISITED 0
#define V_VISITED 1
size_t extract_core_kmers(BFT_kmer* kmer, BFT* graph, va_list args);
size_t extract_dispensable_kmers(BFT_kmer* kmer, BFT* graph, va_list args);
size_t extract_singleton_kmers(BFT_kmer* kmer, BFT* graph, va_list args);
void extract_pangenome_kmers_to_disk(BFT* graph, char* filename_output, BFT_func_ptr f);
void extract_simple_paths_to_disk(BFT* graph, char* filename_output);
void extract_simple_core_paths_to_disk(BFT* graph, double core_ratio, char* filename_output);
size_t BFS(BFT_kmer* kmer, BFT* graph, va_list args);
size_t BFS_subgraph(BFT_kmer* kmer, BFT* graph, va_list args);
size_t DFS(BFT_kmer* kmer, BFT* graph, va_list args);
size_t DFS_subgraph(BFT_kmer* kmer, BFT* graph, va_list args);
void cdbg_traversal(BFT* graph, BFT_func_ptr f, ...);
void get_nb_connected_component(BFT* graph, ...);

#ifdef __cplusplus
}
#endif
#endif
