/* bft_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's query path (GuillaumeHolley/BloomFilterTrie), used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker for the CUDA engine. It shares no code with
 * the product (bloomfiltertrie_b200/csrc): it keeps the reference's own pointer-linked layout (Node -> CC[] + UC,
 * byte-string suffix lines compared with memcmp, in-band cluster flags, annotation bytes decoded per query) where
 * the product flattens everything into integer-keyed SoA arenas.
 *
 * Parity pinned: yes — checked against the unmodified reference compiled from /root/reference (oracle/_ref, see
 * oracle/Makefile) on every case of tests/cases.py, and against the committed golden vectors in tests/golden/
 * (outputs of that same reference, tests/golden/make_golden.py) by tests/test_oracle.py.
 */
#ifndef BFT_ORACLE_H
#define BFT_ORACLE_H
#include <stddef.h>
#include <stdint.h>

typedef struct o_bft o_bft;

o_bft* o_load(const char* path, char* err, size_t errlen);      /* read_BFT_Root, src/write_to_disk.c:264-357 */
void o_free(o_bft* b);
int o_k(const o_bft* b);
int o_n_genomes(const o_bft* b);

/* isKmerPresent + get_annotation + get_list_id_genomes (src/presenceNode.c:1823-1921, src/bft.c:363-387, 622-641).
 * kmer: ceil(2k/8) bytes in the reference layout. Returns 1 if present and writes ids[0] = count, ids[1..] ascending
 * genome ids (ids must hold n_genomes + 1 entries); returns 0 if absent. */
int o_query_kmer(o_bft* b, const uint8_t* kmer, uint32_t* ids);

/* isBranchingRight / isBranchingLeft (src/branchingNode.c:16-110, 240-413): number of successors / predecessors. */
int o_branching_right(o_bft* b, const uint8_t* kmer);
int o_branching_left(o_bft* b, const uint8_t* kmer);

/* query_sequence (src/bft.c:1241-1351). Returns 0, or -1 where the reference would exit(1) on a bad character.
 * ids as in o_query_kmer. */
int o_query_sequence(o_bft* b, const char* seq, double threshold, int canonical, uint32_t* ids);


/* iterate_over_kmers / -extract_kmers (src/extract_kmers.c:3-597): every stored k-mer as ASCII (k characters each, no
 * separators) in the reference's own iteration order. Writes at most cap k-mers; returns how many the BFT holds. */
size_t o_extract_kmers(o_bft* b, char* out, size_t cap);

/* ---- graph traversals (bft_graph_oracle.c; reference src/snippets.c) ------------------------------------------------
 * kmers: every stored k-mer as ASCII, n * k characters, in the order iterate_over_kmers visits them. */
int64_t o_connected_components(o_bft* b, const char* kmers, size_t n, const uint32_t* ids, int n_ids, uint32_t* labels);
char* o_simple_paths(o_bft* b, const char* kmers, size_t n, double core_ratio, int faithful, size_t* n_bytes, int* longest);
void o_free_buf(void* p);

#endif
