/* bft_b200.h — C-ABI of the B200 batched query engine for Bloom Filter Tries.
 *
 * Plain C linkage, plain pointers and sizes; no torch or C++ types. One context = one .bft file flattened into
 * device-resident arenas on one GPU. Every entry point cites the reference interface it stands in for
 * (paths relative to the GuillaumeHolley/BloomFilterTrie tree). All query results are bit-exact with the
 * reference's own query path on the same .bft and inputs.
 *
 * Conventions
 *   - packed k-mers: W = bft_b200_kmer_words() (1, 2 or 4 for k <= 27 / 63 / 126) little-endian uint64 per k-mer, nucleotide i at bits 2i
 *     (A=0 C=1 G=2 T=3), i.e. the reference's byte layout (include/fasta.h:15, src/fasta.c:13-23) zero-padded to W
 *     words; bits above 2k must be zero.
 *     bft_b200_query_records takes the same k-mers as the reference's own ceil(2k/8)-byte records instead.
 *   - colour rows: RW = bft_b200_row_words() uint32 per item, genome g = bit (g & 31) of word g >> 5. A row is the
 *     reference's ascending id list (src/bft.c:622-641) as a bitmap; an absent k-mer has an all-zero row.
 *   - every function returns BFT_B200_OK (0) or a negative status; bft_b200_last_error() gives the message of the
 *     last failure on the calling thread. Where the reference would fprintf+exit(1) (include/useful_macros.h:33-43)
 *     the batched calls return a status / per-item flag instead; the drop-in wrappers in bft_compat.h keep exit(1).
 *   - "host" entry points take ordinary host pointers (pinned memory from bft_b200_host_alloc avoids a staging
 *     copy) and include the host<->device transfers; "_device" entry points take pointers into this context's GPU
 *     and only enqueue work on bft_b200_stream().
 *   - there is no CPU fallback: without a usable CUDA device every call fails with BFT_B200_ERR_CUDA.
 *   - a context serves one caller at a time (the reference's query API is not re-entrant either); use one
 *     context per thread / per GPU. The library changes no device-wide state: the tables every lookup touches are kept in
 *     L2 by the cache hints on the loads themselves (a persisting-L2 set-aside is an opt-in experiment, BFT_B200_L2_PERSIST=1;
 *     it is a device-wide limit that would outlive the context and take L2 away from everyone else).
 *     Environment knobs read by bft_b200_open: BFT_B200_KF_BITS (bits per stored k-mer of the L2-resident negative
 *     filter, default 6, 0 = off), BFT_B200_KF_MAX_MB (its size cap, default 36), BFT_B200_RKF_BITS / BFT_B200_RKF_SECTORS
 *     (fused root directory + filter table of the plain look-ups: filter bits per stored k-mer, default 5.5; or the
 *     number of 32-byte sectors per 9-nt prefix outright, 0 = off), BFT_B200_NO_DEEP=1 (no collapsed subtrees: look-ups
 *     walk the Nodes below the root).
 */
#ifndef BFT_B200_H
#define BFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BFT_B200_OK 0
#define BFT_B200_ERR_ARG (-1)    /* invalid argument */
#define BFT_B200_ERR_FILE (-2)   /* cannot read / not a supported .bft file */
#define BFT_B200_ERR_CUDA (-3)   /* CUDA runtime failure or no device */
#define BFT_B200_ERR_NOMEM (-4)

/* per-sequence status written by bft_b200_query_sequences */
#define BFT_B200_SEQ_OK 0
#define BFT_B200_SEQ_TOO_SHORT 1 /* shorter than k: all-zero row, as the reference (src/bft.c:1276-1277) */
#define BFT_B200_SEQ_BAD_CHAR 2  /* a character the reference would exit(1) on (src/bft.c:239, src/fasta.c:431-434) */

typedef struct bft_b200_ctx bft_b200_ctx;

const char* bft_b200_last_error(void);

/* load_BFT (include/bft.h:176, src/bft.c:1222 -> read_BFT_Root, src/write_to_disk.c:264-357) + the serializer:
 * reads the .bft file, flattens the trie into SoA arenas, uploads them to `device`, and decodes every distinct
 * colour annotation on the GPU (get_id_genomes_from_annot, src/annotation.c:2086-2250). */
int bft_b200_open(const char* bft_path, int device, bft_b200_ctx** out);
/* free_cdbg (include/bft.h:63) */
void bft_b200_close(bft_b200_ctx* ctx);

/* BFT_Root fields (include/Node.h:96-122): k, nb_genomes, filenames */
int bft_b200_k(const bft_b200_ctx* ctx);
int bft_b200_n_genomes(const bft_b200_ctx* ctx);
const char* bft_b200_genome_name(const bft_b200_ctx* ctx, int i);
int bft_b200_kmer_words(const bft_b200_ctx* ctx); /* W */
int bft_b200_row_words(const bft_b200_ctx* ctx);  /* RW = ceil(nb_genomes / 32) */
int bft_b200_device(const bft_b200_ctx* ctx);
void* bft_b200_stream(const bft_b200_ctx* ctx);   /* cudaStream_t the _device calls enqueue on */

/* arena statistics (what printMemoryUsedFromNode reports for the pointer structure, src/printMemory.c:68-254) */
typedef struct {
    uint64_t n_kmers, n_nodes, n_ccs, n_lines, n_prefixes, n_classes, arena_bytes, class_row_bytes;
    int max_cc_per_node, max_depth, n_pools;
    double flatten_seconds, upload_seconds, decode_seconds;
    uint64_t filter_bytes; /* stored-k-mer filter in L2 (0: none) */
    uint64_t rootkf_bytes; /* fused root directory + stored-k-mer filter of the plain look-ups (0: none) */
    uint64_t deep_bytes;   /* collapsed subtrees: one hashed block per root prefix that leads to child Nodes (0: none) */
} bft_b200_stats;
int bft_b200_get_stats(const bft_b200_ctx* ctx, bft_b200_stats* out);

/* pinned host memory for the host entry points */
void* bft_b200_host_alloc(size_t bytes);
void bft_b200_host_free(void* p);

/* ---- k-mer membership + colours -------------------------------------------------------------------------------
 * Batch form of get_kmer + is_kmer_in_cdbg + get_annotation + get_list_id_genomes (include/bft.h:125,126,97,115;
 * src/bft.c:216-248, 363-387, 622-641), i.e. the per-k-mer body of queryBFT_kmerPresences_from_KmerFiles
 * (src/file_io.c:732-752, 810-834). present[i] = 1/0; rows (n*RW) and class_ids (n) may be NULL.
 * class_ids[i] = index of the k-mer's distinct annotation (0xffffffff if absent); see bft_b200_class_rows. */
int bft_b200_query_kmers(bft_b200_ctx* ctx, const uint64_t* kmers, size_t n, uint8_t* present, uint32_t* rows,
                         uint32_t* class_ids);
int bft_b200_query_kmers_device(bft_b200_ctx* ctx, const uint64_t* d_kmers, size_t n, uint8_t* d_present,
                                uint32_t* d_rows, uint32_t* d_class_ids);
/* Same as the _device call with rows, plus the batch's hit count accumulated by the kernel itself into *d_n_present (a
 * device uint64, zeroed by the call): the "Nb k-mers present" of the reference driver (src/file_io.c:813) without a
 * second pass over the presence bytes. Any row width. */
int bft_b200_query_kmers_device_counted(bft_b200_ctx* ctx, const uint64_t* d_kmers, size_t n, uint8_t* d_present,
                                        uint32_t* d_rows, uint64_t* d_n_present);
/* Same, without zeroing: the kernel ADDS the batch's hit count to *d_counter with one system-scope atomic per thread
 * block. d_counter may be a peer mapping of a counter in another GPU's memory (bft_b200_peer_import): every rank of a
 * sharded query then accumulates into the owner's counter over NVLink from inside its query kernel — the reduction of
 * the reference driver's `Nb k-mers present` (src/file_io.c:813) fused into the kernel, no NCCL call on the path. */
int bft_b200_query_kmers_device_accumulate(bft_b200_ctx* ctx, const uint64_t* d_kmers, size_t n, uint8_t* d_present,
                                           uint32_t* d_rows, uint64_t* d_counter);
/* ASCII input, n k-mers of exactly k characters each, back to back: parseKmerCount (src/fasta.c:3-53) on the GPU.
 * valid[i] = 0 for a k-mer with a non-ACGTU character (the reference drops such lines, src/file_io.c:786-862);
 * its present/rows are zero. */
int bft_b200_query_kmers_ascii(bft_b200_ctx* ctx, const char* ascii, size_t n, uint8_t* valid, uint8_t* present,
                               uint32_t* rows, uint32_t* class_ids);
/* the decoded class table: n_classes * RW words, row c = colour set of class c (host copy owned by ctx) */
int bft_b200_class_rows(bft_b200_ctx* ctx, const uint32_t** rows, uint64_t* n_classes);
/* get_count_id_genomes (include/bft.h:116): number of genomes per class (host copy owned by ctx) */
int bft_b200_class_counts(bft_b200_ctx* ctx, const uint32_t** counts, uint64_t* n_classes);

/* ---- annotation set algebra -------------------------------------------------------------------------------------
 * Batch form of intersection_annotations / union_annotations / sym_difference_annotations (include/bft.h:99-113,
 * src/bft.c:421-613, which fold cmp_annots, src/annotation.c:2358-2552, over their arguments from left to right):
 * group g combines the colour sets of the classes class_ids[group_offs[g] .. group_offs[g+1]) (ids as returned by the
 * query calls; 0xffffffff = the empty set of an absent k-mer). rows: n_groups*RW bitmap words, counts: n_groups genome
 * counts (get_count_id_genomes of the result); either may be NULL. An empty group yields the empty set (the reference
 * exit(1)s on "no annotations given"). */
#define BFT_B200_SET_INTERSECTION 0
#define BFT_B200_SET_UNION 1
#define BFT_B200_SET_SYM_DIFFERENCE 2
int bft_b200_annotation_setop(bft_b200_ctx* ctx, int op, const uint32_t* class_ids, const uint64_t* group_offs,
                              size_t n_groups, uint32_t* rows, uint32_t* counts);
int bft_b200_annotation_setop_device(bft_b200_ctx* ctx, int op, const uint32_t* d_class_ids, const uint64_t* d_group_offs,
                                     size_t n_groups, uint32_t* d_rows, uint32_t* d_counts);

/* ---- sequences ------------------------------------------------------------------------------------------------
 * Batch form of query_sequence (include/bft.h:127, src/bft.c:1241-1351) / query_sequences_outputCSV
 * (src/file_io.c:1464-1574): sequence i = chars[offs[i] .. offs[i+1]). For every window: optional canonical pick
 * (reverse_complement + strcmp, src/bft.c:1287-1293), IUPAC windows skipped (src/fasta.c:357-363), lookup, colour
 * decode, per-genome hit count; genome g is reported iff count >= ceil(n_windows * threshold) (double arithmetic,
 * src/bft.c:1279). rows: n_seq*RW. status (may be NULL): BFT_B200_SEQ_*. 0 < threshold <= 1 (src/bft.c:1246-1247).
 * Like the reference (src/bft.c:1319) the scan of a sequence stops once no genome can reach the threshold any more (windows
 * found so far + windows left < need; tested every 32 windows, never earlier than the reference's own test): the row is all
 * zero, and a bad character past that point is not reported (the reference would not have reached it either). */
int bft_b200_query_sequences(bft_b200_ctx* ctx, const char* chars, const uint64_t* offs, size_t n_seq,
                             double threshold, int canonical, uint32_t* rows, uint8_t* status);
int bft_b200_query_sequences_device(bft_b200_ctx* ctx, const char* d_chars, const uint64_t* d_offs, size_t n_seq,
                                    double threshold, int canonical, uint32_t* d_rows, uint8_t* d_status);

/* ---- branching ------------------------------------------------------------------------------------------------
 * Batch form of isBranchingRight / isBranchingLeft (src/branchingNode.c:16-110, 240-413) as driven by
 * queryBFT_kmerBranching_from_KmerFiles (src/file_io.c:897-1020): succ[i] / pred[i] = number of successors /
 * predecessors of k-mer i present in the graph (0..4); *n_branching = #{i : succ[i] > 1 or pred[i] > 1}.
 * succ, pred and n_branching may each be NULL. */
int bft_b200_query_branching(bft_b200_ctx* ctx, const uint64_t* kmers, size_t n, uint8_t* succ, uint8_t* pred,
                             uint64_t* n_branching);
int bft_b200_query_branching_device(bft_b200_ctx* ctx, const uint64_t* d_kmers, size_t n, uint8_t* d_succ,
                                    uint8_t* d_pred, uint64_t* d_n_branching);
/* Batch form of get_neighbors (include/bft.h:156, src/bft.c:804-886): neighbour_classes[8*i + j] = colour class of
 * neighbour j of k-mer i, or 0xffffffff if that neighbour is not in the graph; j = 0..3 predecessors (A,C,G,T
 * prepended), 4..7 successors (A,C,G,T appended) — the reference's order. Host pointers. */
int bft_b200_query_neighbors(bft_b200_ctx* ctx, const uint64_t* kmers, size_t n, uint32_t* neighbour_classes);
/* Successor lookups follow the reference bit for bit by default, including its behaviour at the leaf level of deep
 * tries (presenceNeighborsRight clears nucleotide 7 of the last 9-nt prefix before probing the CCs,
 * src/presenceNode.c:719-723, so there it reports the successors of a neighbouring k-mer). exact != 0 keeps that;
 * exact == 0 switches to plain set membership of the four successors. */
int bft_b200_set_reference_exact_branching(bft_b200_ctx* ctx, int exact);

/* ---- enumeration --------------------------------------------------------------------------------------------------
 * Batch form of iterate_over_kmers / extract_kmers_to_disk (include/bft.h:88,164; src/bft.c:255-283,
 * src/extract_kmers.c): every k-mer stored in the BFT (bft_b200_get_stats().n_kmers of them) with its colour class
 * and, optionally, its colour row. The k-mers come out in arena order — deterministic for a given .bft, but not the
 * reference's trie order; the SET of (k-mer, colours) pairs is identical. capacity = room in the output arrays, in
 * k-mers; kmers is required, class_ids and rows may be NULL. */
int bft_b200_extract_kmers(bft_b200_ctx* ctx, uint64_t* kmers, uint32_t* class_ids, uint32_t* rows, size_t capacity,
                           uint64_t* n_written);
int bft_b200_extract_kmers_device(bft_b200_ctx* ctx, uint64_t* d_kmers, uint32_t* d_class_ids, size_t capacity);
/* extract_kmers_to_disk (src/bft.c:255-283): compressed_output != 0 writes the `kmers_comp` layout (two header lines:
 * k, count; then ceil(2k/8)-byte records), else one ASCII k-mer per line. Same records as the reference's file, in
 * arena order. */
int bft_b200_extract_kmers_file(bft_b200_ctx* ctx, const char* path, int compressed_output);

/* ---- multi-GPU: results written straight into a peer GPU's memory over NVLink ------------------------------------
 * The path shards by query with the arena replicated per GPU; the only exchange is the gather of results on one
 * GPU. Instead of a collective after the kernel, a rank can hand the *_device entry points an output pointer that
 * lives on ANOTHER GPU of the box (mapped through CUDA IPC, one process per GPU): the kernel's own result stores
 * then travel over NVLink/NVSwitch into the right slice of the gathered array — compute and gather in one kernel.
 *   owner rank:  bft_b200_device_alloc -> bft_b200_peer_export (64-byte handle, ship it with any host channel)
 *   other ranks: bft_b200_peer_import -> pointer usable as d_rows / d_present / ... (+ element offset of the shard)
 * Close imported mappings with bft_b200_peer_close and free the buffer with bft_b200_device_free on the owner. */
int bft_b200_device_alloc(bft_b200_ctx* ctx, size_t bytes, void** d_ptr);
int bft_b200_device_free(bft_b200_ctx* ctx, void* d_ptr);
int bft_b200_peer_export(bft_b200_ctx* ctx, void* d_ptr, unsigned char handle[64]);
int bft_b200_peer_import(bft_b200_ctx* ctx, const unsigned char handle[64], void** d_ptr);
int bft_b200_peer_close(bft_b200_ctx* ctx, void* d_ptr);

/* ---- file-level drivers (the CLI-visible bytes) ---------------------------------------------------------------
 * queryBFT_kmerPresences_from_KmerFiles (src/file_io.c:651-895): CSV of colour rows; returns #present via out.
 * queryBFT_kmerBranching_from_KmerFiles (src/file_io.c:897-1020): count of branching k-mers.
 * query_sequences_outputCSV (src/file_io.c:1464-1574); besides the reference's one-sequence-per-line layout the
 * sequence file may be FASTA (first character '>' or ';': headers dropped, the lines of a record joined) or FASTQ
 * (first character '@'): one CSV row per record, byte-identical to what the reference writes for the same records
 * flattened to one per line. A text k-mer file ("kmers") is parsed on the GPU (parseKmerCount, src/fasta.c:3-53). */
int bft_b200_query_kmers_file(bft_b200_ctx* ctx, const char* query_path, int binary_file, const char* csv_path,
                              uint64_t* n_present);
int bft_b200_query_branching_file(bft_b200_ctx* ctx, const char* query_path, int binary_file, uint64_t* n_branching);
int bft_b200_query_sequences_file(bft_b200_ctx* ctx, const char* query_path, const char* csv_path, double threshold,
                                  int canonical);

/* Roofline accounting helper: over a device-resident batch, sums of out[0] Nodes probed, out[1]
 * binary-search depths ceil(log2(lines+1)), out[2] found k-mers, out[3] CC Bloom filters the reference layout would
 * probe, out[4] lines in the searched suffix blocks; and of the arena's own walk: out[5] bucket / Node-UC searches the
 * product path performs (one random HBM access each), out[6] k-mers the stored-k-mer filter rejects; out[7] spare.
 * Diagnostic; synchronous. */
int bft_b200_kmer_walk_stats_device(bft_b200_ctx* ctx, const uint64_t* d_kmers, size_t n, uint64_t out[8]);

/* Random-access roofline probe: rate (loads/s) of independent 8-byte loads at random offsets of a table_bytes table
 * (choose it far larger than the 126 MB L2). Diagnostic; synchronous. */
int bft_b200_random_gather_probe(bft_b200_ctx* ctx, size_t table_bytes, size_t n_loads, double* loads_per_sec);

/* wait for everything enqueued on the context's streams */
int bft_b200_sync(bft_b200_ctx* ctx);

/* number of kernels this context has launched (bench.py's gpu_launches) */
uint64_t bft_b200_launch_count(const bft_b200_ctx* ctx);

/* ---- the reference's own record format ------------------------------------------------------------------------------
 * records: n k-mers of bft_b200_record_bytes() = ceil(2k/8) bytes each, nucleotide i at bits 2(i%4) of byte i/4 — the
 * records of a kmers_comp file and BFT_kmer.kmer_comp (src/fasta.c:3-53, src/file_io.c:721-774); pad bits ignored.
 * rows: bft_b200_row_bytes() = ceil(n_genomes/8) bytes per k-mer, bit (g % 8) of byte g/8 = genome g; all zero for an
 * absent k-mer (a stored k-mer always has at least one genome). present (optional): 1 byte per k-mer.
 * n_present (optional): the number the reference driver prints as "Nb k-mers present".
 * Same answers as bft_b200_query_kmers; 7 + 13 instead of 8 + 17 bytes per k-mer cross PCIe at k = 27 / 100 genomes,
 * and that transfer is what bounds the host-facing call. The _device variant takes 16-byte aligned device buffers. */
int bft_b200_record_bytes(const bft_b200_ctx* ctx);
int bft_b200_row_bytes(const bft_b200_ctx* ctx);
int bft_b200_query_records(bft_b200_ctx* ctx, const uint8_t* records, size_t n, uint8_t* present, uint8_t* rows,
                           uint64_t* n_present);
int bft_b200_query_records_device(bft_b200_ctx* ctx, const uint8_t* d_records, size_t n, uint8_t* d_present,
                                  uint8_t* d_rows, uint64_t* d_n_present);
/* The same answers without the zeros: present_bits = one bit per k-mer (bit i % 8 of byte i / 8; (n + 7) / 8 bytes),
 * rows = the bft_b200_row_bytes()-byte rows of the PRESENT k-mers only, back to back in query order (capacity: n rows),
 * *n_present = how many there are. The row of k-mer i is rows[r * row_bytes] with r = number of set bits before i —
 * a CSV writer (src/file_io.c:740-752) walks both with one running index. On a batch where half the k-mers are absent
 * half the row bytes never cross PCIe. */
int bft_b200_query_records_compact(bft_b200_ctx* ctx, const uint8_t* records, size_t n, uint8_t* present_bits,
                                   uint8_t* rows, uint64_t* n_present);

/* ---- graph traversals (reference src/snippets.c) ----------------------------------------------
 * The coloured de Bruijn graph is materialised on the device once (one vertex per stored k-mer, in the order of
 * bft_b200_extract_kmers; 8 neighbour look-ups per vertex) and the reference's traversal snippets run on it as
 * data-parallel kernels. Neighbourhood is plain set membership of the 8 possible neighbours (get_neighbors,
 * include/bft.h:156). Results that do not depend on the reference's iteration order are identical to its own. */

/* Build the device graph now (otherwise the first traversal call does). bft_b200_graph_release frees it. */
int bft_b200_graph_prepare(bft_b200_ctx* ctx);
int bft_b200_graph_release(bft_b200_ctx* ctx);

/* Vertex index of each k-mer (its position in the order of bft_b200_extract_kmers), 0xffffffff when the k-mer is not
 * stored. This is the handle for per-k-mer state kept by the caller — what the reference stores as marks inside the
 * trie (set_marking / set_flag_kmer / get_flag_kmer, include/bft.h:143-146, src/marking.c). */
int bft_b200_query_vertex_ids(bft_b200_ctx* ctx, const uint64_t* kmers, size_t n, uint32_t* vertex_ids);

/* adj[8 * i + j]: vertex index of the j-th possible neighbour of k-mer i (j = 0-3 predecessors prepending A,C,G,T;
 * 4-7 successors appending A,C,G,T — the order of get_neighbors), or 0xffffffff when it is not in the graph.
 * capacity in k-mers, >= stats.n_kmers. */
int bft_b200_graph_adjacency(bft_b200_ctx* ctx, uint32_t* adj, size_t capacity);

/* get_nb_connected_component (src/snippets.c:937-958) with BFS / DFS (n_ids == 0: the whole graph) or with
 * BFS_subgraph / DFS_subgraph (n_ids > 0: the subgraph of the k-mers whose colour set holds every listed genome id,
 * is_in_subgraph :824-881; ids strictly ascending, as that function requires). labels (optional, stats.n_kmers
 * entries, order of bft_b200_extract_kmers): smallest vertex index of the k-mer's component, 0xffffffff for k-mers
 * outside the subgraph. */
int bft_b200_connected_components(bft_b200_ctx* ctx, const uint32_t* genome_ids, int n_ids, uint64_t* n_components,
                                  uint32_t* labels);

/* extract_simple_core_paths_to_disk (src/snippets.c:346-596): every maximal non-branching path over k-mers with fewer
 * than two successors and fewer than two predecessors, consecutive k-mers sharing at least (int)(core_ratio * genomes)
 * genomes; one line per path (first k-mer, then the last character of each following k-mer, then a newline).
 * core_ratio = 0 is extract_simple_paths_to_disk (:115-344). The lines come in vertex order of their first k-mer, not
 * in the reference's trie order; a cycle of non-branching k-mers is opened at its smallest vertex.
 * *paths is malloc'd by the library (release with bft_b200_free), *n_bytes its length; n_paths / longest (characters
 * of the longest line, the reference's "Longest simple path has %d nuc.") are optional. */
int bft_b200_simple_paths(bft_b200_ctx* ctx, double core_ratio, char** paths, size_t* n_bytes, uint64_t* n_paths,
                          uint64_t* longest);
int bft_b200_simple_paths_file(bft_b200_ctx* ctx, double core_ratio, const char* out_path, uint64_t* n_paths,
                               uint64_t* longest);
void bft_b200_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* BFT_B200_H */
