/* bft_flatten.h — host-side owner of a flattened BFT (see bft_arena.h for the layout). */
#ifndef BFT_FLATTEN_H
#define BFT_FLATTEN_H

#include "bft_arena.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bft_arena {
    /* header of the .bft file (reference src/write_to_disk.c:34-86) */
    int k, W, n_genomes, n_levels;
    int r1, r2, treshold_compression, compressed;
    char** filenames; /* n_genomes NUL-terminated names */

    bft_entry_t* rootdir; /* BFT_ROOTDIR_SIZE */
    bft_node_t* nodes;   size_t n_nodes;
    bft_cc_t* ccs;       size_t n_ccs;
    uint8_t* firstcc;    size_t firstcc_bytes; /* first-CC tables and, for Nodes with few CCs, their Bloom filters */
    uint32_t* hpos;      size_t n_hpos;        /* Bloom-filter bit positions per idx14, one table per trie level in use */
    uint16_t* csr;       size_t n_csr;
    uint8_t* filter3;    size_t filter3_bytes;
    bft_entry_t* pref;   size_t n_pref;
    uint64_t* buckets;   size_t n_buckets; /* n_buckets * BFT_BUCKET_KEYS * W words */
    uint32_t* slotcls;   /* n_buckets * BFT_BUCKET_KEYS; NULL once the classes are embedded (cls_shift != 0) */
    uint64_t* ovf;       size_t n_ovf;     /* n_ovf * W words */
    uint32_t* ovfcls;    /* n_ovf; NULL once embedded */
    size_t n_lines;      /* inline suffix lines stored (in buckets + ovf) */
    uint64_t* uckeys;    /* n_uc_lines * W: Node-UC lines */
    uint32_t* uccls;     size_t n_uc_lines;
    uint8_t* uc_rank;    /* n_uc_lines: position of the line inside its UC as the reference stores it (enumeration order) */
    int cls_shift;
    uint32_t cls_mask;
    /* enumeration side tables (see bft_arena.h) */
    uint32_t* pref_low18; /* n_pref */
    uint32_t* pref_node;  /* n_pref */
    bft_path_t* node_path; /* n_nodes */
    uint64_t* pref_out;   /* n_pref */

    /* colour classes: distinct annotation byte strings (annotation ‖ extended byte, reference src/UC.c:171-239) */
    uint32_t* cls_off;   /* n_classes + 1 */
    uint8_t* cls_bytes;  size_t cls_bytes_len;
    size_t n_classes;
    size_t max_cls_len;

    /* comp_set_colors pools (reference include/annotation.h:51-55, 309-323), concatenated */
    int n_pools;
    int64_t* pool_last_index; /* n_pools */
    int32_t* pool_size_annot; /* n_pools */
    uint64_t* pool_off;       /* n_pools: byte offset of pool i in pool_bytes */
    uint8_t* pool_bytes; size_t pool_bytes_len;

    /* statistics */
    size_t n_kmers;         /* stored k-mers = lines + leaf prefixes */
    size_t n_leaf_prefixes;
    int max_cc_per_node;
    int max_depth;          /* deepest level reached, in Nodes (1 = root only) */
} bft_arena_t;

/* Parse a .bft file written by the reference's write_BFT_Root (src/write_to_disk.c:21-93) and flatten it.
 * Returns NULL and writes a message to err (if non-NULL) on failure. */
bft_arena_t* bft_arena_from_file(const char* path, char* err, size_t errlen);
bft_arena_t* bft_arena_from_memory(const uint8_t* buf, size_t len, char* err, size_t errlen);
void bft_arena_free(bft_arena_t* a);
size_t bft_arena_bytes(const bft_arena_t* a);
void bft_arena_view(const bft_arena_t* a, bft_view_t* v);

#ifdef __cplusplus
}
#endif
#endif
