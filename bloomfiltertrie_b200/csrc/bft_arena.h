/* bft_arena.h — the flattened Bloom Filter Trie ("arena") and the walk over it.
 *
 * The reference keeps a BFT as a pointer-linked burst trie (Node -> CC[] + UC, CC -> UC buckets + child Nodes;
 * reference include/Node.h:55-58, include/CC.h:34-67, include/UC.h:13-20). The serializer (bft_flatten.c)
 * turns one .bft file into the structure-of-arrays arena declared here; the same arrays are uploaded verbatim
 * into HBM. Everything in the arena is derived exactly from the reference's own bytes — no result changes:
 *
 *   nodes[]    one descriptor per trie Node.
 *   firstcc[]  per Node with more than BFT_BF_DIRECT_MAX CCs: 16384 bytes, entry idx14 -> index of the FIRST CC of the node whose Bloom filter
 *              fires for the 14-bit hash index idx14 (0xff: none fires). The reference probes the CC chain in order
 *              and lets the first hit decide (src/presenceNode.c:1354-1550); the probe depends only on idx14
 *              (src/presenceNode.c:1341-1343), so this table is that chain, precomputed (the reference builds the
 *              same table for its root in get_bf_presence_per_cc, src/presenceNode.c:1213-1270).
 *              A Node with at most BFT_BF_DIRECT_MAX CCs (every Node below the root of a deep trie has exactly one)
 *              keeps the regenerated Bloom filters themselves instead (bf_mode = 1: n_cc filters of bf_stride bytes,
 *              188 bytes each at the reference's 1504-bit filters) and is probed the way the reference probes it: the
 *              two bit positions of idx14 (hpos[], 64 KB per trie level, L2-resident) tested in each filter in turn.
 *              16 KB per Node would put the tables of a 10^4-Node trie (160 MB) out of L2; 188 bytes keeps them in.
 *   ccs[]      one descriptor per CC.
 *   csr[]      per CC: 2^p + 1 uint16, csr[pu] = number of stored prefixes whose p_u is < pu. Replaces the
 *              filter2 bit test + rank (SkipFilter2) + select in the cluster-start bits (SkipFilter3 /
 *              extra_filter3 / in-band flags) of findCluster (src/presenceNode.c:1578-1821).
 *   filter3[]  the p_v values, as stored (bytes for s=8, nibbles for s=4; src/presenceNode.c:1478-1479).
 *   pref[]     per stored prefix: where its suffixes live (inline line range, child Node, or leaf colour class).
 *              Replaces children_type prefix sums (count_children / count_nodes, include/CC.h:471-550).
 *   buckets[]  every CC inline suffix line, as W little-endian 64-bit words holding the suffix as an integer
 *              (nucleotide i at bits 2i; replaces memcmp in binary_search_UC, src/UC.c:81-124), hashed by its own
 *              hash into fixed-size buckets of BFT_BUCKET_KEYS = 4 slots — one 32-byte sector for W = 1, 64 bytes for
 *              W = 2, a 128-byte line for W = 4: a prefix with cnt suffixes owns B consecutive buckets and suffix x
 *              lives in bucket (hash(x) * B) >> 64. A lookup therefore touches exactly one aligned bucket: ONE random
 *              DRAM access per k-mer. For W >= 2, B is a power of two >= cnt/2 and a bucket that would hold more than 4
 *              suffixes keeps 3 and an overflow descriptor pointing into ovf[] (rare: the load factor is <= 1/2). For
 *              W = 1 the 32-byte buckets work in PAIRS, because HBM delivers 64 bytes per fetch anyway: B is even,
 *              about cnt/2.4 (load 0.6), a suffix whose own bucket is full sits in the other half of the same 64-byte
 *              line, and only a pair that would hold more than 8 falls back to descriptors.
 *              When the colour-class ids fit above the widest suffix (cls_shift != 0) the line's class is stored in
 *              the spare top bits of its most significant word, so a hit needs no second load; otherwise
 *              slotcls[]/ovfcls[] hold it. Bit 63 of the top word marks empty slots / overflow descriptors.
 *   uckeys[] / uccls[]  the Node-UC lines (whole k-mer remainders, src/presenceNode.c:1554-1573) and their classes.
 *   rootdir[]  262144 entries: the complete answer of the root Node probe for every possible 9-nt prefix, indexed
 *              by the low 18 bits of the packed k-mer. Every query passes through the root, so its probe is
 *              collapsed into one 8-byte load.
 *   rootdir_fast[] / dbuckets[]  (device only, built by k_deep_insert from the arena's own enumeration) a second root
 *              directory in which every prefix whose suffixes live in a child Node (BFT_KIND_NODE) points at ONE hashed block
 *              holding every k-mer stored anywhere below it (BFT_KIND_DEEP): the whole subtree — Nodes, their Bloom filters,
 *              csr, filter3, pref and per-prefix buckets, five or six dependent random accesses per level — collapsed into
 *              one bucket access, like an inline prefix with thousands of suffixes. A block is 2^lb buckets of
 *              BFT_BUCKET_KEYS slots at load <= 1/2; a suffix that finds its bucket full moves to the next one (linear
 *              probing), so there is no overflow area. Valid for every look-up that is plain set membership — all of them
 *              except the leaf-level successor quirk — and that does not need the storage location (the graph build and the
 *              enumeration keep walking the structure). On the 100 x 5 Mbp pan-genome at k = 63 (136 584 Nodes) this took
 *              the DRAM traffic of a neighbour look-up from 5.3 transactions to one.
 *   colour classes: distinct annotation byte strings (cls_off/cls_bytes) + the comp_set_colors pools; decoded on
 *              the device once per arena into class rows (bft_kernels.cu: k_decode_classes).
 *   kfilter[]  (device only, built by k_kf_insert from the arena's own enumeration) a blocked Bloom filter over the
 *              STORED K-MERS — not one of the reference's Bloom filters, which cover 9-nt prefixes: 32-byte blocks,
 *              4 bits per k-mer (one in each 64-bit word of the block), ~8 bits per stored k-mer, sized to stay in
 *              L2. A k-mer the filter rejects is in no Node, so the walk (and its one random HBM access) is
 *              skipped; a k-mer it accepts is looked up as before, so answers never change. The reference spends
 *              its Bloom filters on choosing a CC; an absent k-mer still costs it the full descent.
 *   rootkf[]   (device only, built by k_rkf_insert) the root directory and a stored-k-mer filter FUSED into one table:
 *              every 9-nt prefix owns S 32-byte sectors, each holding the prefix's rootdir entry (8 bytes, repeated) and 192
 *              filter bits; a k-mer reads the ONE sector its hash names and gets both the root probe's answer and the
 *              filter's verdict. Measured on B200 (tools/mix_probe.cu) the walk is bound as much by the number of random
 *              L2 requests per k-mer as by its one HBM access: a random sector served by L2 costs about 1/300 G s
 *              chip-wide, so rootdir + kfilter as two requests cost 0.5 ms per 125 M k-mers more than one. Used by the
 *              plain look-ups (k-mers, records, sequences); the neighbour engine keeps kfilter, whose block is chosen by
 *              the k-mer's middle so that 8 neighbours share two blocks.
 *
 * The walk functions below are plain C, compiled for the device by nvcc (the product path) and for the host by
 * the flattener (to fill rootdir) and by tests/tools (to debug the arena without a GPU). The shipped library
 * exposes no host query path.
 */
#ifndef BFT_ARENA_H
#define BFT_ARENA_H

#include <stdint.h>
#include <stddef.h>

#ifdef __CUDACC__
#define BFT_HD __host__ __device__ __forceinline__
#else
#define BFT_HD static inline
#endif

/* read-only loads: non-coherent path on the device, plain loads on the host */
#ifdef __CUDA_ARCH__
#define BFT_LD8(p) __ldg((const unsigned char*)(p))
#define BFT_LD16(p) __ldg((const unsigned short*)(p))
#define BFT_LD32(p) __ldg((const unsigned int*)(p))
#define BFT_LD64(p) __ldg((const unsigned long long*)(p))
#else
#define BFT_LD8(p) (*(const uint8_t*)(p))
#define BFT_LD16(p) (*(const uint16_t*)(p))
#define BFT_LD32(p) (*(const uint32_t*)(p))
#define BFT_LD64(p) (*(const uint64_t*)(p))
#endif

/* BFT_OPAQUE(x): hides a 64-bit value from the optimiser at this point (no instruction is emitted). Used where a word of a
 * W-word key is picked by a run-time index through "select by value" loops: LLVM recognises the pattern and turns it back
 * into an indexed access, which puts the key in local memory (LDL/STL on the lookup path). */
#ifdef __CUDA_ARCH__
#define BFT_OPAQUE(x) asm volatile("" : "+l"(x))
#else
#define BFT_OPAQUE(x) ((void)0)
#endif

#define BFT_NB_CHAR_SUF_PREF 9         /* reference include/default_param.h:12 */
#define BFT_PREFIX_BITS 18
#define BFT_N_IDX14 16384
#define BFT_ROOTDIR_SIZE (1u << BFT_PREFIX_BITS)
#define BFT_FIRSTCC_NONE 0xffu
#define BFT_BF_DIRECT_MAX 4            /* Nodes with at most this many CCs keep their Bloom filters instead of a first-CC table */
#define BFT_MAX_WORDS 4                /* k <= 126 (252 bits), the reference's KMER_LENGTH_MAX */
#define BFT_BUCKET_KEYS 4              /* slots per bucket */
#define BFT_MAX_LB 8                   /* at most 256 buckets per prefix */
#define BFT_SLOT_EMPTY 0xffffffffffffffffULL
#define BFT_SLOT_SPECIAL (1ULL << 63)  /* top word: empty slot (all ones) or overflow descriptor (count << 32 | start) */

/* kinds of a prefix-probe answer */
#define BFT_KIND_ABSENT 0u /* a CC's Bloom filter fired but the prefix is not stored there: k-mer absent */
#define BFT_KIND_UC 1u     /* no CC fired: search the Node's own UC lines [a, a+n) for the whole remainder */
#define BFT_KIND_INLINE 2u /* prefix stored, suffixes inline: search lines [a, a+n) for the shifted remainder */
#define BFT_KIND_NODE 3u   /* prefix stored, suffixes in child Node a */
#define BFT_KIND_LEAF 4u   /* leaf level (9 nt left): prefix stored, a = colour class */
#define BFT_KIND_DEEP 5u   /* rootdir_fast only: every k-mer below this prefix in one block of 2^lb dbuckets starting at a */
#define BFT_KIND_SHIFT 28
#define BFT_LB_SHIFT 24                /* DEEP entries: log2(#buckets) of the block */
#define BFT_LB_MASK 0xfu
#define BFT_CNT_MASK ((1u << BFT_LB_SHIFT) - 1u)
#define BFT_NBK_SHIFT 8                /* INLINE entries: count of suffixes in bits 0-7 (at most 255), number of buckets of */
#define BFT_NBK_MASK 0xffffu           /* the prefix's block in bits 8-23 */
#define BFT_INLINE_CNT(e_) ((e_).b & 0xffu)
#define BFT_INLINE_NBK(e_) (((e_).b >> BFT_NBK_SHIFT) & BFT_NBK_MASK)
/* One-word keys (k <= 27): a bucket is a 32-byte sector but HBM delivers 64 bytes per fetch, so buckets work in PAIRS — a block
 * has an even number of them, a suffix that finds its own bucket full goes to the other half of the 64-byte line (fetched by the
 * same DRAM transaction), and the block can run at a load of 0.6 instead of 0.25-0.5 with a power-of-two bucket count. */
#define BFT_PAIRED(W_) ((W_) == 1)

#define BFT_CLS_NONE 0xffffffffu

typedef struct {
    uint32_t a; /* first bucket (INLINE) / first UC line (UC) / node id (NODE) / class id (LEAF) */
    uint32_t b; /* kind << 28 | log2(#buckets) << 24 | count */
} bft_entry_t;

typedef struct {
    uint32_t cc_begin; /* first CC descriptor */
    uint32_t n_cc;
    uint32_t fc_off;   /* byte offset of this node's firstcc table (valid when n_cc > 0) */
    uint32_t uc_begin; /* first line of the Node's own UC */
    uint32_t uc_n;     /* number of lines in the Node's own UC */
    uint32_t bf_mode;  /* 0: fc_off is a 16384-byte first-CC table; 1: fc_off is n_cc Bloom filters of bf_stride bytes */
    uint32_t hp_off;   /* bf_mode 1: offset (uint32 units) of this level's bit-position table in hpos[] */
    uint32_t bf_stride;
} bft_node_t;

typedef struct {
    uint32_t csr_off;  /* offset into csr[] (uint16 units) */
    uint32_t f3_off;   /* byte offset into filter3[] */
    uint32_t pref_off; /* offset into pref[] */
    uint16_t nb_elem;
    uint8_t s;         /* bits of p_v: 8 or 4 (reference CC.type bits 1-5) */
    uint8_t pad;
} bft_cc_t;

typedef struct {
    uint64_t acc[BFT_MAX_WORDS]; /* nucleotides of the 9-nt prefixes above this Node, at their final bit positions */
    uint32_t depth;              /* number of prefixes above (0 for the root) */
    uint32_t uc_out_lo, uc_out_hi; /* enumeration index of this Node's first own-UC k-mer (64 bits, split) */
    uint32_t pad;
} bft_path_t;

/* Read-only view of an arena; pointers are host pointers on the host and device pointers in kernels. */
typedef struct {
    const bft_entry_t* rootdir;
    const bft_node_t* nodes;
    const bft_cc_t* ccs;
    const uint8_t* firstcc;
    const uint32_t* hpos;     /* per trie level in use: 16384 entries h1 | h2 << 16, the two Bloom-filter bit positions of idx14 */
    const uint16_t* csr;
    const uint8_t* filter3;
    const bft_entry_t* pref;
    const uint64_t* buckets;  /* n_buckets * BFT_BUCKET_KEYS * W words (inline suffix lines, bucketed) */
    const uint64_t* ovf;      /* n_ovf * W words (suffixes that did not fit their bucket) */
    const uint32_t* slotcls;  /* n_buckets * BFT_BUCKET_KEYS, only when cls_shift == 0 */
    const uint32_t* ovfcls;   /* n_ovf, only when cls_shift == 0 */
    const uint64_t* uckeys;   /* n_uc_lines * W words (Node-UC lines) */
    const uint32_t* uccls;    /* n_uc_lines */
    const uint8_t* uc_rank;   /* n_uc_lines: index of the line inside its UC in the reference's stored order */
    /* enumeration side tables (iterate_over_kmers / -extract_kmers): where each stored prefix sits in the trie */
    const uint32_t* pref_low18;  /* per stored prefix: its 9 nucleotides as they appear in the packed k-mer */
    const uint32_t* pref_node;   /* per stored prefix: the Node whose CC holds it */
    const bft_path_t* node_path; /* per Node: the k-mer bits fixed by the path from the root, and the depth */
    const uint64_t* pref_out;    /* per stored prefix: index of its first k-mer in the enumeration order (the reference's
                                  * iterate_over_kmers order, depth first; NODE prefixes: first k-mer of the subtree) */
    int k;
    int W;             /* 64-bit words per key: 1 for k <= 27, 2 for k <= 63, 4 for k <= 126 */
    int cls_shift;     /* != 0: class id of an inline line = (top word >> cls_shift) & cls_mask, suffix = the bits below */
    uint32_t cls_mask;
    /* storage locations ("loc"): one number per place a k-mer can be stored — bucket slot g -> g, overflow line i ->
     * loc_ovf + i, Node-UC line i -> loc_uc + i, leaf prefix i -> loc_leaf + i; n_loc = loc_leaf + n_pref. The graph
     * traversals (bft_graph.cuh) key their vertex table and marks on it, the way the reference keeps its marks inside
     * the UC a k-mer is stored in (src/marking.c). */
    uint32_t loc_ovf, loc_uc, loc_leaf;
    /* stored-k-mer filter (see above); kf_blocks == 0: none. kf_quirk_safe != 0: no Node sits at the leaf level, so a
     * successor lookup with the reference's leaf-level quirk (bft_node_probe_ex) is plain set membership too. */
    const uint64_t* kfilter;
    uint32_t kf_blocks;
    uint32_t kf_quirk_safe;
    /* fused root directory + stored-k-mer filter (see above); rkf_sectors == 0: none. Sector (prefix * rkf_sectors + j):
     * word 0 = the rootdir entry of the prefix, words 1-3 = 3 x 64 filter bits. */
    const uint64_t* rootkf;
    uint32_t rkf_sectors;
    uint32_t rows_keep; /* != 0: the class-row table is small enough to be worth an evict-last priority in L2 (device only) */
    /* collapsed subtrees (see above); rootdir_fast == NULL: none (no Node below the root, or no memory for the blocks) */
    const bft_entry_t* rootdir_fast;
    const uint64_t* dbuckets;  /* n_dbuckets * BFT_BUCKET_KEYS * W words */
    const uint32_t* dslotcls;  /* n_dbuckets * BFT_BUCKET_KEYS, only when cls_shift == 0 */
} bft_view_t;

/* ---- prefix bit manipulation --------------------------------------------------------------------------------
 * low18 = the 9 leading nucleotides as stored in the packed k-mer (nuc i at bits 2i).
 * The reference works on the same 9 nucleotides MSB-first (reverse_word_8, src/presenceNode.c:1327-1329), hashes
 * nuc1..nuc7 (14 bits, :1341-1343) and stores the prefix rotated as nuc1..nuc8,nuc0 (:1367-1371). */
BFT_HD uint32_t bft_msb_first18(uint32_t low18) {
    uint32_t r = 0;
    for (int i = 0; i < 9; i++) r |= ((low18 >> (2 * i)) & 3u) << (2 * (8 - i));
    return r;
}
BFT_HD uint32_t bft_idx14(uint32_t r18) { return (r18 >> 2) & 0x3fffu; }
BFT_HD uint32_t bft_rot18(uint32_t r18) { return ((r18 << 2) & 0x3ffffu) | (r18 >> 16); }

BFT_HD bft_entry_t bft_ld_entry(const bft_entry_t* p) {
    bft_entry_t e;
#ifdef __CUDA_ARCH__
    /* the 2 MB root directory is every look-up's first stop: kept in L2 (evict-last policy) against the streamed batch. Loads
     * narrower than 256 bits take the priority through a cache-policy operand (createpolicy), not a qualifier. */
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm("ld.global.nc.L2::cache_hint.v2.u32 {%0,%1}, [%2], %3;" : "=r"(e.a), "=r"(e.b) : "l"(p), "l"(pol));
#else
    e = *p;
#endif
    return e;
}

BFT_HD bft_entry_t bft_mk_entry(uint32_t kind, uint32_t a, uint32_t n) {
    bft_entry_t e;
    e.a = a;
    e.b = (kind << BFT_KIND_SHIFT) | (n & BFT_CNT_MASK);
    return e;
}

/* a bucket's 4*W words. On the device each 32-byte half is ONE 256-bit load (LDG.E.256, new on sm_100): measured on
 * B200 (tools/gather_probe.cu) a random 256-bit load sustains the same 37.9 G accesses/s as a random 8-byte load,
 * while two 128-bit loads of the same sector reach only 31-35 G/s. The L2::64B qualifier caps the sector promotion
 * on a miss (128 B of HBM traffic per random access by default, 64 B with it). A bucket is read once per lookup and
 * 771 MB of them never fit L2: evict-first / no L1 allocation keeps them from displacing the L2-resident tables. */
BFT_HD void bft_ld_bucket(const uint64_t* p, uint64_t* out, const int W) {
#ifdef __CUDA_ARCH__
    for (int i = 0; i < W; i++)
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(out[4 * i]), "=l"(out[4 * i + 1]), "=l"(out[4 * i + 2]), "=l"(out[4 * i + 3])
                     : "l"(p + 4 * i));
#else
    for (int i = 0; i < BFT_BUCKET_KEYS * W; i++) out[i] = p[i];
#endif
}

/* ---- stored-k-mer filter ---------------------------------------------------------------------------------------
 * The BLOCK of a k-mer is chosen by a hash of its middle k-2 nucleotides (first and last dropped); the four bit
 * positions inside the block by a hash of that plus the two end nucleotides. The four successors of a k-mer differ
 * only in their last nucleotide and its four predecessors only in their first, so the eight neighbour look-ups of a
 * branching query (k_query_branching, k_graph_adjacency) read TWO filter blocks — two L2 sectors — instead of eight. */
typedef struct { uint32_t block; uint32_t b0, b1, b2, b3; } bft_kf_pos_t;

/* stage 1: 64-bit mix of the middle (W words, bits above 2(k-2) zero) */
BFT_HD uint64_t bft_kf_mix_mid(const uint64_t* mid, const int W) {
    uint64_t x = mid[0];
    for (int w = 1; w < W; w++) x = (x ^ (x >> 29)) * 0x9FB21C651E98DF25ULL + mid[w];
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDULL;
    x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ULL;
    x ^= x >> 33;
    return x;
}

/* stage 2: block from the middle's hash, bit positions from that and the two end nucleotides (ends = first | last << 2) */
BFT_HD bft_kf_pos_t bft_kf_finish(const uint64_t x, const uint32_t ends, const uint32_t n_blocks) {
    uint64_t y = (x ^ ((uint64_t)(ends + 1u) * 0xD6E8FEB86659FD93ULL)) * 0x9E3779B97F4A7C15ULL;
    y ^= y >> 29;
    bft_kf_pos_t p;
    p.block = (uint32_t)(((x >> 32) * (uint64_t)n_blocks) >> 32);
    p.b0 = (uint32_t)(y >> 58);
    p.b1 = (uint32_t)(y >> 52) & 63u;
    p.b2 = (uint32_t)(y >> 46) & 63u;
    p.b3 = (uint32_t)(y >> 40) & 63u;
    return p;
}

BFT_HD bft_kf_pos_t bft_kf_pos(const uint64_t* kmer, const int W, const int k, const uint32_t n_blocks) {
    /* middle = (kmer >> 2) without its top nucleotide; word indices are compared, never computed, so that W-word arrays
     * stay in registers on the device */
    uint64_t mid[BFT_MAX_WORDS];
    const int mid_bits = 2 * (k - 2), top = 2 * (k - 1);
    uint64_t last = 0;
    for (int w = 0; w < W; w++) {
        uint64_t m = kmer[w] >> 2;
        if (w + 1 < W) m |= kmer[w + 1] << 62;
        const int b = mid_bits - 64 * w;
        if (b < 64) m &= b <= 0 ? 0ULL : ((1ULL << b) - 1ULL);
        mid[w] = m;
        uint64_t cand = (kmer[w] >> (top & 63)) & 3ULL;
        BFT_OPAQUE(cand);
        last |= ((top >> 6) == w) ? cand : 0ULL;
    }
    return bft_kf_finish(bft_kf_mix_mid(mid, W), (uint32_t)(kmer[0] & 3ULL) | ((uint32_t)last << 2), n_blocks);
}

/* the filter block at position q.block, and the test of one k-mer's four bits in it */
BFT_HD void bft_kf_load(const bft_view_t* v, uint32_t block, uint64_t* w) {
    const uint64_t* p = v->kfilter + (size_t)block * 4;
#ifdef __CUDA_ARCH__
    /* one 32-byte sector, kept in L2 (evict-last) */
    asm("ld.global.nc.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p));
#else
    w[0] = p[0]; w[1] = p[1]; w[2] = p[2]; w[3] = p[3];
#endif
}
BFT_HD int bft_kf_bits(const uint64_t* w, const bft_kf_pos_t q) {
    return (int)(((w[0] >> q.b0) & (w[1] >> q.b1) & (w[2] >> q.b2) & (w[3] >> q.b3)) & 1ULL);
}

/* 1: the k-mer may be stored; 0: it is certainly not */
BFT_HD int bft_kf_test(const bft_view_t* v, const uint64_t* kmer, const int W) {
    const bft_kf_pos_t q = bft_kf_pos(kmer, W, v->k, v->kf_blocks);
    uint64_t w[4];
    bft_kf_load(v, q.block, w);
    return bft_kf_bits(w, q);
}

/* ---- fused root directory + filter ------------------------------------------------------------------------------
 * The sector of a k-mer inside its prefix's group and its three bit positions (one in each filter word), all from one
 * 64-bit hash of the whole k-mer. */
typedef struct { uint32_t j; uint32_t b1, b2, b3; } bft_rkf_pos_t;

BFT_HD bft_rkf_pos_t bft_rkf_pos(const uint64_t* kmer, const int W, const uint32_t n_sectors) {
    uint64_t x = kmer[0];
    for (int w = 1; w < W; w++) x = (x ^ (x >> 29)) * 0x9FB21C651E98DF25ULL + kmer[w];
    x *= 0xD6E8FEB86659FD93ULL;
    x ^= x >> 32;
    x *= 0x9E3779B97F4A7C15ULL;
    bft_rkf_pos_t p;
    p.b1 = (uint32_t)(x >> 58);
    p.b2 = (uint32_t)(x >> 52) & 63u;
    p.b3 = (uint32_t)(x >> 46) & 63u;
    p.j = (uint32_t)((((x >> 14) & 0xffffffffULL) * (uint64_t)n_sectors) >> 32);
    return p;
}

/* index of the first CC of a Node whose Bloom filter fires for idx14, or BFT_FIRSTCC_NONE (src/presenceNode.c:1354-1362) */
BFT_HD uint32_t bft_first_cc(const bft_view_t* v, const bft_node_t* nd, uint32_t idx14) {
    if (!nd->bf_mode) return BFT_LD8(v->firstcc + nd->fc_off + idx14);
    const uint32_t hp = BFT_LD32(v->hpos + nd->hp_off + idx14);
    const uint32_t h1 = hp & 0xffffu, h2 = hp >> 16;
    const uint8_t* bf = v->firstcc + nd->fc_off;
    for (uint32_t i = 0; i < nd->n_cc; i++, bf += nd->bf_stride)
        if ((BFT_LD8(bf + (h1 >> 3)) >> (h1 & 7u)) & (BFT_LD8(bf + (h2 >> 3)) >> (h2 & 7u)) & 1u) return i;
    return BFT_FIRSTCC_NONE;
}

/* One Node probe: the reference's presenceKmer (src/presenceNode.c:1284-1576) on the flattened layout.
 * succ_leaf_quirk != 0 reproduces presenceNeighborsRight at the leaf level (size_kmer == 9,
 * src/presenceNode.c:719-723): the reference clears nucleotide 7 of the prefix (`& 0xfc` on the second byte) before
 * hashing and before forming p_u/p_v, so the CC path answers for the prefix with nuc 7 = A; the Node-UC path
 * (:1164-1208) compares the unmodified k-mer. Only the successor lookups of the branching queries pass it. */
BFT_HD bft_entry_t bft_node_probe_ex(const bft_view_t* v, uint32_t node_id, uint32_t low18, int succ_leaf_quirk, uint32_t* pref_idx) {
    bft_node_t nd;
#ifdef __CUDA_ARCH__
    {
        const uint4 t = __ldg((const uint4*)(v->nodes + node_id));
        const uint4 u = __ldg((const uint4*)(v->nodes + node_id) + 1);
        nd.cc_begin = t.x; nd.n_cc = t.y; nd.fc_off = t.z; nd.uc_begin = t.w;
        nd.uc_n = u.x; nd.bf_mode = u.y; nd.hp_off = u.z; nd.bf_stride = u.w;
    }
#else
    nd = v->nodes[node_id];
#endif
    uint32_t r18 = bft_msb_first18(low18);
    if (succ_leaf_quirk) r18 &= ~0xcu; /* nuc 7 sits at bits 2-3 of the MSB-first prefix */
    if (nd.n_cc) {
        const uint32_t c = bft_first_cc(v, &nd, bft_idx14(r18));
        if (c != BFT_FIRSTCC_NONE) {
            bft_cc_t cc;
#ifdef __CUDA_ARCH__
            {
                const uint4 t = __ldg((const uint4*)(v->ccs + nd.cc_begin + c));
                cc.csr_off = t.x; cc.f3_off = t.y; cc.pref_off = t.z;
                cc.nb_elem = (uint16_t)(t.w & 0xffffu); cc.s = (uint8_t)((t.w >> 16) & 0xffu); cc.pad = 0;
            }
#else
            cc = v->ccs[nd.cc_begin + c];
#endif
            const uint32_t rot = bft_rot18(r18);
            const uint32_t pu = rot >> cc.s;
            const uint32_t pv = rot & ((1u << cc.s) - 1u);
            uint32_t lo = BFT_LD16(v->csr + cc.csr_off + pu);
            uint32_t hi = BFT_LD16(v->csr + cc.csr_off + pu + 1);
            if (lo >= hi) return bft_mk_entry(BFT_KIND_ABSENT, 0, 0);
            /* lower_bound of pv in the cluster [lo, hi) of filter3 (src/presenceNode.c:1396-1410 / 1475-1488) */
            const uint8_t* f3 = v->filter3 + cc.f3_off;
            uint32_t end = hi;
            if (cc.s == 8) {
                while (lo < hi) {
                    uint32_t mid = (lo + hi) >> 1;
                    if (BFT_LD8(f3 + mid) < pv) lo = mid + 1; else hi = mid;
                }
                if (lo >= end || BFT_LD8(f3 + lo) != pv) return bft_mk_entry(BFT_KIND_ABSENT, 0, 0);
            } else {
                while (lo < hi) {
                    uint32_t mid = (lo + hi) >> 1;
                    uint32_t t = (mid & 1u) ? (uint32_t)(BFT_LD8(f3 + (mid >> 1)) >> 4) : (uint32_t)(BFT_LD8(f3 + (mid >> 1)) & 0xf);
                    if (t < pv) lo = mid + 1; else hi = mid;
                }
                if (lo >= end) return bft_mk_entry(BFT_KIND_ABSENT, 0, 0);
                uint32_t t = (lo & 1u) ? (uint32_t)(BFT_LD8(f3 + (lo >> 1)) >> 4) : (uint32_t)(BFT_LD8(f3 + (lo >> 1)) & 0xf);
                if (t != pv) return bft_mk_entry(BFT_KIND_ABSENT, 0, 0);
            }
            if (pref_idx) *pref_idx = cc.pref_off + lo;
            return bft_ld_entry(v->pref + cc.pref_off + lo);
        }
    }
    return bft_mk_entry(BFT_KIND_UC, nd.uc_begin, nd.uc_n);
}

BFT_HD bft_entry_t bft_node_probe(const bft_view_t* v, uint32_t node_id, uint32_t low18, int succ_leaf_quirk) {
    return bft_node_probe_ex(v, node_id, low18, succ_leaf_quirk, (uint32_t*)0);
}

/* Search the Node-UC lines [begin, begin+n) for `key` (W words, word W-1 most significant); returns the line index
 * or 0xffffffff (binary_search_UC + equality test, src/UC.c:81-124, src/presenceNode.c:1554-1570). */
BFT_HD uint32_t bft_search_uc(const bft_view_t* v, uint32_t begin, uint32_t n, const uint64_t* key, const int W) {
    uint32_t lo = begin, hi = begin + n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const uint64_t* p = v->uckeys + (size_t)mid * W;
        int less = 0;
        for (int w = W - 1; w >= 0; w--) {
            const uint64_t x = BFT_LD64(p + w);
            if (x != key[w]) { less = x < key[w]; break; }
        }
        if (less) lo = mid + 1; else hi = mid;
    }
    if (lo < begin + n) {
        const uint64_t* p = v->uckeys + (size_t)lo * W;
        int eq = 1;
        for (int w = 0; w < W; w++) eq &= (BFT_LD64(p + w) == key[w]);
        if (eq) return lo;
    }
    return 0xffffffffu;
}

/* Searching one prefix's inline block for `key` (the suffix left after the 9-nt prefix) returns the colour class of the matching
 * line, or BFT_CLS_NONE (binary_search_UC over the block + equality, src/presenceNode.c:1876-1914): a hash of the key names the
 * bucket that holds it. */
BFT_HD uint64_t bft_key_hash(const uint64_t* key, const int W) {
    /* multiplicative hash of the whole suffix: the suffixes of one prefix are near-duplicates of each other in a
     * pan-genome (SNP variants share all but one nucleotide), so raw leading bits would pile them into one bucket */
    uint64_t x = key[0];
    for (int w = 1; w < W; w++) x = (x ^ (x >> 31)) * 0xC2B2AE3D27D4EB4FULL + key[w];
    x ^= x >> 29;
    return x * 0x9E3779B97F4A7C15ULL;
}

/* bucket of a key in a block of 2^lb buckets (collapsed subtrees) */
BFT_HD uint32_t bft_bucket_of(const uint64_t* key, const int W, const uint32_t lb) {
    return lb ? (uint32_t)(bft_key_hash(key, W) >> (64 - lb)) : 0u;
}

/* bucket of a key in a block of nbk buckets, any nbk >= 1 (for a power of two this is the top bits of the hash) */
BFT_HD uint32_t bft_bucket_idx(const uint64_t* key, const int W, const uint32_t nbk) {
    if (nbk <= 1u) return 0u;
#ifdef __CUDA_ARCH__
    return (uint32_t)__umul64hi(bft_key_hash(key, W), (uint64_t)nbk);
#else
    return (uint32_t)(((unsigned __int128)bft_key_hash(key, W) * (unsigned __int128)nbk) >> 64);
#endif
}

/* the four slots of one bucket against `key`: class of the matching slot (and its location), or BFT_CLS_NONE; *last receives the
 * top word of the last slot (a key: the bucket is full; all ones: it has room; otherwise an overflow descriptor) */
BFT_HD uint32_t bft_match_bucket(const bft_view_t* v, size_t bucket, const uint64_t* key, const int W, uint32_t* loc, const int want_cls,
                                 uint64_t* last) {
    uint64_t s[BFT_BUCKET_KEYS * BFT_MAX_WORDS];
    bft_ld_bucket(v->buckets + bucket * (size_t)(BFT_BUCKET_KEYS * W), s, W);
    const int shift = v->cls_shift;
    const uint64_t top_mask = shift ? ((1ULL << shift) - 1ULL) : ~BFT_SLOT_SPECIAL;
    uint32_t found = BFT_CLS_NONE;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int j = 0; j < BFT_BUCKET_KEYS; j++) {
        const uint64_t top = s[j * W + W - 1];
        int eq = !(top & BFT_SLOT_SPECIAL) && (top & top_mask) == key[W - 1];
        for (int w = 0; w < W - 1; w++) eq = eq && s[j * W + w] == key[w];
        if (eq) {
            found = shift ? ((uint32_t)(top >> shift) & v->cls_mask) : (want_cls ? BFT_LD32(v->slotcls + bucket * BFT_BUCKET_KEYS + j) : 0u);
            if (loc) *loc = (uint32_t)(bucket * BFT_BUCKET_KEYS + j);
        }
    }
    *last = s[(BFT_BUCKET_KEYS - 1) * W + W - 1];
    return found;
}

/* Search one prefix's inline block of `nbk` buckets starting at bucket `base`: the bucket the key's hash names; its overflow run
 * if it carries a descriptor; for one-word keys the other half of its 64-byte line if it is full (slots fill from the front, so
 * a key in the last slot means no room and no descriptor). */
BFT_HD uint32_t bft_search_block_ex(const bft_view_t* v, uint32_t base, uint32_t nbk, const uint64_t* key, const int W, uint32_t* loc,
                                    const int want_cls) {
    const uint32_t idx = bft_bucket_idx(key, W, nbk);
    uint64_t last;
    uint32_t found = bft_match_bucket(v, (size_t)base + idx, key, W, loc, want_cls, &last);
    if (found != BFT_CLS_NONE) return found;
    if (last & BFT_SLOT_SPECIAL) {
        if (last == BFT_SLOT_EMPTY) return BFT_CLS_NONE;
        /* overflow run */
        const int shift = v->cls_shift;
        const uint64_t top_mask = shift ? ((1ULL << shift) - 1ULL) : ~BFT_SLOT_SPECIAL;
        const uint32_t start = (uint32_t)last, cnt = (uint32_t)(last >> 32) & 0x7fffffffu;
        for (uint32_t i = 0; i < cnt; i++) {
            const uint64_t* p = v->ovf + ((size_t)start + i) * W;
            const uint64_t top = BFT_LD64(p + W - 1);
            int eq = (top & top_mask) == key[W - 1];
            for (int w = 0; w < W - 1; w++) eq = eq && BFT_LD64(p + w) == key[w];
            if (eq) {
                found = shift ? ((uint32_t)(top >> shift) & v->cls_mask) : (want_cls ? BFT_LD32(v->ovfcls + start + i) : 0u);
                if (loc) *loc = v->loc_ovf + start + i;
            }
        }
        return found;
    }
    if (BFT_PAIRED(W) && nbk > 1u) /* full, no descriptor: the suffix may have moved to the other half of the line */
        found = bft_match_bucket(v, (size_t)base + (idx ^ 1u), key, W, loc, want_cls, &last);
    return found;
}

BFT_HD uint32_t bft_search_block(const bft_view_t* v, uint32_t base, uint32_t nbk, const uint64_t* key, const int W) {
    return bft_search_block_ex(v, base, nbk, key, W, (uint32_t*)0, 1);
}

/* Search a collapsed subtree's block (BFT_KIND_DEEP) for `key` (the k-mer without its first 9 nucleotides): buckets of
 * BFT_BUCKET_KEYS slots, linear probing over the 2^lb buckets of the block; an empty slot ends the search. want_cls == 0: the
 * caller only needs presence (any value other than BFT_CLS_NONE), which spares the slotcls load when ids are not embedded. */
BFT_HD uint32_t bft_search_deep(const bft_view_t* v, uint32_t base, uint32_t lb, const uint64_t* key, const int W, const int want_cls) {
    const uint32_t mask = (1u << lb) - 1u;
    uint32_t b = bft_bucket_of(key, W, lb);
    const int shift = v->cls_shift;
    const uint64_t top_mask = shift ? ((1ULL << shift) - 1ULL) : ~BFT_SLOT_SPECIAL;
    for (uint32_t probe = 0; probe <= mask; probe++) {
        const size_t bucket = (size_t)base + b;
        uint64_t s[BFT_BUCKET_KEYS * BFT_MAX_WORDS];
        bft_ld_bucket(v->dbuckets + bucket * (size_t)(BFT_BUCKET_KEYS * W), s, W);
        uint32_t found = BFT_CLS_NONE;
        int empty = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int j = 0; j < BFT_BUCKET_KEYS; j++) {
            const uint64_t top = s[j * W + W - 1];
            empty |= top == BFT_SLOT_EMPTY;
            int eq = !(top & BFT_SLOT_SPECIAL) && (top & top_mask) == key[W - 1];
            for (int w = 0; w < W - 1; w++) eq = eq && s[j * W + w] == key[w];
            if (eq) found = shift ? ((uint32_t)(top >> shift) & v->cls_mask) : (want_cls ? BFT_LD32(v->dslotcls + bucket * BFT_BUCKET_KEYS + j) : 0u);
        }
        if (found != BFT_CLS_NONE || empty) return found;
        b = (b + 1u) & mask;
    }
    return BFT_CLS_NONE;
}

BFT_HD void bft_shift18(uint64_t* cur, int W) {
    for (int w = 0; w < W; w++) {
        uint64_t x = cur[w] >> BFT_PREFIX_BITS;
        if (w + 1 < W) x |= cur[w + 1] << (64 - BFT_PREFIX_BITS);
        cur[w] = x;
    }
}

/* Full lookup: the reference's isKmerPresent (src/presenceNode.c:1823-1921).
 * kmer: W words (bits above 2k must be zero). Returns the colour class of the k-mer, or BFT_CLS_NONE if absent.
 * W is passed explicitly so device callers can make it a compile-time constant. */
/* st (optional, NULL on the product path): walk statistics for the roofline accounting of SURVEY.md §8(d) —
 * st[0] += Nodes probed, st[1] += sum of ceil(log2(lines+1)) over the line searches, st[2] += 1 if found,
 * st[3] += CCs whose Bloom filter the REFERENCE would probe in those Nodes (index of the first CC that fires + 1, or
 * all of them), st[4] += lines in the searched blocks; and for the arena's own walk: st[5] += bucket / UC searches the
 * product path performs (those the stored-k-mer filter does not cut short), st[6] += k-mers the filter rejects. */
BFT_HD uint32_t bft_cc_probed(const bft_view_t* v, uint32_t node_id, uint32_t low18) {
    const bft_node_t* nd = v->nodes + node_id;
    if (!nd->n_cc) return 0;
    const uint32_t c = bft_first_cc(v, nd, bft_idx14(bft_msb_first18(low18)));
    return c == BFT_FIRSTCC_NONE ? nd->n_cc : c + 1;
}

BFT_HD uint32_t bft_ceil_log2p1(uint32_t n) { /* ceil(log2(n + 1)) */
    uint32_t b = 0;
    while ((1u << b) < n + 1u) b++;
    return b;
}

/* flags of bft_lookup_loc */
#define BFT_LK_SUCC_QUIRK 1   /* reproduce presenceNeighborsRight at the leaf level (see bft_node_probe_ex) */
#define BFT_LK_FILTER_FIRST 2 /* test the stored-k-mer filter BEFORE fetching the root directory entry: for look-ups that mostly
                               * miss (the 8 neighbours of a branching query); otherwise both loads are issued back to back */

#define BFT_LK_NO_FILTER 4    /* skip the stored-k-mer filter: for look-ups known to mostly hit (see k_query_sequences) */
#define BFT_LK_PRESENCE 8     /* the caller only tests the result against BFT_CLS_NONE (branching counts): a found k-mer may come
                               * back with class 0 instead of its own when that spares a load */

/* loc (optional): receives the storage location of the k-mer when it is found (see bft_view_t). */
BFT_HD uint32_t bft_lookup_loc(const bft_view_t* v, const uint64_t* kmer, const int W, const int flags, uint32_t* st, uint32_t* loc) {
    const int succ_leaf_quirk = flags & BFT_LK_SUCC_QUIRK;
    uint64_t cur[BFT_MAX_WORDS];
    for (int w = 0; w < BFT_MAX_WORDS; w++) cur[w] = w < W ? kmer[w] : 0;
    int sz = v->k;
    uint32_t pref_idx = 0;
    bft_entry_t e;
    const int may_filter = !(flags & BFT_LK_NO_FILTER) && (!succ_leaf_quirk || v->kf_quirk_safe);
    const int filtered = v->kf_blocks && may_filter;
    int rejected = 0; /* statistics mode only: the product path stops at a rejection */
    if (filtered && (flags & BFT_LK_FILTER_FIRST)) {
        if (!bft_kf_test(v, kmer, W)) {
            if (!st) return BFT_CLS_NONE;
            rejected = 1;
            st[6]++;
        }
    }
    /* the collapsed view of the root serves plain set membership: not the leaf-level successor quirk (unless the trie has no
     * leaf-level Node, where the quirk never applies), not look-ups that want the storage location, not the statistics mode */
    const int fast = v->rootdir_fast && !loc && !st && (!succ_leaf_quirk || v->kf_quirk_safe);
    const bft_entry_t* const rd = fast ? v->rootdir_fast : v->rootdir;
    /* the fused table carries the entries of the fast view when there is one */
    const int fused = may_filter && v->rkf_sectors && !(flags & BFT_LK_FILTER_FIRST) && !(loc && sz == BFT_NB_CHAR_SUF_PREF) &&
                      (fast || !v->rootdir_fast);
    if (fused) {
        /* ONE L2 sector: the root entry of the prefix and the filter bits of this k-mer */
        const bft_rkf_pos_t q = bft_rkf_pos(kmer, W, v->rkf_sectors);
        const uint64_t* p = v->rootkf + ((size_t)((uint32_t)cur[0] & (BFT_ROOTDIR_SIZE - 1u)) * v->rkf_sectors + q.j) * 4;
        uint64_t w0, w1, w2, w3;
#ifdef __CUDA_ARCH__
        asm("ld.global.nc.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "l"(p));
#else
        w0 = p[0]; w1 = p[1]; w2 = p[2]; w3 = p[3];
#endif
        e.a = (uint32_t)w0;
        e.b = (uint32_t)(w0 >> 32);
        if (!(((w1 >> q.b1) & (w2 >> q.b2) & (w3 >> q.b3)) & 1ULL)) {
            if (!st) return BFT_CLS_NONE;
            rejected = 1;
            st[6]++;
        }
    } else {
        /* a 9-mer trie keeps its k-mers as leaf prefixes of the root: rootdir holds the entry but not its index */
        if (loc && sz == BFT_NB_CHAR_SUF_PREF) e = bft_node_probe_ex(v, 0, (uint32_t)cur[0] & (BFT_ROOTDIR_SIZE - 1u), 0, &pref_idx);
        else e = bft_ld_entry(rd + ((uint32_t)cur[0] & (BFT_ROOTDIR_SIZE - 1u)));
        /* the filter block is fetched right behind the root entry (two independent L2 loads in flight) */
        if (filtered && !(flags & BFT_LK_FILTER_FIRST)) {
            if (!bft_kf_test(v, kmer, W)) {
                if (!st) return BFT_CLS_NONE;
                rejected = 1;
                st[6]++;
            }
        }
    }
    int deep_counted = 0; /* statistics mode: this look-up is one bucket search of a collapsed block on the product path */
    if (st) {
        st[0]++;
        st[3] += bft_cc_probed(v, 0, (uint32_t)cur[0] & (BFT_ROOTDIR_SIZE - 1u));
        if (v->rootdir_fast && (!succ_leaf_quirk || v->kf_quirk_safe) &&
            (bft_ld_entry(v->rootdir_fast + ((uint32_t)cur[0] & (BFT_ROOTDIR_SIZE - 1u))).b >> BFT_KIND_SHIFT) == BFT_KIND_DEEP) {
            deep_counted = 1;
            st[5] += !rejected;
        }
    }
    for (;;) {
        const uint32_t kind = e.b >> BFT_KIND_SHIFT;
        const uint32_t n = kind == BFT_KIND_INLINE ? BFT_INLINE_CNT(e) : (e.b & BFT_CNT_MASK);
        if (kind == BFT_KIND_ABSENT) return BFT_CLS_NONE;
        if (kind == BFT_KIND_DEEP) { /* fast view only (never in statistics mode) */
            bft_shift18(cur, W);
            return bft_search_deep(v, e.a, (e.b >> BFT_LB_SHIFT) & BFT_LB_MASK, cur, W, !(flags & BFT_LK_PRESENCE));
        }
        if (kind == BFT_KIND_LEAF) {
            if (st) st[2]++;
            if (loc) *loc = v->loc_leaf + pref_idx;
            return e.a;
        }
        if (kind == BFT_KIND_UC) {
            if (n == 0) return BFT_CLS_NONE;
            if (st) { st[1] += bft_ceil_log2p1(n); st[4] += n; st[5] += !rejected && !deep_counted; }
            const uint32_t ln = bft_search_uc(v, e.a, n, cur, W);
            if (st && ln != 0xffffffffu) st[2]++;
            if (loc && ln != 0xffffffffu) *loc = v->loc_uc + ln;
            return ln == 0xffffffffu ? BFT_CLS_NONE : BFT_LD32(v->uccls + ln);
        }
        bft_shift18(cur, W);
        sz -= BFT_NB_CHAR_SUF_PREF;
        if (kind == BFT_KIND_INLINE) {
            if (st) { st[1] += bft_ceil_log2p1(n); st[4] += n; st[5] += !rejected && !deep_counted; }
            const uint32_t cls = bft_search_block_ex(v, e.a, BFT_INLINE_NBK(e), cur, W, loc, !(flags & BFT_LK_PRESENCE));
            if (st && cls != BFT_CLS_NONE) st[2]++;
            return cls;
        }
        /* BFT_KIND_NODE */
        if (st) { st[0]++; st[3] += bft_cc_probed(v, e.a, (uint32_t)cur[0] & (BFT_ROOTDIR_SIZE - 1u)); }
        e = bft_node_probe_ex(v, e.a, (uint32_t)cur[0] & (BFT_ROOTDIR_SIZE - 1u), succ_leaf_quirk && sz == BFT_NB_CHAR_SUF_PREF,
                              loc ? &pref_idx : (uint32_t*)0);
    }
}

BFT_HD uint32_t bft_lookup_ex(const bft_view_t* v, const uint64_t* kmer, const int W, const int flags, uint32_t* st) {
    return bft_lookup_loc(v, kmer, W, flags, st, (uint32_t*)0);
}

BFT_HD uint32_t bft_lookup_w(const bft_view_t* v, const uint64_t* kmer, const int W) { return bft_lookup_ex(v, kmer, W, 0, (uint32_t*)0); }

BFT_HD uint32_t bft_lookup(const bft_view_t* v, const uint64_t* kmer) { return bft_lookup_w(v, kmer, v->W); }

#endif /* BFT_ARENA_H */
