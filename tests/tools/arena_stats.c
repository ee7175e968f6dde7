/* arena_stats — TEST TOOL (host, no GPU): prints the occupancy figures of a flattened arena (buckets, slots in use, overflow
 * lines) and checks the block invariants of the paired layout of one-word keys: a block of more than one bucket has an even
 * number of buckets and starts on a 64-byte line, slots fill from the front, and its lines add up to the declared count. */
#include <stdio.h>
#include <stdlib.h>
#include "bft_flatten.h"

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s file.bft\n", argv[0]); return 2; }
    char err[256];
    bft_arena_t* a = bft_arena_from_file(argv[1], err, sizeof err);
    if (!a) { fprintf(stderr, "%s\n", err); return 1; }
    const int W = a->W;
    size_t blocks = 0, odd_blocks = 0, misaligned = 0, holes = 0, count_mismatch = 0, slots_used = 0, in_ovf = 0, block_buckets = 0;
    for (size_t j = 0; j < a->n_pref; j++) {
        const bft_entry_t e = a->pref[j];
        if ((e.b >> BFT_KIND_SHIFT) != BFT_KIND_INLINE) continue;
        const uint32_t nbk = BFT_INLINE_NBK(e), cnt = BFT_INLINE_CNT(e);
        blocks++;
        block_buckets += nbk;
        if (BFT_PAIRED(W) && nbk > 1) {
            odd_blocks += nbk & 1u;
            misaligned += e.a & 1u;
        }
        uint32_t n = 0;
        for (uint32_t b = 0; b < nbk; b++) {
            int seen_empty = 0;
            for (int s = 0; s < BFT_BUCKET_KEYS; s++) {
                const uint64_t top = a->buckets[(((size_t)e.a + b) * BFT_BUCKET_KEYS + s) * W + W - 1];
                if (top == BFT_SLOT_EMPTY) { seen_empty = 1; continue; }
                if (seen_empty) holes++; /* something after an empty slot */
                if (top & BFT_SLOT_SPECIAL) { const uint32_t m = (uint32_t)(top >> 32) & 0x7fffffffu; n += m; in_ovf += m; }
                else { n++; slots_used++; }
            }
        }
        count_mismatch += n != cnt;
    }
    printf("k=%d W=%d kmers=%zu buckets=%zu block_buckets=%zu blocks=%zu slots_used=%zu in_ovf=%zu n_ovf=%zu odd_blocks=%zu misaligned=%zu holes=%zu "
           "count_mismatch=%zu load_permille=%zu arena_bytes=%zu\n", a->k, W, (size_t)a->n_kmers, (size_t)a->n_buckets, block_buckets, blocks, slots_used, in_ovf,
           (size_t)a->n_ovf, odd_blocks, misaligned, holes, count_mismatch,
           block_buckets ? (slots_used + in_ovf) * 1000 / (block_buckets * BFT_BUCKET_KEYS) : 0, (size_t)bft_arena_bytes(a));
    bft_arena_free(a);
    return 0;
}
