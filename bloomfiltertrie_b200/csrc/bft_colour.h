/* bft_colour.h — colour-set codec, read side: annotation bytes -> genome-id bitmap.
 *
 * Restates the reference decoder get_id_genomes_from_annot (src/annotation.c:2086-2250) with decomp_annotation
 * (:1840-1922) and extract_from_annotation_array_elem (include/annotation.h:309-323) for the bytes get_annot hands
 * it (src/UC.c:171-239: annotation, then the optional extended byte appended, src/annotation.c:2117-2124).
 * Output is a bitmap row (bit g of word g/32 = genome g) instead of the reference's ascending id list
 * (src/bft.c:622-641) — the same set. Compiled for the device (k_decode_classes, run once per arena over every
 * distinct annotation) and, by the flattener tests, for the host.
 *
 * First byte, low 2 bits = mode:
 *   0  bit-vector: genome g <=> bit g+2 of the byte string                                 (:2134-2144)
 *   1  ranges: ids as 6-bit big-endian chunks, first chunk flagged 0x1, continuation 0x2;
 *      ids come in (start, stop) pairs, inclusive                                           (:2145-2178)
 *   2  id list: first chunk flagged 0x2, continuation 0x1                                  (:2228-2244)
 *   3  indirection: varint position into the comp_set_colors pools (:2097-2112); the pooled annotation is mode 0,
 *      or mode 1/2 with delta-coded ids (decomp_annotation prefix-sums them, :1877-1916)
 */
#ifndef BFT_COLOUR_H
#define BFT_COLOUR_H

#include "bft_arena.h"

typedef struct {
    int n_pools;
    const int64_t* last_index; /* per pool: last global position it holds */
    const int32_t* size_annot; /* per pool: bytes per entry */
    const uint64_t* off;       /* per pool: byte offset in bytes[] */
    const uint8_t* bytes;
} bft_pools_t;

BFT_HD void bft_row_set(uint32_t* row, int row_words, uint32_t id) {
    if ((id >> 5) < (uint32_t)row_words) row[id >> 5] |= 1u << (id & 31u);
}
BFT_HD void bft_row_set_range(uint32_t* row, int row_words, uint32_t first, uint32_t last) {
    const uint32_t lim = (uint32_t)row_words * 32u;
    if (lim == 0 || first >= lim) return;
    if (last >= lim) last = lim - 1;
    for (uint32_t id = first; id <= last; id++) row[id >> 5] |= 1u << (id & 31u);
}

/* parse one id: first byte carries f_first, continuation bytes carry f_cont; returns 0 when no id starts at *i */
BFT_HD int bft_next_id(const uint8_t* a, int size, int* i, uint32_t f_first, uint32_t f_cont, uint32_t* id) {
    if (*i >= size || !(a[*i] & f_first)) return 0;
    uint32_t v = a[*i] >> 2;
    (*i)++;
    while (*i < size && (a[*i] & f_cont)) {
        v = (v << 6) | (a[*i] >> 2);
        (*i)++;
    }
    *id = v;
    return 1;
}

/* Decode `size` annotation bytes into row[0..row_words) (must be zeroed by the caller).
 * Returns 0, or -1 for a malformed annotation (a mode-3 entry pointing at another mode-3 entry / outside the
 * pools: the reference calls ERROR() there, src/annotation.c:2247). */
BFT_HD int bft_decode_annotation(const uint8_t* annot, int size, const bft_pools_t* pools, uint32_t* row, int row_words) {
    if (size <= 0) return 0;
    int delta = 0;
    uint32_t mode = annot[0] & 3u;
    if (mode == 3) {
        uint32_t position = annot[0] >> 2;
        int i = 1;
        while (i < size && (annot[i] & 1u)) {
            position |= ((uint32_t)(annot[i] >> 1)) << (6 + (i - 1) * 7);
            i++;
        }
        int p = 0;
        while (p < pools->n_pools && (int64_t)position > pools->last_index[p]) p++;
        if (p >= pools->n_pools) return -1;
        const int64_t first = p ? pools->last_index[p - 1] + 1 : 0;
        size = pools->size_annot[p];
        annot = pools->bytes + pools->off[p] + (uint64_t)((int64_t)position - first) * (uint64_t)size;
        if (size <= 0) return 0;
        mode = annot[0] & 3u;
        if (mode == 3) return -1;
        delta = 1;
    }
    if (mode == 0) {
        for (int b = 0; b < size; b++) {
            uint32_t byte = annot[b];
            if (b == 0) byte &= ~3u;
            while (byte) {
                uint32_t bit = 0;
                while (!((byte >> bit) & 1u)) bit++;
                byte &= byte - 1;
                bft_row_set(row, row_words, (uint32_t)(b * 8) + bit - 2u);
            }
        }
        return 0;
    }
    const uint32_t f_first = mode == 1 ? 1u : 2u, f_cont = mode == 1 ? 2u : 1u;
    int i = 0;
    uint32_t id, prev = 0;
    int have_prev = 0;
    if (mode == 2) {
        while (bft_next_id(annot, size, &i, f_first, f_cont, &id)) {
            if (delta && have_prev) id += prev;
            bft_row_set(row, row_words, id);
            prev = id;
            have_prev = 1;
        }
        return 0;
    }
    /* mode 1: (start, stop) pairs; a trailing unpaired start stands alone (:2150-2176, :2193-2222) */
    int is_stop = 0;
    uint32_t start = 0;
    while (bft_next_id(annot, size, &i, f_first, f_cont, &id)) {
        if (delta && have_prev) id += prev;
        prev = id;
        have_prev = 1;
        if (is_stop) {
            if (id >= start) bft_row_set_range(row, row_words, start, id);
        } else {
            start = id;
            bft_row_set(row, row_words, id);
        }
        is_stop = !is_stop;
    }
    return 0;
}

BFT_HD int bft_popc32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

#endif /* BFT_COLOUR_H */
