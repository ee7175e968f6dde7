/* bft_xxh64.h — XXH64 written from the published xxHash specification (Yann Collet, xxHash64 algorithm
 * description). The reference vendors xxHash v0.6.2 and uses XXH64 only to fill the Bloom-filter hash table
 * hash_v (reference include/Node.h:158-185): for every 18-bit value i, XXH64(3 bytes of i MSB-first, seed).
 * The .bft file does not store Bloom-filter bits (reference src/write_to_disk.c:228, 656-683), so the flattener
 * must regenerate them with the same hash. */
#ifndef BFT_XXH64_H
#define BFT_XXH64_H
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define BFT_XXH_P1 0x9E3779B185EBCA87ULL
#define BFT_XXH_P2 0xC2B2AE3D27D4EB4FULL
#define BFT_XXH_P3 0x165667B19E3779F9ULL
#define BFT_XXH_P4 0x85EBCA77C2B2AE63ULL
#define BFT_XXH_P5 0x27D4EB2F165667C5ULL

static inline uint64_t bft_xxh_rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t bft_xxh_rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; } /* little-endian hosts */
static inline uint32_t bft_xxh_rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t bft_xxh_round(uint64_t acc, uint64_t in) {
    acc += in * BFT_XXH_P2;
    acc = bft_xxh_rotl(acc, 31);
    return acc * BFT_XXH_P1;
}
static inline uint64_t bft_xxh_merge(uint64_t acc, uint64_t val) {
    acc ^= bft_xxh_round(0, val);
    return acc * BFT_XXH_P1 + BFT_XXH_P4;
}

static inline uint64_t bft_xxh64(const void* data, size_t len, uint64_t seed) {
    const uint8_t* p = (const uint8_t*)data;
    const uint8_t* end = p + len;
    uint64_t h;
    if (len >= 32) {
        const uint8_t* limit = end - 32;
        uint64_t v1 = seed + BFT_XXH_P1 + BFT_XXH_P2, v2 = seed + BFT_XXH_P2, v3 = seed, v4 = seed - BFT_XXH_P1;
        do {
            v1 = bft_xxh_round(v1, bft_xxh_rd64(p)); p += 8;
            v2 = bft_xxh_round(v2, bft_xxh_rd64(p)); p += 8;
            v3 = bft_xxh_round(v3, bft_xxh_rd64(p)); p += 8;
            v4 = bft_xxh_round(v4, bft_xxh_rd64(p)); p += 8;
        } while (p <= limit);
        h = bft_xxh_rotl(v1, 1) + bft_xxh_rotl(v2, 7) + bft_xxh_rotl(v3, 12) + bft_xxh_rotl(v4, 18);
        h = bft_xxh_merge(h, v1); h = bft_xxh_merge(h, v2); h = bft_xxh_merge(h, v3); h = bft_xxh_merge(h, v4);
    } else {
        h = seed + BFT_XXH_P5;
    }
    h += (uint64_t)len;
    while (p + 8 <= end) {
        h ^= bft_xxh_round(0, bft_xxh_rd64(p));
        h = bft_xxh_rotl(h, 27) * BFT_XXH_P1 + BFT_XXH_P4;
        p += 8;
    }
    if (p + 4 <= end) {
        h ^= (uint64_t)bft_xxh_rd32(p) * BFT_XXH_P1;
        h = bft_xxh_rotl(h, 23) * BFT_XXH_P2 + BFT_XXH_P3;
        p += 4;
    }
    while (p < end) {
        h ^= (uint64_t)(*p) * BFT_XXH_P5;
        h = bft_xxh_rotl(h, 11) * BFT_XXH_P1;
        p++;
    }
    h ^= h >> 33; h *= BFT_XXH_P2;
    h ^= h >> 29; h *= BFT_XXH_P3;
    h ^= h >> 32;
    return h;
}
#endif
