/* bft_graph_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see bft_oracle.h).
 *
 * CPU restatement of the reference's graph-traversal snippets (src/snippets.c) on top of the oracle's own k-mer
 * look-up (o_query_kmer): sequential, one k-mer at a time, with a visit mark per k-mer — the reference's control flow,
 * where the product (bloomfiltertrie_b200/csrc/bft_graph.cuh) uses union-find and pointer doubling on a device graph.
 *
 * The caller passes every stored k-mer as ASCII (n * k characters, no separators) in the order iterate_over_kmers
 * would visit them (the order of the reference's `-extract_kmers kmers` output); marks are found by binary search in
 * a sorted copy instead of inside the trie (src/marking.c).
 *
 * Pinned against the unmodified reference (oracle/_ref/ref_graph) by tests/test_oracle.py:
 *   - o_connected_components == get_nb_connected_component(BFS) and (DFS) on tries without leaf-level Nodes
 *     (the reference's successor probe deviates from set membership at the leaf level, and its BFS and DFS then
 *     disagree with each other; the oracle uses set membership);
 *   - o_simple_paths(faithful = 1) == the bytes extract_simple_core_paths_to_disk writes, line for line.
 * Not pinnable: BFS_subgraph / DFS_subgraph abort in the reference (is_in_subgraph frees list_ids twice,
 * src/snippets.c:867,875) and extract_simple_paths exits on an out-of-range index (src/snippets.c:224 reads pred[i]
 * with the successor loop's i); for those the oracle restates the documented behaviour.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "bft_oracle.h"

typedef struct {
    o_bft* b;
    int k, nbytes, G, rw;
    size_t n;
    const char* kmers;  /* n * k, iteration order */
    uint32_t* sorted;   /* indices sorted by k-mer string */
    uint8_t* visited;
    uint32_t* ids;      /* scratch for o_query_kmer */
} graph_t;

static int g_k;
static const char* g_base;
static int cmp_idx(const void* a, const void* b) {
    return memcmp(g_base + (size_t)(*(const uint32_t*)a) * g_k, g_base + (size_t)(*(const uint32_t*)b) * g_k, (size_t)g_k);
}

static int graph_init(graph_t* g, o_bft* b, const char* kmers, size_t n) {
    memset(g, 0, sizeof *g);
    g->b = b; g->k = o_k(b); g->nbytes = (2 * g->k + 7) / 8; g->G = o_n_genomes(b); g->rw = (g->G + 31) / 32 > 0 ? (g->G + 31) / 32 : 1;
    g->n = n; g->kmers = kmers;
    g->sorted = malloc((n + 1) * sizeof(uint32_t));
    g->visited = calloc(n + 1, 1);
    g->ids = malloc(((size_t)g->G + 2) * sizeof(uint32_t));
    if (!g->sorted || !g->visited || !g->ids) return -1;
    for (size_t i = 0; i < n; i++) g->sorted[i] = (uint32_t)i;
    g_k = g->k; g_base = kmers;
    qsort(g->sorted, n, sizeof(uint32_t), cmp_idx);
    return 0;
}
static void graph_done(graph_t* g) { free(g->sorted); free(g->visited); free(g->ids); }

/* index of a k-mer string among the stored k-mers, or -1 (stands in for the reference's marks inside the trie) */
static int64_t find_idx(const graph_t* g, const char* s) {
    size_t lo = 0, hi = g->n;
    while (lo < hi) {
        const size_t mid = lo + (hi - lo) / 2;
        const int c = memcmp(g->kmers + (size_t)g->sorted[mid] * g->k, s, (size_t)g->k);
        if (c == 0) return g->sorted[mid];
        if (c < 0) lo = mid + 1; else hi = mid;
    }
    return -1;
}

static void pack(const graph_t* g, const char* s, uint8_t* out) { /* parseKmerCount layout, src/fasta.c:3-53 */
    memset(out, 0, (size_t)g->nbytes);
    for (int i = 0; i < g->k; i++) {
        const int c = s[i] == 'A' ? 0 : s[i] == 'C' ? 1 : s[i] == 'G' ? 2 : 3;
        out[i >> 2] |= (uint8_t)(c << (2 * (i & 3)));
    }
}

/* is_kmer_in_cdbg of one k-mer string through the oracle's trie walk; fills row (rw words) when present */
static int present(graph_t* g, const char* s, uint32_t* row) {
    uint8_t km[40];
    pack(g, s, km);
    if (!o_query_kmer(g->b, km, g->ids)) return 0;
    if (row) {
        memset(row, 0, (size_t)g->rw * 4);
        for (uint32_t i = 1; i <= g->ids[0]; i++) row[g->ids[i] >> 5] |= 1u << (g->ids[i] & 31);
    }
    return 1;
}

/* get_neighbors order (src/bft.c:804-1003): 0-3 predecessors (A,C,G,T prepended), 4-7 successors (appended).
 * out[j] = index of the neighbour among the stored k-mers or -1. */
static void neighbors(graph_t* g, size_t v, int64_t out[8]) {
    char s[160];
    const char* x = g->kmers + v * g->k;
    for (int j = 0; j < 8; j++) {
        if (j < 4) { s[0] = "ACGT"[j]; memcpy(s + 1, x, (size_t)g->k - 1); }
        else { memcpy(s, x + 1, (size_t)g->k - 1); s[g->k - 1] = "ACGT"[j - 4]; }
        out[j] = present(g, s, NULL) ? find_idx(g, s) : -1;
    }
}

static int row_of(graph_t* g, size_t v, uint32_t* row) { return present(g, g->kmers + v * g->k, row); }

static int in_subgraph(graph_t* g, size_t v, const uint32_t* ids, int n_ids) { /* is_in_subgraph, src/snippets.c:824-881 */
    if (n_ids <= 0) return 0;
    uint32_t row[64];
    if (!row_of(g, v, row)) return 0;
    for (int i = 0; i < n_ids; i++)
        if (ids[i] >= (uint32_t)g->G || !(row[ids[i] >> 5] >> (ids[i] & 31) & 1u)) return 0;
    return 1;
}

/* get_nb_connected_component with BFS (src/snippets.c:605-665, 915-958) or, when n_ids > 0, BFS_subgraph (:667-741).
 * labels (optional, n entries): component number of each k-mer in order of discovery, 0xffffffff if in none. */
int64_t o_connected_components(o_bft* b, const char* kmers, size_t n, const uint32_t* ids, int n_ids, uint32_t* labels) {
    graph_t g;
    if (graph_init(&g, b, kmers, n)) { graph_done(&g); return -1; }
    uint32_t* queue = malloc((n + 1) * sizeof(uint32_t));
    int64_t n_comp = 0;
    if (labels) memset(labels, 0xff, n * sizeof(uint32_t));
    for (size_t s = 0; s < n && queue; s++) {
        if (g.visited[s]) continue;
        g.visited[s] = 1;
        if (n_ids > 0 && !in_subgraph(&g, s, ids, n_ids)) continue;
        size_t head = 0, tail = 0;
        queue[tail++] = (uint32_t)s;
        if (labels) labels[s] = (uint32_t)n_comp;
        while (head < tail) {
            const size_t cur = queue[head++];
            int64_t nb[8];
            neighbors(&g, cur, nb);
            for (int j = 0; j < 8; j++) {
                if (nb[j] < 0 || g.visited[nb[j]]) continue;
                g.visited[nb[j]] = 1;
                if (n_ids > 0 && !in_subgraph(&g, (size_t)nb[j], ids, n_ids)) continue;
                if (labels) labels[nb[j]] = (uint32_t)n_comp;
                queue[tail++] = (uint32_t)nb[j];
            }
        }
        n_comp++;
    }
    free(queue);
    graph_done(&g);
    return n_comp;
}

static int degree(const int64_t* nb4, int* last) {
    int d = 0;
    for (int j = 0; j < 4; j++)
        if (nb4[j] >= 0) { d++; *last = j; }
    return d;
}

static uint32_t shared(const graph_t* g, const uint32_t* a, const uint32_t* b) {
    uint32_t c = 0;
    for (int w = 0; w < g->rw; w++) c += (uint32_t)__builtin_popcount(a[w] & b[w]);
    return c;
}

typedef struct { char* p; size_t len, cap; } buf_t;
static void buf_put(buf_t* o, const char* s, size_t n) {
    if (o->len + n + 1 > o->cap) { o->cap = (o->len + n + 1) * 2; o->p = realloc(o->p, o->cap); }
    memcpy(o->p + o->len, s, n);
    o->len += n;
}

/* extract_simple_core_paths_to_disk (src/snippets.c:346-596); core_ratio 0 gives what extract_simple_paths_to_disk
 * (:115-344) documents. faithful != 0 follows the reference statement by statement, including that the k-mer appended
 * (prepended) in the second and later steps of an extension has only its in-degree (out-degree) tested before it joins
 * the path (:413-455, :492-535), so a line may end (begin) with a branching k-mer depending on where the iteration
 * first entered the path. faithful == 0 applies the test the comments state (:149, :169-171): every k-mer of a path
 * has fewer than two successors and fewer than two predecessors — independent of the iteration order.
 * Returns a malloc'd buffer of '\n'-terminated lines (*n_bytes long), *longest = longest line in characters. */
char* o_simple_paths(o_bft* b, const char* kmers, size_t n, double core_ratio, int faithful, size_t* n_bytes, int* longest) {
    graph_t g;
    buf_t out = {NULL, 0, 0};
    *n_bytes = 0; *longest = 0;
    if (graph_init(&g, b, kmers, n)) { graph_done(&g); return NULL; }
    const int k = g.k;
    const uint32_t core = (uint32_t)(int)(core_ratio * g.G); /* :366 */
    char* path = malloc(2 * n + (size_t)k + 2); /* grows both ways from the middle */
    for (size_t s = 0; s < n && path; s++) {
        if (g.visited[s]) continue;
        g.visited[s] = 1;
        uint32_t seed_row[64], cur_row[64], nb_row[64];
        row_of(&g, s, seed_row);
        if (shared(&g, seed_row, seed_row) < core) continue; /* :376 */
        int64_t nb[8], nb2[8];
        neighbors(&g, s, nb);
        int last = 0;
        const int n_succ = degree(nb + 4, &last), n_pred = degree(nb, &last);
        if (!(n_succ < 2 && n_pred < 2)) continue; /* :389 */
        size_t lo = n, hi = n + (size_t)k; /* path[lo, hi) */
        memcpy(path + lo, kmers + s * k, (size_t)k);
        /* forward (:400-470) */
        memcpy(cur_row, seed_row, sizeof cur_row);
        if (n_succ == 1) {
            degree(nb + 4, &last);
            int64_t cand = nb[4 + last];
            if (!g.visited[cand]) {
                neighbors(&g, (size_t)cand, nb2);
                int l2 = 0;
                int out_deg = degree(nb2 + 4, &l2); /* of the first candidate */
                while (out_deg <= 1 && !g.visited[cand]) {
                    neighbors(&g, (size_t)cand, nb2);
                    if (degree(nb2, &l2) != 1) { g.visited[cand] = 1; break; } /* :458-461 */
                    if (!faithful && degree(nb2 + 4, &l2) > 1) break;
                    row_of(&g, (size_t)cand, nb_row);
                    if (shared(&g, cur_row, nb_row) < core) break;
                    g.visited[cand] = 1;
                    memcpy(cur_row, nb_row, sizeof cur_row);
                    path[hi++] = kmers[(size_t)cand * k + k - 1];
                    out_deg = degree(nb2 + 4, &l2);
                    if (out_deg == 0) break;
                    cand = nb2[4 + l2]; /* the last successor present (:446-455) */
                }
            }
        }
        /* backward (:472-540) */
        memcpy(cur_row, seed_row, sizeof cur_row);
        if (n_pred == 1) {
            degree(nb, &last);
            int64_t cand = nb[last];
            if (!g.visited[cand]) {
                neighbors(&g, (size_t)cand, nb2);
                int l2 = 0;
                int in_deg = degree(nb2, &l2);
                while (in_deg <= 1 && !g.visited[cand]) {
                    neighbors(&g, (size_t)cand, nb2);
                    if (degree(nb2 + 4, &l2) != 1) { g.visited[cand] = 1; break; }
                    if (!faithful && degree(nb2, &l2) > 1) break;
                    row_of(&g, (size_t)cand, nb_row);
                    if (shared(&g, cur_row, nb_row) < core) break;
                    g.visited[cand] = 1;
                    memcpy(cur_row, nb_row, sizeof cur_row);
                    path[--lo] = kmers[(size_t)cand * k];
                    in_deg = degree(nb2, &l2);
                    if (in_deg == 0) break;
                    cand = nb2[l2];
                }
            }
        }
        path[hi] = '\n';
        buf_put(&out, path + lo, hi - lo + 1);
        if ((int)(hi - lo) > *longest) *longest = (int)(hi - lo);
    }
    free(path);
    graph_done(&g);
    if (!out.p) out.p = calloc(1, 1);
    *n_bytes = out.len;
    return out.p;
}

void o_free_buf(void* p) { free(p); }
