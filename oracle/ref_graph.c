/* ref_graph — TEST INFRASTRUCTURE (not product code).
 *
 * Drives the UNMODIFIED reference's graph-traversal snippets (src/snippets.c) on a .bft file:
 *   ref_graph components   file.bft {bfs|dfs} [genome_id ...]   get_nb_connected_component (src/snippets.c:937) with
 *                          BFS/DFS, or BFS_subgraph/DFS_subgraph when genome ids are given; prints "REF_COMPONENTS n"
 *   ref_graph simple_paths file.bft out.txt                     extract_simple_paths_to_disk (src/snippets.c:310)
 *   ref_graph core_paths   file.bft ratio out.txt               extract_simple_core_paths_to_disk (src/snippets.c:572)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "bft.h"
#include "snippets.h"

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: see header of oracle/ref_graph.c\n"); return 2; }
    BFT* g = load_BFT(argv[2]);
    if (strcmp(argv[1], "components") == 0) {
        int n = 0;
        const int dfs = strcmp(argv[3], "dfs") == 0;
        const int nid = argc - 4;
        uint32_t id[4] = {0, 0, 0, 0};
        if (nid > 4) { fprintf(stderr, "at most 4 genome ids\n"); return 2; }
        for (int i = 0; i < nid; i++) id[i] = (uint32_t)atoi(argv[4 + i]);
        if (nid == 0) get_nb_connected_component(g, &n, dfs ? DFS : BFS);
        else if (nid == 1) get_nb_connected_component(g, &n, dfs ? DFS_subgraph : BFS_subgraph, 1, id[0]);
        else if (nid == 2) get_nb_connected_component(g, &n, dfs ? DFS_subgraph : BFS_subgraph, 2, id[0], id[1]);
        else if (nid == 3) get_nb_connected_component(g, &n, dfs ? DFS_subgraph : BFS_subgraph, 3, id[0], id[1], id[2]);
        else get_nb_connected_component(g, &n, dfs ? DFS_subgraph : BFS_subgraph, 4, id[0], id[1], id[2], id[3]);
        printf("REF_COMPONENTS %d\n", n);
    } else if (strcmp(argv[1], "simple_paths") == 0) {
        extract_simple_paths_to_disk(g, argv[3]);
    } else if (strcmp(argv[1], "core_paths") == 0 && argc >= 5) {
        extract_simple_core_paths_to_disk(g, atof(argv[3]), argv[4]);
    } else {
        fprintf(stderr, "unknown mode\n");
        return 2;
    }
    fflush(stdout);
    return 0;
}
