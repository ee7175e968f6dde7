"""Adds the graph-traversal fixtures graph_<case>.npz next to the golden .bft files, from the UNMODIFIED reference's
src/snippets.c driven by oracle/ref_graph.c (oracle/_ref/ref_graph):

    python tests/golden/make_golden_graph.py

  n_components   get_nb_connected_component(graph, &n, BFS)  (== with DFS; asserted here)
  paths_r<ratio> the bytes extract_simple_core_paths_to_disk(graph, ratio, file) wrote, for the ratios on which the
                 reference runs to completion and its intersection_annotations returns the true intersection
                 (it drops genome ids for some annotation encodings)
The leaf-level fixture (golden_lowcomplex_k18_g3) is left out: the reference's BFS and DFS disagree on it.
Also writes extract_sha256.json (see extract_digests)."""
import os
import sys
import tempfile

import glob
import hashlib
import json

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refutil  # noqa: E402

GRAPH_GOLDEN = {
    "golden_shallow_k27_g4": (0.0, 0.5, 0.9),
    "golden_canon_k27_g8": (0.0,),
    "golden_deep_k63_g12": (),
    "golden_pan_k27_g100": (),
}


def main():
    assert refutil.have_ref_graph(), "build oracle/_ref first (make -C oracle ref)"
    wd = tempfile.mkdtemp()
    for name, ratios in GRAPH_GOLDEN.items():
        bft = os.path.join(HERE, name + ".bft")
        n_bfs, n_dfs = refutil.ref_components(bft, "bfs"), refutil.ref_components(bft, "dfs")
        assert n_bfs == n_dfs
        out = dict(n_components=n_bfs, ratios=np.array(ratios, dtype=np.float64))
        for r in ratios:
            data, longest = refutil.ref_core_paths(bft, r, wd)
            out[f"paths_r{r}"] = np.frombuffer(data, dtype=np.uint8)
            out[f"longest_r{r}"] = longest
        np.savez_compressed(os.path.join(HERE, "graph_" + name + ".npz"), **out)
        print(name, n_bfs, {r: int(out[f"longest_r{r}"]) for r in ratios}, os.path.getsize(os.path.join(HERE, "graph_" + name + ".npz")))


def extract_digests():
    """extract_sha256.json: per golden .bft the number of k-mers and the SHA-256 of the reference's own
    `-extract_kmers kmers` output (k-mers in iterate_over_kmers order, newlines removed)."""
    wd = tempfile.mkdtemp()
    out = {}
    for bft in sorted(glob.glob(os.path.join(HERE, "golden_*.bft"))):
        name = os.path.basename(bft)[:-4]
        data = refutil.ref_extract_ascii(bft, wd)
        k = int(np.load(os.path.join(HERE, name + ".npz"))["k"])
        out[name] = {"n_kmers": len(data) // k, "sha256": hashlib.sha256(data).hexdigest()}
    with open(os.path.join(HERE, "extract_sha256.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(out)


if __name__ == "__main__":
    extract_digests()
    main()
