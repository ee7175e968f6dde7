"""CPU suite, part 1: the oracle. The plain-C restatement (oracle/bft_oracle.c) must reproduce (a) the committed golden
vectors — outputs of the unmodified reference — and (b), when the compiled reference (oracle/_ref) is present, the
reference itself on the larger seeded cases. This is what "parity pinned" in oracle/bft_oracle.h rests on."""
import glob
import os

import numpy as np
import pytest

import cases
import refutil

NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(refutil.GOLDEN, "golden_*.npz")))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name, workdir):
    z = np.load(os.path.join(refutil.GOLDEN, name + ".npz"))
    bft = os.path.join(refutil.GOLDEN, name + ".bft")
    k, G = int(z["k"]), int(z["n_genomes"])
    present, rows = refutil.oracle_kmers(bft, z["queries"], k, G, workdir)
    np.testing.assert_array_equal(present, z["present"])
    np.testing.assert_array_equal(rows, z["rows"])
    succ, pred = refutil.oracle_branching(bft, z["queries"], k, workdir)
    np.testing.assert_array_equal(succ, z["succ"])
    np.testing.assert_array_equal(pred, z["pred"])
    seqs = refutil.split_seqs(z["seq_chars"], z["seq_offs"])
    for canonical in (0, 1):
        for t in z["thresholds"]:
            got = refutil.oracle_sequences(bft, seqs, float(t), bool(canonical), G, workdir)
            np.testing.assert_array_equal(got, z[f"seqrows_c{canonical}_t{t}"], err_msg=f"canonical={canonical} thr={t}")


@pytest.mark.skipif(not refutil.have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("name", ["deep_k45_g4", "structured_k18_g3", "classes_k27_g40", "repeats_k36_g3", "canon_k27_g16"])
def test_oracle_matches_compiled_reference(name, workdir):
    c = cases.make_case(name)
    k, G = c["k"], c["n_genomes"]
    bft = refutil.build_bft(workdir, "o_" + name, c["genome_words"], k)
    assert refutil.oracle_extract_ascii(bft) == refutil.ref_extract_ascii(bft, workdir)   # iterate_over_kmers order
    q = c["queries"][:1500]
    rp, rr = refutil.ref_kmers(bft, q, k, G)
    op, orow = refutil.oracle_kmers(bft, q, k, G, workdir)
    np.testing.assert_array_equal(op, rp)
    np.testing.assert_array_equal(orow, rr)
    rs, rpred = refutil.ref_branching(bft, q[:800], k)
    os_, opred = refutil.oracle_branching(bft, q[:800], k, workdir)
    np.testing.assert_array_equal(os_, rs)
    np.testing.assert_array_equal(opred, rpred)
    seqs = c["seqs"][:60]
    for canonical in (False, True):
        np.testing.assert_array_equal(refutil.oracle_sequences(bft, seqs, 0.8, canonical, G, workdir),
                                      refutil.ref_sequences(bft, seqs, 0.8, canonical, G))


# ---- graph traversals (src/snippets.c): oracle/bft_graph_oracle.c ----------------------------------------------------
import graphutil  # noqa: E402

GRAPH_NAMES = sorted(os.path.basename(p)[len("graph_"):-4] for p in glob.glob(os.path.join(refutil.GOLDEN, "graph_*.npz")))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_enumeration_reproduces_golden(name):
    """o_extract_kmers (iterate_over_kmers restated): the same k-mers in the same order as the reference's own
    `-extract_kmers kmers` output, whose SHA-256 is committed (tests/golden/extract_sha256.json)."""
    import hashlib
    import json
    want = json.load(open(os.path.join(refutil.GOLDEN, "extract_sha256.json")))[name]
    data = refutil.oracle_extract_ascii(os.path.join(refutil.GOLDEN, name + ".bft"))
    k = int(np.load(os.path.join(refutil.GOLDEN, name + ".npz"))["k"])
    assert len(data) == want["n_kmers"] * k
    assert hashlib.sha256(data).hexdigest() == want["sha256"]
    c = cases.make_golden_case(name)           # and they are exactly the k-mers the case inserted
    assert sorted(graphutil.kmer_list(data, k)) == sorted(graphutil.kmer_list(graphutil.case_kmers_ascii(c), k))


@pytest.mark.parametrize("name", GRAPH_NAMES)
def test_graph_oracle_reproduces_golden_bytes(name):
    """With the oracle's own enumeration as the iteration order, the statement-by-statement path extraction must write
    the very bytes the reference wrote (committed in graph_<name>.npz) — no reference needed at run time."""
    z = np.load(os.path.join(refutil.GOLDEN, "graph_" + name + ".npz"))
    bft = os.path.join(refutil.GOLDEN, name + ".bft")
    og = refutil.OracleGraph(bft, refutil.oracle_extract_ascii(bft))
    try:
        assert og.components() == int(z["n_components"])
        for r in z["ratios"]:
            mine, longest = og.simple_paths(float(r), faithful=True)
            assert mine == z[f"paths_r{r}"].tobytes()
            assert longest == int(z[f"longest_r{r}"])
    finally:
        og.close()


@pytest.mark.parametrize("name", GRAPH_NAMES)
def test_graph_oracle_reproduces_golden(name):
    """Component counts equal the reference's; its path lines, with a branching end k-mer dropped (graphutil), are the
    oracle's order-independent paths — the k-mer list here comes from the case generator, not from the reference."""
    z = np.load(os.path.join(refutil.GOLDEN, "graph_" + name + ".npz"))
    c = cases.make_golden_case(name)
    k = c["k"]
    asc = graphutil.case_kmers_ascii(c)
    og = refutil.OracleGraph(os.path.join(refutil.GOLDEN, name + ".bft"), asc)
    try:
        assert og.components() == int(z["n_components"])
        kset = set(graphutil.kmer_list(asc, k))
        for r in z["ratios"]:
            ref_lines = z[f"paths_r{r}"].tobytes().split(b"\n")
            mine, longest = og.simple_paths(float(r), faithful=False)
            assert graphutil.normal_paths(graphutil.trim_branching_ends(ref_lines, kset, k), k) == graphutil.normal_paths(mine.split(b"\n"), k)
            assert longest <= int(z[f"longest_r{r}"]) <= longest + 2
    finally:
        og.close()


@pytest.mark.skipif(not (refutil.have_ref() and refutil.have_ref_graph()), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("name", ["cycles_k27_g3", "cycles_k63_g2", "canon_k45_g3", "repeats_k36_g3"])
def test_graph_oracle_matches_compiled_reference(name, workdir):
    """Same iteration order as the reference (its own -extract_kmers output): the statement-by-statement restatement
    must write the very bytes extract_simple_core_paths_to_disk writes, and count the components BFS and DFS count."""
    c = cases.make_case(name)
    k = c["k"]
    bft = refutil.build_bft(workdir, "og_" + name, c["genome_words"], k)
    asc = refutil.ref_extract_ascii(bft, workdir)
    assert refutil.oracle_extract_ascii(bft) == asc        # the oracle's enumeration, same k-mers in the same order
    assert sorted(graphutil.kmer_list(asc, k)) == sorted(graphutil.kmer_list(graphutil.case_kmers_ascii(c), k))
    og = refutil.OracleGraph(bft, asc)
    try:
        n = og.components()
        assert n == refutil.ref_components(bft, "bfs") == refutil.ref_components(bft, "dfs")
        ref_bytes, ref_longest = refutil.ref_core_paths(bft, 0.0, workdir)
        mine, longest = og.simple_paths(0.0, faithful=True)
        assert mine == ref_bytes and longest == ref_longest
        kset = set(graphutil.kmer_list(asc, k))
        clean, _ = og.simple_paths(0.0, faithful=False)
        assert graphutil.normal_paths(graphutil.trim_branching_ends(ref_bytes.split(b"\n"), kset, k), k) == graphutil.normal_paths(clean.split(b"\n"), k)
    finally:
        og.close()
