"""CPU suite, part 1: the oracle. The plain-C restatement (oracle/bft_oracle.c) must reproduce (a) the committed golden
vectors — outputs of the unmodified reference — and (b), when the compiled reference (oracle/_ref) is present, the
reference itself on the larger seeded cases. This is what "parity pinned" in oracle/bft_oracle.h rests on."""
import glob
import os

import numpy as np
import pytest

import cases
import refutil

NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(refutil.GOLDEN, "*.npz")))


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
