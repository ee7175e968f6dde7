"""Regenerates the golden fixtures in this directory by running the UNMODIFIED reference (oracle/_ref, compiled
from /root/reference by oracle/Makefile) on the seeded cases in tests/cases.py:

    python tests/golden/make_golden.py

Per case it commits <name>.bft (written by the reference's `bft build`) and <name>.npz holding the query inputs and
the reference's own answers: presence bits + colour rows (isKmerPresent/get_annotation/get_list_id_genomes),
successor/predecessor counts (isBranchingRight/Left) and per-sequence threshold rows (query_sequence) for both
canonical modes. The parity tests replay them without needing the reference at run time."""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cases  # noqa: E402
import refutil  # noqa: E402


def main():
    assert refutil.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    wd = tempfile.mkdtemp()
    for name in cases.GOLDEN_CASES:
        c = cases.make_golden_case(name)
        k, G = c["k"], c["n_genomes"]
        bft = refutil.build_bft(wd, name, c["genome_words"], k)
        shutil.copy(bft, os.path.join(HERE, name + ".bft"))
        present, rows = refutil.ref_kmers(bft, c["queries"], k, G, threads=1)
        succ, pred = refutil.ref_branching(bft, c["queries"], k, threads=1)
        out = dict(k=k, n_genomes=G, queries=c["queries"], present=present, rows=rows, succ=succ, pred=pred,
                   seq_chars=np.frombuffer(b"".join(c["seqs"]), dtype=np.uint8),
                   seq_offs=np.concatenate([[0], np.cumsum([len(s) for s in c["seqs"]])]).astype(np.uint64),
                   thresholds=np.array(cases.GOLDEN_THRESHOLDS))
        for canonical in (0, 1):
            for t in cases.GOLDEN_THRESHOLDS:
                out[f"seqrows_c{canonical}_t{t}"] = refutil.ref_sequences(bft, c["seqs"], t, bool(canonical), G, threads=1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, os.path.getsize(os.path.join(HERE, name + ".bft")), "bytes; present", float(present.mean()))
    shutil.rmtree(wd)


if __name__ == "__main__":
    main()
