"""Seeded parity cases shared by the GPU tests and the golden-fixture generator.

Each case = a k-mer set per genome (what the reference `bft build` inserts) + query k-mers + query sequences,
chosen to cover the layouts SURVEY.md §4 lists: shallow and forced-deep tries, s=8 and s=4 CCs, level_min 0/1
levels, one- and two-word k-mers, the leaf level, 1-byte bit-vector annotations up to compressed (mode 3,
delta-coded pool) annotations, and near-miss queries."""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from bloomfiltertrie_b200 import synth


def _reads_with_edge_cases(genomes, k: int, seed: int, n_reads: int = 300, read_len: int = 150) -> List[bytes]:
    rng = np.random.default_rng(seed)
    reads = synth.sample_reads(genomes, n_reads, read_len, seed, err=0.01, random_strand=True, frac_random=0.1)
    g0 = synth.codes_to_ascii(genomes[0][: 4 * read_len])
    extra = [
        b"",                                   # empty line
        g0[: k - 1],                           # shorter than k
        g0[:k],                                # exactly one window
        g0[: k + 1],                           # two windows
        g0[:read_len].lower(),                 # all lower case
        bytes(c + 32 if i % 3 == 0 else c for i, c in enumerate(g0[:read_len])),   # mixed case
        g0[:read_len].replace(b"T", b"U"),     # U for T
        g0[:60] + b"N" + g0[61:read_len],      # one IUPAC letter: the windows over it are skipped
        g0[:40] + b"RYSWKMBDHVN" + g0[51:read_len],
        b"N" * read_len,                       # nothing but IUPAC
        g0[read_len: 3 * read_len],            # a longer sequence
        g0[: 4 * read_len],                    # spans several tiles? (no: 600 < tile) still long
    ]
    # palindromic windows (canonical tie: strcmp == 0)
    half = g0[: (k + 1) // 2]
    comp = bytes({65: 84, 67: 71, 71: 67, 84: 65}[c] for c in reversed(half))
    extra.append((half + comp)[:k] if k % 2 == 0 else g0[:k])
    order = rng.permutation(len(reads) + len(extra))
    allr = reads + extra
    return [allr[i] for i in order]


def case_shallow(k: int = 27, n_genomes: int = 4, length: int = 60_000, seed: int = 101, canonical: bool = False) -> Dict:
    genomes = synth.make_pangenome(n_genomes, length, 0.01, indel=0.001, seed=seed)
    per = []
    for g in genomes:
        w = synth.pack_windows(g, k)
        per.append(synth.canonical_words(w, k) if canonical else w)
    allw = np.unique(np.concatenate(per), axis=0)
    q = np.concatenate([synth.sample_kmer_queries(genomes, k, 6000, seed + 1, frac_present=0.5, frac_mismatch=0.3),
                        synth.near_miss_queries(allw, k, 3000, seed + 2)])
    return dict(k=k, genome_words=per, queries=q, seqs=_reads_with_edge_cases(genomes, k, seed + 3), canonical=canonical)


def case_deep(k: int, n_genomes: int, n_kmers: int, pools, seed: int, membership: float = 0.5) -> Dict:
    words, per = synth.deep_kmer_sets(k, n_kmers, n_genomes, seed, pool_sizes=pools, membership=membership)
    q = synth.near_miss_queries(words, k, 9000, seed + 1)
    # sequences: chains of overlapping members are rare in random sets; use the k-mers themselves + short joins
    rng = np.random.default_rng(seed + 2)
    asc = synth.words_to_ascii(words[rng.integers(0, len(words), size=120)], k)
    seqs = [bytes(a) for a in asc] + [bytes(a) + bytes(b) for a, b in zip(asc[:40], asc[40:80])]
    return dict(k=k, genome_words=per, queries=q, seqs=seqs, canonical=False)


def case_pangenome(k: int = 27, n_genomes: int = 100, length: int = 8_000, seed: int = 303) -> Dict:
    genomes = synth.make_pangenome(n_genomes, length, 0.002, indel=0.0002, seed=seed, tree=True)
    per = [synth.pack_windows(g, k) for g in genomes]
    q = synth.sample_kmer_queries(genomes, k, 8000, seed + 1, frac_present=0.5, frac_mismatch=0.25)
    return dict(k=k, genome_words=per, queries=q, seqs=_reads_with_edge_cases(genomes, k, seed + 3, n_reads=150), canonical=False)


def case_repeats(k: int, n_genomes: int, vocab: int, n_words: int, seed: int) -> Dict:
    """Genomes spelled from a small vocabulary of 9-mers: a few 9-nt prefixes carry hundreds of suffixes (child
    Nodes, several levels deep) AND consecutive windows are true de Bruijn neighbours, so the deep levels are
    exercised by the branching and sequence queries too."""
    rng = np.random.default_rng(seed)
    words9 = rng.integers(0, 4, size=(vocab, 9), dtype=np.uint8)
    founder = words9[rng.integers(0, vocab, size=n_words)].reshape(-1)
    genomes = [founder] + [synth.mutate(rng, founder, 0.003) for _ in range(n_genomes - 1)]
    per = [synth.pack_windows(g, k) for g in genomes]
    allw = np.unique(np.concatenate(per), axis=0)
    q = np.concatenate([synth.sample_kmer_queries(genomes, k, 6000, seed + 1, frac_present=0.6, frac_mismatch=0.3),
                        synth.near_miss_queries(allw, k, 3000, seed + 2)])
    return dict(k=k, genome_words=per, queries=q, seqs=_reads_with_edge_cases(genomes, k, seed + 3, n_reads=200), canonical=False)


def case_lowcomplexity(k: int, n_genomes: int, length: int, seed: int) -> Dict:
    """Random sequences over the two letters A/C: only 512 distinct 9-nt prefixes, so prefixes overflow into child
    Nodes down to the leaf level, and the de Bruijn graph is dense (most k-mers have several neighbours)."""
    rng = np.random.default_rng(seed)
    genomes = [rng.integers(0, 2, size=length, dtype=np.uint8) for _ in range(n_genomes)]
    per = [synth.pack_windows(g, k) for g in genomes]
    allw = np.unique(np.concatenate(per), axis=0)
    q = np.concatenate([synth.sample_kmer_queries(genomes, k, 5000, seed + 1, frac_present=0.6, frac_mismatch=0.3),
                        synth.near_miss_queries(allw, k, 3000, seed + 2),
                        synth.pack_windows(rng.integers(0, 2, size=3000 + k - 1, dtype=np.uint8), k)])
    return dict(k=k, genome_words=per, queries=q, seqs=_reads_with_edge_cases(genomes, k, seed + 3, n_reads=150), canonical=False)


def case_structured(k: int, n_genomes: int, vocab: int, n_units: int, seed: int) -> Dict:
    """Genomes made of units [one of `vocab` fixed 9-mers][k-9 random nt]: the unit-aligned windows pile thousands
    of four-letter suffixes onto a few prefixes (child Nodes down to the leaf level) while every window keeps true
    de Bruijn neighbours — the layout on which the reference's leaf-level successor rule is visible."""
    rng = np.random.default_rng(seed)
    heads = rng.integers(0, 4, size=(vocab, 9), dtype=np.uint8)
    genomes = []
    for _ in range(n_genomes):
        units = np.concatenate([heads[rng.integers(0, vocab, size=n_units)],
                                rng.integers(0, 4, size=(n_units, k - 9), dtype=np.uint8)], axis=1)
        genomes.append(units.reshape(-1))
    per = [synth.pack_windows(g, k) for g in genomes]
    allw = np.unique(np.concatenate(per), axis=0)
    q = np.concatenate([synth.sample_kmer_queries(genomes, k, 6000, seed + 1, frac_present=0.7, frac_mismatch=0.2),
                        synth.near_miss_queries(allw, k, 3000, seed + 2)])
    return dict(k=k, genome_words=per, queries=q, seqs=_reads_with_edge_cases(genomes, k, seed + 3, n_reads=150), canonical=False)


def case_cycles(k: int, n_genomes: int, seed: int) -> Dict:
    """Graph shapes for the traversal tests: closed loops of non-branching k-mers (circular replicons: every window of
    the circle including the ones across the junction), a loop with a tail, linear pieces with SNP bubbles, a tandem
    repeat, isolated k-mers and a homopolymer self-loop."""
    rng = np.random.default_rng(seed)

    def circ_windows(codes):
        return synth.pack_windows(np.concatenate([codes, codes[: k - 1]]), k)

    circles = [rng.integers(0, 4, size=n, dtype=np.uint8) for n in (k + 5, 300, 1200, 2 * k)]
    base = synth.make_pangenome(n_genomes, 6000, 0.01, indel=0.001, seed=seed + 1)
    unit = rng.integers(0, 4, size=40, dtype=np.uint8)
    tandem = np.concatenate([rng.integers(0, 4, size=100, dtype=np.uint8), np.tile(unit, 6), rng.integers(0, 4, size=100, dtype=np.uint8)])
    lollipop = rng.integers(0, 4, size=500, dtype=np.uint8)
    lollipop = np.concatenate([lollipop, lollipop[200: 200 + k + 30]])   # runs back into itself: a loop with a tail
    homo = np.zeros(k + 3, dtype=np.uint8)                                 # AAAA...: one k-mer, its own neighbour
    singles = [rng.integers(0, 4, size=k, dtype=np.uint8) for _ in range(20)]
    per = []
    for g in range(n_genomes):
        parts = [synth.pack_windows(base[g], k), synth.pack_windows(tandem, k)]
        parts += [circ_windows(c) for i, c in enumerate(circles) if (i + g) % 2 == 0 or g == 0]
        if g != 1:
            parts += [synth.pack_windows(lollipop, k), synth.pack_windows(homo, k)]
        parts += [synth.pack_windows(s, k) for s in singles[g::n_genomes]]
        per.append(np.unique(np.concatenate(parts), axis=0))
    allw = np.unique(np.concatenate(per), axis=0)
    q = np.concatenate([synth.sample_kmer_queries(base, k, 3000, seed + 2, frac_present=0.6, frac_mismatch=0.3),
                        synth.near_miss_queries(allw, k, 2000, seed + 3)])
    return dict(k=k, genome_words=per, queries=q, seqs=_reads_with_edge_cases(base, k, seed + 4, n_reads=80), canonical=False)


CASES = {
    # name: (factory, kwargs)
    "shallow_k27_g4": (case_shallow, dict(k=27, n_genomes=4, seed=101)),
    "canon_k27_g16": (case_shallow, dict(k=27, n_genomes=16, length=30_000, seed=111, canonical=True)),
    "shallow_k63_g5": (case_shallow, dict(k=63, n_genomes=5, length=40_000, seed=121)),
    "canon_k45_g3": (case_shallow, dict(k=45, n_genomes=3, length=30_000, seed=131, canonical=True)),
    "deep_k27_g4": (case_deep, dict(k=27, n_genomes=4, n_kmers=150_000, pools=(40,), seed=201)),
    "deep_k63_g12": (case_deep, dict(k=63, n_genomes=12, n_kmers=150_000, pools=(30, 3, 3, 3, 3), seed=202)),
    "deep_k36_g4": (case_deep, dict(k=36, n_genomes=4, n_kmers=120_000, pools=(30, 4), seed=203)),
    "deep_k45_g4": (case_deep, dict(k=45, n_genomes=4, n_kmers=120_000, pools=(30, 4, 4), seed=204)),
    "deep_k54_g3": (case_deep, dict(k=54, n_genomes=3, n_kmers=120_000, pools=(30, 4, 3, 3), seed=205)),
    "leaf_k18_g4": (case_deep, dict(k=18, n_genomes=4, n_kmers=150_000, pools=(40,), seed=206)),
    "leaf_k9_g5": (case_deep, dict(k=9, n_genomes=5, n_kmers=60_000, pools=(400,), seed=207)),
    "classes_k27_g40": (case_deep, dict(k=27, n_genomes=40, n_kmers=100_000, pools=(60,), seed=208)),
    "sparse_k27_g200": (case_deep, dict(k=27, n_genomes=200, n_kmers=60_000, pools=(60,), seed=209, membership=0.03)),
    "repeats_k27_g4": (case_repeats, dict(k=27, n_genomes=4, vocab=12, n_words=12_000, seed=401)),
    "repeats_k36_g3": (case_repeats, dict(k=36, n_genomes=3, vocab=8, n_words=12_000, seed=402)),
    "lowcomplex_k18_g3": (case_lowcomplexity, dict(k=18, n_genomes=3, length=120_000, seed=404)),
    "structured_k18_g3": (case_structured, dict(k=18, n_genomes=3, vocab=6, n_units=6_000, seed=407)),
    "structured_k27_g2": (case_structured, dict(k=27, n_genomes=2, vocab=3, n_units=5_000, seed=408)),
    "lowcomplex_k27_g3": (case_lowcomplexity, dict(k=27, n_genomes=3, length=100_000, seed=405)),
    "lowcomplex_k45_g2": (case_lowcomplexity, dict(k=45, n_genomes=2, length=100_000, seed=406)),
    "cycles_k27_g3": (case_cycles, dict(k=27, n_genomes=3, seed=601)),
    "cycles_k63_g2": (case_cycles, dict(k=63, n_genomes=2, seed=602)),
    "pan_k27_g100": (case_pangenome, dict(k=27, n_genomes=100, seed=303)),
    "pan_k27_g1000": (case_pangenome, dict(k=27, n_genomes=1000, length=2_000, seed=304)),
    # four-word keys (63 < k <= 126, the reference's KMER_LENGTH_MAX); only inputs on which the reference's own
    # insertion stays consistent (see test_duplicate_kmers_are_refused)
    "shallow_k72_g3": (case_shallow, dict(k=72, n_genomes=3, length=30_000, seed=141)),
    "shallow_k126_g4": (case_shallow, dict(k=126, n_genomes=4, length=25_000, seed=143)),
    "pan_k63_g130": (case_pangenome, dict(k=63, n_genomes=130, length=6_000, seed=305)),
}


def make_case(name: str) -> Dict:
    f, kw = CASES[name]
    c = f(**kw)
    c["name"] = name
    c["n_genomes"] = len(c["genome_words"])
    return c


# Small variants whose reference outputs are committed under tests/golden/ (made by tests/golden/make_golden.py).
GOLDEN_CASES = {
    "golden_shallow_k27_g4": (case_shallow, dict(k=27, n_genomes=4, length=12_000, seed=501)),
    "golden_canon_k27_g8": (case_shallow, dict(k=27, n_genomes=8, length=8_000, seed=502, canonical=True)),
    "golden_deep_k63_g12": (case_deep, dict(k=63, n_genomes=12, n_kmers=30_000, pools=(12, 3, 3, 3, 3), seed=503)),
    "golden_lowcomplex_k18_g3": (case_lowcomplexity, dict(k=18, n_genomes=3, length=60_000, seed=504)),
    "golden_pan_k27_g100": (case_pangenome, dict(k=27, n_genomes=100, length=2_500, seed=505)),
}
GOLDEN_THRESHOLDS = [0.8, 0.5, 1.0]


def make_golden_case(name: str) -> Dict:
    f, kw = GOLDEN_CASES[name]
    c = f(**kw)
    c["name"] = name
    c["n_genomes"] = len(c["genome_words"])
    c["queries"] = c["queries"][:3000]
    c["seqs"] = c["seqs"][:120]
    return c
