"""Helpers of the graph-traversal tests (reference src/snippets.c): order-independent normal forms of the outputs."""
from __future__ import annotations

from collections import Counter
from typing import Dict, Iterable, List

import numpy as np

from bloomfiltertrie_b200 import synth

NUC = (b"A", b"C", b"G", b"T")


def case_kmers_ascii(c: Dict) -> bytes:
    """Every k-mer a case inserts (union over genomes), as n * k ASCII characters."""
    allw = np.unique(np.concatenate(c["genome_words"]), axis=0)
    return synth.words_to_ascii(allw, c["k"]).tobytes()


def kmer_list(ascii_kmers: bytes, k: int) -> List[bytes]:
    return [ascii_kmers[i:i + k] for i in range(0, len(ascii_kmers), k)]


def out_degree(km: bytes, kset) -> int:
    return sum((km[1:] + c) in kset for c in NUC)


def in_degree(km: bytes, kset) -> int:
    return sum((c + km[:-1]) in kset for c in NUC)


def trim_branching_ends(lines: Iterable[bytes], kset, k: int) -> List[bytes]:
    """The reference's extension loops test only the in-degree (out-degree) of the k-mer they append (prepend) from
    the second step on (src/snippets.c:413-455, 492-535), so a line may end (begin) with one branching k-mer depending
    on where iterate_over_kmers entered the path. Dropping such an end k-mer gives the order-independent path."""
    out = []
    for l in lines:
        if not l:
            continue
        if len(l) > k and out_degree(l[-k:], kset) >= 2:
            l = l[:-1]
        if len(l) > k and in_degree(l[:k], kset) >= 2:
            l = l[1:]
        out.append(l)
    return out


def canon_cycle(line: bytes, k: int) -> bytes:
    """A closed loop of non-branching k-mers is written starting at whichever k-mer was reached first; rotate such a
    line to the lexicographically smallest start."""
    m = len(line) - k + 1                     # k-mers on the path
    if m <= 1 or line[-(k - 1):] != line[:k - 1]:
        return line
    circ = line[:m]                           # the loop read once; the line is this string continued for k - 1 more characters
    dbl = circ + circ
    best = min(dbl[i:i + m] for i in range(m))
    return (best * ((m + k - 2) // m + 1))[:m + k - 1]


def normal_paths(lines: Iterable[bytes], k: int) -> Counter:
    return Counter(canon_cycle(l, k) for l in lines if l)


def partition(labels: np.ndarray, kmers: List[bytes], none: int = 0xFFFFFFFF):
    """Set of frozensets of k-mers, one per component label (k-mers labelled `none` are outside the subgraph)."""
    groups: Dict[int, list] = {}
    for l, km in zip(labels.tolist(), kmers):
        if l != none:
            groups.setdefault(l, []).append(km)
    return {frozenset(v) for v in groups.values()}
