"""Seeded synthetic pan-genome data for the BFT query engine.

No datasets are reachable offline, so every test/bench input is generated here: random founder
genomes, SNP/indel-mutated strains (optionally along a random tree), the per-genome k-mer files the
reference `bft build` consumes, and k-mer / read query sets.

K-mer packing follows the reference codec: 2 bits per nucleotide, A=0 C=1 G=2 T=3, nucleotide i in
byte i/4 at bits 2*(i%4) (reference `include/fasta.h:15`, `src/fasta.c:13-23`), i.e. the packed k-mer
read as a little-endian integer is sum(code_i << 2i). The binary k-mer file layout ("kmers_comp") is
two text header lines (k, count) followed by ceil(2k/8)-byte records
(reference `src/file_io.c:132-147`).
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np

ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[_c + 32] = _i  # lowercase
_CODE[ord("U")] = 3
_CODE[ord("u")] = 3


def kmer_nbytes(k: int) -> int:
    return (2 * k + 7) // 8


def kmer_nwords(k: int) -> int:
    """64-bit words per packed k-mer in the engine's convention: 1, 2 or 4 (k <= 32 / 64 / 126)."""
    return 1 if 2 * k <= 64 else (2 if 2 * k <= 128 else 4)


def mask_words(words: np.ndarray, k: int) -> None:
    """Clear, in place, every bit above 2k in [n, nwords] packed k-mers."""
    for w in range(words.shape[1]):
        bits = min(64, max(0, 2 * k - 64 * w))
        if bits < 64:
            words[:, w] &= np.uint64((1 << bits) - 1)


def random_genome(rng: np.random.Generator, length: int) -> np.ndarray:
    """Uniform random genome as 2-bit codes (uint8 array of 0..3)."""
    return rng.integers(0, 4, size=length, dtype=np.uint8)


def mutate(rng: np.random.Generator, g: np.ndarray, snp: float, indel: float = 0.0) -> np.ndarray:
    """Strain of `g`: per-base substitution rate `snp`, per-base indel rate `indel` (length 1..10)."""
    out = g.copy()
    n = len(out)
    if snp > 0:
        pos = np.nonzero(rng.random(n) < snp)[0]
        out[pos] = (out[pos] + rng.integers(1, 4, size=len(pos), dtype=np.uint8)) & 3
    if indel > 0:
        pos = np.nonzero(rng.random(n) < indel)[0]
        if len(pos):
            pieces = []
            last = 0
            for p in pos:
                if p < last:
                    continue
                pieces.append(out[last:p])
                ln = int(rng.integers(1, 11))
                if rng.random() < 0.5:
                    pieces.append(rng.integers(0, 4, size=ln, dtype=np.uint8))  # insertion
                    last = p
                else:
                    last = min(n, p + ln)  # deletion
            pieces.append(out[last:])
            out = np.concatenate(pieces)
    return out


def make_pangenome(n_genomes: int, length: int, snp: float, indel: float = 0.0, seed: int = 12345,
                   tree: bool = False) -> List[np.ndarray]:
    """Founder + (n_genomes-1) strains. tree=False: every strain derives from the founder;
    tree=True: each strain derives from a uniformly chosen earlier genome (random tree, rates per edge)."""
    rng = np.random.default_rng(seed)
    genomes = [random_genome(rng, length)]
    for i in range(1, n_genomes):
        parent = genomes[int(rng.integers(0, i))] if tree else genomes[0]
        genomes.append(mutate(rng, parent, snp, indel))
    return genomes


def pack_windows(codes: np.ndarray, k: int) -> np.ndarray:
    """All k-mer windows of a 2-bit code array as little-endian words: uint64 array [n, nwords]."""
    n = len(codes) - k + 1
    nw = kmer_nwords(k)
    if n <= 0:
        return np.zeros((0, nw), dtype=np.uint64)
    out = np.zeros((n, nw), dtype=np.uint64)
    c64 = codes.astype(np.uint64)
    for j in range(k):
        out[:, j // 32] |= c64[j:j + n] << np.uint64(2 * (j % 32))
    return out


def revcomp_words(words: np.ndarray, k: int) -> np.ndarray:
    """Reverse complement of packed k-mers (uint64 [n, nwords])."""
    n, nw = words.shape
    out = np.zeros_like(words)
    for j in range(k):
        src = (words[:, j // 32] >> np.uint64(2 * (j % 32))) & np.uint64(3)
        d = k - 1 - j
        out[:, d // 32] |= (np.uint64(3) - src) << np.uint64(2 * (d % 32))
    return out


def msb_first_key(words: np.ndarray, k: int) -> np.ndarray:
    """Nucleotide-lexicographic sort key (nuc 0 most significant), same shape as `words`.
    Comparing these as big integers equals strcmp on upper-case ACGT strings (A<C<G<T)."""
    n, nw = words.shape
    out = np.zeros_like(words)
    for j in range(k):
        src = (words[:, j // 32] >> np.uint64(2 * (j % 32))) & np.uint64(3)
        d = k - 1 - j  # position counted from the least significant end
        out[:, d // 32] |= src << np.uint64(2 * (d % 32))
    return out


def canonical_words(words: np.ndarray, k: int) -> np.ndarray:
    """Canonical form under the reference rule (`src/bft.c:1287-1293`): strcmp(fwd, rc) >= 0 -> rc."""
    rc = revcomp_words(words, k)
    kf = msb_first_key(words, k)
    kr = msb_first_key(rc, k)
    nw = words.shape[1]
    use_rc = np.zeros(len(words), dtype=bool)
    decided = np.zeros(len(words), dtype=bool)
    for w in range(nw - 1, -1, -1):
        gt = (kf[:, w] > kr[:, w]) & ~decided
        lt = (kf[:, w] < kr[:, w]) & ~decided
        use_rc |= gt
        decided |= gt | lt
    use_rc |= ~decided  # equal (palindrome): strcmp == 0 -> rc (identical anyway)
    return np.where(use_rc[:, None], rc, words)


def words_to_bytes(words: np.ndarray, k: int) -> np.ndarray:
    """uint64 [n, nwords] -> uint8 [n, ceil(2k/8)] in the reference's byte layout."""
    n, nw = words.shape
    b = np.ascontiguousarray(words.astype("<u8")).view(np.uint8).reshape(n, nw * 8)
    return np.ascontiguousarray(b[:, :kmer_nbytes(k)])


def bytes_to_words(b: np.ndarray, k: int) -> np.ndarray:
    n = b.shape[0]
    nw = kmer_nwords(k)
    buf = np.zeros((n, nw * 8), dtype=np.uint8)
    buf[:, :kmer_nbytes(k)] = b
    return buf.view("<u8").reshape(n, nw).astype(np.uint64)


def words_to_ascii(words: np.ndarray, k: int) -> np.ndarray:
    """uint64 [n, nwords] -> uint8 [n, k] ASCII."""
    n = words.shape[0]
    out = np.empty((n, k), dtype=np.uint8)
    for j in range(k):
        out[:, j] = ALPHABET[((words[:, j // 32] >> np.uint64(2 * (j % 32))) & np.uint64(3)).astype(np.intp)]
    return out


def ascii_to_codes(seq: bytes) -> np.ndarray:
    return _CODE[np.frombuffer(seq, dtype=np.uint8)]


def codes_to_ascii(codes: np.ndarray) -> bytes:
    return ALPHABET[codes].tobytes()


def write_kmers_comp(path: str, words: np.ndarray, k: int) -> None:
    """Binary k-mer file the reference reads with `kmers_comp` (`src/file_io.c:132-147`, 721-774)."""
    with open(path, "wb") as f:
        f.write(f"{k}\n{len(words)}\n".encode())
        f.write(words_to_bytes(words, k).tobytes())


def write_kmers_text(path: str, words: np.ndarray, k: int) -> None:
    a = words_to_ascii(words, k)
    lines = np.empty((len(a), k + 1), dtype=np.uint8)
    lines[:, :k] = a
    lines[:, k] = ord("\n")
    with open(path, "wb") as f:
        f.write(lines.tobytes())


def write_genome_kmer_files(outdir: str, genomes: Sequence[np.ndarray], k: int, canonical: bool = False,
                            binary: bool = True, prefix: str = "genome") -> str:
    """One k-mer file per genome (all windows, optional canonicalisation) + the list file `bft build` wants.
    Returns the list file path. Genome names in the BFT are the basenames (`src/file_io.c:123-125`)."""
    os.makedirs(outdir, exist_ok=True)
    paths = []
    for i, g in enumerate(genomes):
        w = pack_windows(g, k)
        if canonical:
            w = canonical_words(w, k)
        p = os.path.join(outdir, f"{prefix}_{i:04d}.{'kc' if binary else 'txt'}")
        (write_kmers_comp if binary else write_kmers_text)(p, w, k)
        paths.append(p)
    lst = os.path.join(outdir, f"{prefix}_list.txt")
    with open(lst, "w") as f:
        f.write("\n".join(paths) + "\n")
    return lst


def sample_kmer_queries(genomes: Sequence[np.ndarray], k: int, n: int, seed: int,
                        frac_present: float = 0.5, frac_mismatch: float = 0.0) -> np.ndarray:
    """Query k-mers: `frac_present` windows sampled from the genomes, `frac_mismatch` sampled windows with one
    substituted nucleotide (near misses), the rest uniform random k-mers; shuffled. uint64 [n, nwords]."""
    rng = np.random.default_rng(seed)
    n_p = int(n * frac_present)
    n_m = int(n * frac_mismatch)
    n_r = n - n_p - n_m
    nw = kmer_nwords(k)

    def sample(cnt):
        out = np.zeros((cnt, nw), dtype=np.uint64)
        gi = rng.integers(0, len(genomes), size=cnt)
        for g_idx in np.unique(gi):
            sel = np.nonzero(gi == g_idx)[0]
            g = genomes[g_idx]
            pos = rng.integers(0, len(g) - k + 1, size=len(sel))
            c64 = g.astype(np.uint64)
            for j in range(k):
                out[sel, j // 32] |= c64[pos + j] << np.uint64(2 * (j % 32))
        return out

    parts = [sample(n_p)]
    if n_m:
        m = sample(n_m)
        pos = rng.integers(0, k, size=n_m)
        delta = rng.integers(1, 4, size=n_m).astype(np.uint64)
        for w in range(nw):
            in_w = (pos // 32) == w
            sh = (2 * (pos % 32)).astype(np.uint64)
            cur = (m[:, w] >> sh) & np.uint64(3)
            new = (cur + delta) & np.uint64(3)
            m[:, w] = np.where(in_w, (m[:, w] & ~(np.uint64(3) << sh)) | (new << sh), m[:, w])
        parts.append(m)
    if n_r:
        r = rng.integers(0, 1 << 63, size=(n_r, nw), dtype=np.uint64) * np.uint64(2) + \
            rng.integers(0, 2, size=(n_r, nw), dtype=np.uint64)
        mask_words(r, k)
        parts.append(r)
    q = np.concatenate(parts)
    return q[rng.permutation(len(q))]


def sample_reads(genomes: Sequence[np.ndarray], n_reads: int, read_len: int, seed: int,
                 err: float = 0.005, random_strand: bool = True, frac_random: float = 0.0) -> List[bytes]:
    """Synthetic reads (ASCII, upper-case) sampled uniformly from the genomes with substitution errors."""
    rng = np.random.default_rng(seed)
    reads = []
    gi = rng.integers(0, len(genomes), size=n_reads)
    for i in range(n_reads):
        if rng.random() < frac_random:
            codes = rng.integers(0, 4, size=read_len, dtype=np.uint8)
        else:
            g = genomes[gi[i]]
            p = int(rng.integers(0, len(g) - read_len + 1))
            codes = g[p:p + read_len].copy()
            if err > 0:
                e = np.nonzero(rng.random(read_len) < err)[0]
                codes[e] = (codes[e] + rng.integers(1, 4, size=len(e), dtype=np.uint8)) & 3
            if random_strand and rng.random() < 0.5:
                codes = (3 - codes)[::-1]
        reads.append(codes_to_ascii(codes))
    return reads


def kmer_set_oracle(genomes: Sequence[np.ndarray], k: int, canonical: bool = False):
    """Implementation-independent ground truth: dict packed-kmer-bytes -> set of genome ids.
    Only for small inputs (pure Python dict)."""
    table = {}
    for gid, g in enumerate(genomes):
        w = pack_windows(g, k)
        if canonical:
            w = canonical_words(w, k)
        b = words_to_bytes(w, k)
        for row in np.unique(b, axis=0):
            table.setdefault(row.tobytes(), set()).add(gid)
    return table


def deep_kmer_sets(k: int, n_kmers: int, n_genomes: int, seed: int, pool_sizes: Sequence[int] = (40,),
                   membership: float = 0.5) -> Tuple[np.ndarray, List[np.ndarray]]:
    """K-mers that force a deep trie: the i-th 9-nt block of every k-mer is drawn from a pool of pool_sizes[i]
    random blocks (blocks beyond len(pool_sizes) are uniform random), so few prefixes carry many suffixes and the
    reference bursts them into child Nodes. Returns (all distinct k-mers [n, nwords], per-genome
    subsets); each k-mer joins each genome with probability `membership` (at least one genome)."""
    rng = np.random.default_rng(seed)
    nb = k // 9
    codes = rng.integers(0, 4, size=(n_kmers, k), dtype=np.uint8)
    for lvl, ps in enumerate(pool_sizes):
        if lvl >= nb:
            break
        pool = rng.integers(0, 4, size=(ps, 9), dtype=np.uint8)
        codes[:, 9 * lvl:9 * lvl + 9] = pool[rng.integers(0, ps, size=n_kmers)]
    nw = kmer_nwords(k)
    words = np.zeros((n_kmers, nw), dtype=np.uint64)
    for j in range(k):
        words[:, j // 32] |= codes[:, j].astype(np.uint64) << np.uint64(2 * (j % 32))
    words = np.unique(words, axis=0)
    words = words[rng.permutation(len(words))]
    member = rng.random((len(words), n_genomes)) < membership
    none = ~member.any(axis=1)
    member[none, rng.integers(0, n_genomes, size=int(none.sum()))] = True
    return words, [words[member[:, g]] for g in range(n_genomes)]


def near_miss_queries(words: np.ndarray, k: int, n: int, seed: int) -> np.ndarray:
    """Queries around a known k-mer set: 1/3 members, 1/3 members with one nucleotide changed, 1/3 random."""
    rng = np.random.default_rng(seed)
    nw = words.shape[1]
    a = words[rng.integers(0, len(words), size=n // 3)]
    m = words[rng.integers(0, len(words), size=n // 3)].copy()
    pos = rng.integers(0, k, size=len(m))
    delta = rng.integers(1, 4, size=len(m)).astype(np.uint64)
    for w in range(nw):
        in_w = (pos // 32) == w
        sh = (2 * (pos % 32)).astype(np.uint64)
        cur = (m[:, w] >> sh) & np.uint64(3)
        new = (cur + delta) & np.uint64(3)
        m[:, w] = np.where(in_w, (m[:, w] & ~(np.uint64(3) << sh)) | (new << sh), m[:, w])
    n_r = n - len(a) - len(m)
    r = rng.integers(0, 1 << 63, size=(n_r, nw), dtype=np.uint64) * np.uint64(2) + \
        rng.integers(0, 2, size=(n_r, nw), dtype=np.uint64)
    mask_words(r, k)
    q = np.concatenate([a, m, r])
    return q[rng.permutation(len(q))]
