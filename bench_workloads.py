"""Benchmark/test infrastructure (NOT part of the product package): the workloads of BASELINE.json's configs
 — synthetic pan-genome BFTs and query batches.

The BFT itself is always built by the UNMODIFIED reference (`oracle/_ref/bft build`, graph construction stays on the
reference host path); this module only generates the seeded inputs, caches the resulting .bft under data/ and
generates query batches with torch (on the GPU for the engine, on the CPU for the bounded reference sample).
"""
from __future__ import annotations

import os
import subprocess
import sys
import time
from typing import List, Tuple

import numpy as np

from bloomfiltertrie_b200 import synth

ROOT = os.path.dirname(os.path.abspath(__file__))
DATA = os.path.join(ROOT, "data")
REF_BFT = os.path.join(ROOT, "oracle", "_ref", "bft")
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")

# config[0]: 4 genomes x 5 Mbp (founder + 3 strains at 1 % SNPs), the reference's own CPU-runnable case
C1 = dict(name="c1_g4", n_genomes=4, snp=0.01, indel=0.0, seed=12345, tree=False)
# config[2] of BASELINE.json: 100-genome synthetic bacterial pan-genome, tree-structured SNP/indel strains
C3 = dict(name="c3_g100", n_genomes=100, snp=0.002, indel=0.0002, seed=12345, tree=True)
# config[4]: 1000-colour pan-genome (annotation-compression heavy: mode-3 annotations + delta-coded colour pools)
C5 = dict(name="c5_g1000", n_genomes=1000, snp=0.002, indel=0.0002, seed=54321, tree=True)
# config[1]: 16-genome pan-genome, canonical k-mers inserted, queried with 150 bp reads (threshold 0.8, canonical)
C2 = dict(name="c2_g16_canon", n_genomes=16, snp=0.005, indel=0.0005, seed=2345, tree=False, canonical=True)


# forced-deep trie (not a BASELINE config): random 63-mers whose first three 9-nt blocks are drawn from pools of
# 120, 20 and 6 blocks, so every 9-nt prefix carries far more than 255 suffixes and the reference bursts it into child Nodes, four
# levels deep: 6 M k-mers, 16 921 Nodes. (Repeat-rich GENOMES of the same size make the reference's own insertion lose 1-3 % of the
# k-mers it stores — it then cannot find them itself — and the serializer refuses such files; this set is one it builds consistently.)
DEEP = dict(name="deep_g8_p120-20-6", n_genomes=8, seed=93, pools=(120, 20, 6), n_kmers=6_000_000)


def log(*a):
    print("[workloads]", *a, file=sys.stderr, flush=True)


def kmer_sets(cfg: dict, k: int):
    """Pool-based k-mer sets of a forced-deep config: (all distinct k-mers [n, W], per-genome subsets)."""
    return synth.deep_kmer_sets(k, cfg["n_kmers"], cfg["n_genomes"], cfg["seed"], pool_sizes=cfg["pools"], membership=0.5)


def pangenome(cfg: dict, genome_len: int) -> List[np.ndarray]:
    return synth.make_pangenome(cfg["n_genomes"], genome_len, cfg["snp"], cfg["indel"], seed=cfg["seed"], tree=cfg["tree"])


def bft_path(cfg: dict, k: int, genome_len: int) -> str:
    return os.path.join(DATA, f"{cfg['name']}_k{k}_L{genome_len}.bft")


def available_lengths(cfg: dict, k: int) -> List[int]:
    out = []
    if os.path.isdir(DATA):
        pre = f"{cfg['name']}_k{k}_L"
        for f in os.listdir(DATA):
            if f.startswith(pre) and (f.endswith(".bft") or f.endswith(".bft.xz")):
                try:
                    out.append(int(f[len(pre):].split(".")[0]))
                except ValueError:
                    pass
    return sorted(set(out))


def ensure_bft(cfg: dict, k: int, genome_len: int, genomes=None) -> str:
    """Path of the cached .bft for (cfg, k, genome_len); builds it with the reference binary when absent."""
    path = bft_path(cfg, k, genome_len)
    if os.path.exists(path):
        return path
    if os.path.exists(path + ".xz"):  # shipped compressed (the snapshot sent to the GPU box is size-limited)
        t0 = time.time()
        tmp_out = f"{path}.{os.getpid()}.tmp"
        try:
            with open(tmp_out, "wb") as f:
                subprocess.run(["xz", "-d", "-c", "-T0", path + ".xz"], stdout=f, check=True)
        except (OSError, subprocess.CalledProcessError):
            import lzma
            with lzma.open(path + ".xz", "rb") as src, open(tmp_out, "wb") as f:
                while True:
                    b = src.read(1 << 24)
                    if not b:
                        break
                    f.write(b)
        os.replace(tmp_out, path)
        log(f"decompressed {os.path.basename(path)}.xz in {time.time() - t0:.1f}s")
        return path
    if not os.access(REF_BFT, os.X_OK):
        raise RuntimeError(f"{path} is absent and the reference binary {REF_BFT} is not built: cannot construct the BFT "
                           "(graph construction is the reference's job)")
    os.makedirs(DATA, exist_ok=True)
    tmp = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"bft_build_{cfg['name']}_k{k}_L{genome_len}_{os.getpid()}")
    os.makedirs(tmp, exist_ok=True)
    t0 = time.time()
    if cfg.get("pools"):  # k-mer sets, not genomes: one kmers_comp file per genome
        _, per = kmer_sets(cfg, k)
        paths = []
        for g, w in enumerate(per):
            paths.append(os.path.join(tmp, f"genome_{g:04d}.kc"))
            synth.write_kmers_comp(paths[-1], w, k)
        lst = os.path.join(tmp, "genome_list.txt")
        with open(lst, "w") as f:
            f.write("\n".join(paths) + "\n")
    else:
        if genomes is None:
            genomes = pangenome(cfg, genome_len)
        lst = synth.write_genome_kmer_files(tmp, genomes, k, canonical=bool(cfg.get("canonical")))
    log(f"building {os.path.basename(path)} with the reference ({cfg['n_genomes']} genomes x {genome_len} bp)...")
    out = os.path.join(tmp, "out.bft")
    p = subprocess.run([REF_BFT, "build", str(k), "kmers_comp", lst, out], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if p.returncode != 0:
        raise RuntimeError("reference build failed:\n" + p.stdout.decode(errors="replace")[-2000:])
    os.replace(out, path)
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)
    log(f"built in {time.time() - t0:.1f}s, {os.path.getsize(path) / 1e6:.1f} MB")
    return path


def genomes_to_torch(genomes: List[np.ndarray], device):
    import torch
    cat = torch.from_numpy(np.concatenate(genomes)).to(device)
    lens = torch.tensor([len(g) for g in genomes], dtype=torch.int64, device=device)
    starts = torch.cumsum(lens, 0) - lens
    return cat, starts, lens


def gen_kmer_queries(cat, starts, lens, k: int, n: int, seed: int, mix: Tuple[float, float, float] = (0.5, 0.25, 0.25)):
    """Query batch on cat.device: mix = (windows present in some genome, windows with one substituted nucleotide,
    uniform random k-mers), shuffled. Returns (int64 [n, W] whose bit pattern is the packed k-mer words,
    uint8 [n] kind: 0 window / 1 mismatch / 2 random)."""
    import torch
    dev = cat.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    W = synth.kmer_nwords(k)
    n_p = int(n * mix[0])
    n_m = int(n * mix[1])
    n_r = n - n_p - n_m
    m = n_p + n_m
    gi = torch.randint(0, len(lens), (m,), generator=g, device=dev)
    span = (lens[gi] - k + 1).to(torch.float64)
    pos = (torch.rand(m, generator=g, device=dev, dtype=torch.float64) * span).to(torch.int64) + starts[gi]
    q = torch.zeros((m, W), dtype=torch.int64, device=dev)
    for j in range(k):
        q[:, j // 32] |= cat[pos + j].to(torch.int64) << (2 * (j % 32))
    if n_m:
        sub = q[n_p:]
        p = torch.randint(0, k, (n_m,), generator=g, device=dev)
        d = torch.randint(1, 4, (n_m,), generator=g, device=dev)
        for w in range(W):
            in_w = (p // 32) == w
            sh = 2 * (p % 32)
            cur = (sub[:, w] >> sh) & 3
            new = (cur + d) & 3
            sub[:, w] = torch.where(in_w, (sub[:, w] & ~(torch.full_like(sh, 3) << sh)) | (new << sh), sub[:, w])
    parts = [q]
    if n_r:
        r = torch.empty((n_r, W), dtype=torch.int64, device=dev)
        for w in range(W):
            bits = max(0, min(64, 2 * k - 64 * w))
            lo = torch.randint(0, 1 << min(32, max(bits, 1)), (n_r,), generator=g, device=dev) if bits > 0 else torch.zeros(n_r, dtype=torch.int64, device=dev)
            hi = torch.randint(0, 1 << min(32, bits - 32), (n_r,), generator=g, device=dev) if bits > 32 else torch.zeros_like(lo)
            r[:, w] = lo | (hi << 32)
        parts.append(r)
    out = torch.cat(parts)
    kind = torch.cat([torch.zeros(n_p, dtype=torch.uint8, device=dev), torch.ones(n_m, dtype=torch.uint8, device=dev),
                      torch.full((n_r,), 2, dtype=torch.uint8, device=dev)])
    perm = torch.randperm(n, generator=g, device=dev)
    return out[perm].contiguous(), kind[perm].contiguous()


def gen_set_queries(words: np.ndarray, k: int, n: int, seed: int, device):
    """Query batch around a k-mer SET (forced-deep configs): 1/3 members, 1/3 members with one nucleotide changed, 1/3 uniform random,
    shuffled. Returns (int64 [n, W] on `device`, uint8 [n] kind: 0 member / 1 mismatch / 2 random)."""
    import torch
    rng = np.random.default_rng(seed)
    nw = words.shape[1]
    third = n // 3
    a = words[rng.integers(0, len(words), size=third)]
    m = words[rng.integers(0, len(words), size=third)].copy()
    pos = rng.integers(0, k, size=third)
    delta = rng.integers(1, 4, size=third).astype(np.uint64)
    for w in range(nw):
        in_w = (pos // 32) == w
        sh = (2 * (pos % 32)).astype(np.uint64)
        cur = (m[:, w] >> sh) & np.uint64(3)
        m[:, w] = np.where(in_w, (m[:, w] & ~(np.uint64(3) << sh)) | (((cur + delta) & np.uint64(3)) << sh), m[:, w])
    r = rng.integers(0, 1 << 63, size=(n - 2 * third, nw), dtype=np.uint64) * np.uint64(2) + \
        rng.integers(0, 2, size=(n - 2 * third, nw), dtype=np.uint64)
    synth.mask_words(r, k)
    q = np.concatenate([a, m, r])
    kind = np.concatenate([np.zeros(third, np.uint8), np.ones(third, np.uint8), np.full(n - 2 * third, 2, np.uint8)])
    perm = rng.permutation(n)
    return torch.from_numpy(q[perm].view(np.int64)).to(device), torch.from_numpy(kind[perm]).to(device)


def gen_reads(cat, starts, lens, n_reads: int, read_len: int, seed: int, err: float = 0.005, frac_random: float = 0.0):
    """Synthetic reads on cat.device: uniform positions over the genomes, random strand, substitution errors.
    Returns (uint8 chars [n_reads * read_len], int64 offsets [n_reads + 1])."""
    import torch
    dev = cat.device
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    gi = torch.randint(0, len(lens), (n_reads,), generator=g, device=dev)
    span = (lens[gi] - read_len + 1).to(torch.float64)
    pos = (torch.rand(n_reads, generator=g, device=dev, dtype=torch.float64) * span).to(torch.int64) + starts[gi]
    idx = pos[:, None] + torch.arange(read_len, device=dev)[None, :]
    codes = cat[idx]
    if frac_random > 0:
        rnd = torch.rand(n_reads, generator=g, device=dev) < frac_random
        codes = torch.where(rnd[:, None], torch.randint(0, 4, codes.shape, generator=g, device=dev, dtype=torch.uint8), codes)
    if err > 0:
        e = torch.rand(codes.shape, generator=g, device=dev) < err
        codes = torch.where(e, (codes + torch.randint(1, 4, codes.shape, generator=g, device=dev, dtype=torch.uint8)) & 3, codes)
    flip = torch.rand(n_reads, generator=g, device=dev) < 0.5
    codes = torch.where(flip[:, None], 3 - codes.flip(1), codes)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    chars = lut[codes.long()].reshape(-1).contiguous()
    offs = (torch.arange(n_reads + 1, device=dev, dtype=torch.int64) * read_len).contiguous()
    return chars, offs
