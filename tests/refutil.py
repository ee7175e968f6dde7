"""Test helpers around the oracle: the compiled unmodified reference (oracle/_ref, built from /root/reference by
oracle/Makefile — it travels to the GPU box as a prebuilt binary) and the committed golden fixtures."""
from __future__ import annotations

import os
import re
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

from bloomfiltertrie_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_BFT = os.path.join(REF_DIR, "bft")
REF_HARNESS = os.path.join(REF_DIR, "ref_harness")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def have_ref() -> bool:
    return os.access(REF_BFT, os.X_OK) and os.access(REF_HARNESS, os.X_OK)


def _run(cmd: List[str], cwd: str, timeout: int = 1800) -> str:
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=timeout)
    out = p.stdout.decode(errors="replace")
    if p.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed ({p.returncode}):\n{out[-2000:]}")
    return out


def build_bft(workdir: str, name: str, genome_words: Sequence[np.ndarray], k: int) -> str:
    """Reference `bft build k kmers_comp list out` (src/main.c:160-199) over one k-mer file per genome."""
    d = os.path.join(workdir, name)
    os.makedirs(d, exist_ok=True)
    paths = []
    for i, w in enumerate(genome_words):
        p = os.path.join(d, f"genome_{i:04d}.kc")
        synth.write_kmers_comp(p, w, k)
        paths.append(p)
    lst = os.path.join(d, "genomes.txt")
    with open(lst, "w") as f:
        f.write("\n".join(paths) + "\n")
    out = os.path.join(d, f"{name}.bft")
    _run([REF_BFT, "build", str(k), "kmers_comp", lst, out], cwd=d)
    for p in paths:
        os.remove(p)
    return out


def _write_queries(workdir: str, words: np.ndarray, k: int, tag: str) -> str:
    p = os.path.join(workdir, f"q_{tag}_{os.getpid()}.kc")
    synth.write_kmers_comp(p, words, k)
    return p


def ref_kmers(bft_path: str, words: np.ndarray, k: int, n_genomes: int, threads: int = 4) -> Tuple[np.ndarray, np.ndarray]:
    """(present uint8 [n], rows uint32 [n, RW]) from the reference's isKmerPresent/get_annotation/get_list_id_genomes."""
    d = os.path.dirname(bft_path)
    q = _write_queries(d, words, k, "k")
    out = q + ".out"
    _run([REF_HARNESS, "kmers", bft_path, q, out, str(threads)], cwd=d)
    rw = max(1, (n_genomes + 31) // 32)
    n = len(words)
    raw = np.fromfile(out, dtype=np.uint8)
    present = raw[:n].copy()
    rows = raw[n:n + 4 * n * rw].view(np.uint32).reshape(n, rw).copy()
    os.remove(q)
    os.remove(out)
    return present, rows


def ref_branching(bft_path: str, words: np.ndarray, k: int, threads: int = 4) -> Tuple[np.ndarray, np.ndarray]:
    d = os.path.dirname(bft_path)
    q = _write_queries(d, words, k, "b")
    out = q + ".out"
    _run([REF_HARNESS, "branching", bft_path, q, out, str(threads)], cwd=d)
    n = len(words)
    raw = np.fromfile(out, dtype=np.uint8)
    os.remove(q)
    os.remove(out)
    return raw[:n].copy(), raw[n:2 * n].copy()


def ref_sequences(bft_path: str, seqs: Sequence[bytes], threshold: float, canonical: bool, n_genomes: int,
                  threads: int = 4) -> np.ndarray:
    d = os.path.dirname(bft_path)
    q = os.path.join(d, f"s_{os.getpid()}.txt")
    with open(q, "wb") as f:
        f.write(b"\n".join(seqs) + b"\n")
    out = q + ".out"
    _run([REF_HARNESS, "sequences", bft_path, q, repr(float(threshold)), "canonical" if canonical else "non_canonical",
          out, str(threads)], cwd=d)
    rw = max(1, (n_genomes + 31) // 32)
    rows = np.fromfile(out, dtype=np.uint32).reshape(len(seqs), rw).copy()
    os.remove(q)
    os.remove(out)
    return rows


def ref_cli(bft_path: str, args: List[str], cwd: str) -> str:
    """Run the reference CLI `bft load file ...` in cwd (CSV outputs land there, src/main.c:258-264)."""
    return _run([REF_BFT, "load", bft_path] + args, cwd=cwd)


def parse_count(out: str, what: str) -> int:
    m = re.search(rf"{what} = (\d+)", out)
    return int(m.group(1)) if m else -1


def rows_from_sets(sets: Sequence[set], n_genomes: int) -> np.ndarray:
    rw = max(1, (n_genomes + 31) // 32)
    rows = np.zeros((len(sets), rw), dtype=np.uint32)
    for i, s in enumerate(sets):
        for g in s:
            rows[i, g >> 5] |= np.uint32(1 << (g & 31))
    return rows


# ---- the plain-C restatement (oracle/bft_oracle.c), same conventions as the reference harness -------------------
ORACLE_CLI = os.path.join(ROOT, "oracle", "oracle_cli")


def ensure_oracle() -> str:
    if not os.access(ORACLE_CLI, os.X_OK):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return ORACLE_CLI


def oracle_kmers(bft_path: str, words: np.ndarray, k: int, n_genomes: int, workdir: str):
    q = _write_queries(workdir, words, k, "ok")
    out = q + ".out"
    _run([ensure_oracle(), "kmers", bft_path, q, out], cwd=workdir)
    rw = max(1, (n_genomes + 31) // 32)
    n = len(words)
    raw = np.fromfile(out, dtype=np.uint8)
    os.remove(q)
    os.remove(out)
    return raw[:n].copy(), raw[n:n + 4 * n * rw].view(np.uint32).reshape(n, rw).copy()


def oracle_branching(bft_path: str, words: np.ndarray, k: int, workdir: str):
    q = _write_queries(workdir, words, k, "ob")
    out = q + ".out"
    _run([ensure_oracle(), "branching", bft_path, q, out], cwd=workdir)
    n = len(words)
    raw = np.fromfile(out, dtype=np.uint8)
    os.remove(q)
    os.remove(out)
    return raw[:n].copy(), raw[n:2 * n].copy()


def oracle_sequences(bft_path: str, seqs: Sequence[bytes], threshold: float, canonical: bool, n_genomes: int, workdir: str):
    q = os.path.join(workdir, f"os_{os.getpid()}.txt")
    with open(q, "wb") as f:
        f.write(b"\n".join(seqs) + b"\n")
    out = q + ".out"
    _run([ensure_oracle(), "sequences", bft_path, q, repr(float(threshold)), "canonical" if canonical else "non_canonical", out],
         cwd=workdir)
    rw = max(1, (n_genomes + 31) // 32)
    rows = np.fromfile(out, dtype=np.uint32).reshape(len(seqs), rw).copy()
    os.remove(q)
    os.remove(out)
    return rows


def split_seqs(chars: np.ndarray, offs: np.ndarray) -> List[bytes]:
    b = chars.tobytes()
    return [b[int(offs[i]):int(offs[i + 1])] for i in range(len(offs) - 1)]
