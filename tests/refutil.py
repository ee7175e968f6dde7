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


# ---- graph traversals (reference src/snippets.c) ------------------------------------------------------------------
REF_GRAPH = os.path.join(REF_DIR, "ref_graph")
ORACLE_LIB = os.path.join(ROOT, "oracle", "liboracle_bft.so")


def have_ref_graph() -> bool:
    return os.access(REF_GRAPH, os.X_OK)


def ref_extract_ascii(bft_path: str, workdir: str) -> bytes:
    """Every stored k-mer in the reference's iterate_over_kmers order (`-extract_kmers kmers`), concatenated (n * k chars)."""
    out = os.path.join(workdir, f"extract_{os.getpid()}.txt")
    ref_cli(bft_path, ["-extract_kmers", "kmers", out], cwd=workdir)
    with open(out, "rb") as f:
        data = f.read()
    os.remove(out)
    return data.replace(b"\n", b"")


def ref_components(bft_path: str, mode: str = "bfs", ids: Sequence[int] = ()) -> int:
    out = _run([REF_GRAPH, "components", bft_path, mode] + [str(i) for i in ids], cwd=os.path.dirname(bft_path))
    m = re.search(r"REF_COMPONENTS (\d+)", out)
    return int(m.group(1))


def ref_core_paths(bft_path: str, ratio: float, workdir: str) -> Tuple[bytes, int]:
    """Bytes extract_simple_core_paths_to_disk writes, and the length it prints."""
    out = os.path.join(workdir, f"paths_{os.getpid()}.txt")
    log = _run([REF_GRAPH, "core_paths", bft_path, repr(float(ratio)), out], cwd=workdir)
    with open(out, "rb") as f:
        data = f.read()
    os.remove(out)
    m = re.search(r"Longest simple core path has (\d+) nuc", log)
    return data, int(m.group(1))


_olib = None


def oracle_lib():
    import ctypes as C
    global _olib
    if _olib is None:
        ensure_oracle()
        lib = C.CDLL(ORACLE_LIB)
        lib.o_load.restype = C.c_void_p
        lib.o_load.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        lib.o_free.argtypes = [C.c_void_p]
        lib.o_k.argtypes = [C.c_void_p]
        lib.o_connected_components.restype = C.c_int64
        lib.o_connected_components.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
        lib.o_simple_paths.restype = C.c_void_p
        lib.o_simple_paths.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_double, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
        lib.o_free_buf.argtypes = [C.c_void_p]
        lib.o_extract_kmers.restype = C.c_size_t
        lib.o_extract_kmers.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        _olib = lib
    return _olib


def oracle_extract_ascii(bft_path: str) -> bytes:
    """Every stored k-mer in iterate_over_kmers order from the oracle's own restatement (n * k characters)."""
    import ctypes as C
    lib = oracle_lib()
    err = C.create_string_buffer(256)
    h = lib.o_load(os.fsencode(bft_path), err, 256)
    if not h:
        raise RuntimeError(err.value.decode())
    try:
        k = lib.o_k(h)
        n = lib.o_extract_kmers(h, None, 0)
        buf = C.create_string_buffer(max(1, n * k))
        assert lib.o_extract_kmers(h, buf, n) == n
        return buf.raw[:n * k]
    finally:
        lib.o_free(h)


class OracleGraph:
    """The oracle's sequential restatement of src/snippets.c over a given k-mer list (ASCII, n * k characters)."""

    def __init__(self, bft_path: str, kmers_ascii: bytes):
        import ctypes as C
        self.lib = oracle_lib()
        err = C.create_string_buffer(256)
        self.h = self.lib.o_load(os.fsencode(bft_path), err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.k = self.lib.o_k(self.h)
        self.kmers = kmers_ascii
        self.n = len(kmers_ascii) // self.k

    def close(self):
        if self.h:
            self.lib.o_free(self.h)
            self.h = None

    def components(self, ids: Sequence[int] = (), want_labels: bool = False):
        arr = np.asarray(list(ids), dtype=np.uint32)
        labels = np.empty(self.n, dtype=np.uint32) if want_labels else None
        r = self.lib.o_connected_components(self.h, self.kmers, self.n, arr.ctypes.data if len(arr) else None, len(arr),
                                            labels.ctypes.data if want_labels else None)
        assert r >= 0
        return (int(r), labels) if want_labels else int(r)

    def simple_paths(self, ratio: float, faithful: bool) -> Tuple[bytes, int]:
        import ctypes as C
        nb, longest = C.c_size_t(), C.c_int()
        p = self.lib.o_simple_paths(self.h, self.kmers, self.n, float(ratio), int(faithful), C.byref(nb), C.byref(longest))
        assert p
        data = C.string_at(p, nb.value)
        self.lib.o_free_buf(p)
        return data, int(longest.value)
