"""ctypes binding of the C-ABI in include/bft_b200.h (libbft_b200.so, built in-tree by __graft_entry__.build()).

This is plumbing for tests and bench.py: NumPy arrays for the host entry points, torch CUDA tensors (by
data_ptr) for the *_device entry points. All query work happens in the CUDA library; there is no Python or
CPU fallback — a missing library or GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbft_b200.so")

SEQ_OK, SEQ_TOO_SHORT, SEQ_BAD_CHAR = 0, 1, 2
CLS_NONE = 0xFFFFFFFF

# every symbol include/bft_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "bft_b200_last_error", "bft_b200_open", "bft_b200_close", "bft_b200_k", "bft_b200_n_genomes",
    "bft_b200_genome_name", "bft_b200_kmer_words", "bft_b200_row_words", "bft_b200_device", "bft_b200_stream",
    "bft_b200_get_stats", "bft_b200_host_alloc", "bft_b200_host_free", "bft_b200_query_kmers",
    "bft_b200_query_kmers_device", "bft_b200_query_kmers_device_counted", "bft_b200_query_kmers_device_accumulate", "bft_b200_query_kmers_ascii", "bft_b200_class_rows", "bft_b200_class_counts",
    "bft_b200_query_sequences", "bft_b200_query_sequences_device", "bft_b200_query_branching",
    "bft_b200_query_branching_device", "bft_b200_query_neighbors", "bft_b200_set_reference_exact_branching",
    "bft_b200_query_kmers_file", "bft_b200_query_branching_file",
    "bft_b200_query_sequences_file", "bft_b200_sync", "bft_b200_launch_count", "bft_b200_kmer_walk_stats_device", "bft_b200_random_gather_probe",
    "bft_b200_extract_kmers", "bft_b200_extract_kmers_device", "bft_b200_extract_kmers_file",
    "bft_b200_device_alloc", "bft_b200_device_free", "bft_b200_peer_export", "bft_b200_peer_import", "bft_b200_peer_close",
    "bft_b200_graph_prepare", "bft_b200_graph_release", "bft_b200_graph_adjacency", "bft_b200_connected_components",
    "bft_b200_simple_paths", "bft_b200_simple_paths_file", "bft_b200_free", "bft_b200_query_vertex_ids",
    "bft_b200_record_bytes", "bft_b200_row_bytes", "bft_b200_query_records", "bft_b200_query_records_device",
    "bft_b200_query_records_compact", "bft_b200_annotation_setop", "bft_b200_annotation_setop_device",
]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_kmers", "n_nodes", "n_ccs", "n_lines", "n_prefixes", "n_classes",
                                           "arena_bytes", "class_row_bytes")] + \
               [(n, C.c_int) for n in ("max_cc_per_node", "max_depth", "n_pools")] + \
               [(n, C.c_double) for n in ("flatten_seconds", "upload_seconds", "decode_seconds")] + \
               [("filter_bytes", C.c_uint64), ("rootkf_bytes", C.c_uint64), ("deep_bytes", C.c_uint64)]


class BFTError(RuntimeError):
    pass


_lib = None


def load_library() -> C.CDLL:
    """Load libbft_b200.so; fails loudly when it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BFTError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(the BFT query engine is CUDA-only and has no fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, sz, u64p, u8p, u32p = C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p
    lib.bft_b200_last_error.restype = C.c_char_p
    lib.bft_b200_open.argtypes = [C.c_char_p, C.c_int, C.POINTER(vp)]
    lib.bft_b200_close.argtypes = [vp]
    lib.bft_b200_close.restype = None
    for f in ("k", "n_genomes", "kmer_words", "row_words", "device"):
        getattr(lib, "bft_b200_" + f).argtypes = [vp]
    lib.bft_b200_genome_name.argtypes = [vp, C.c_int]
    lib.bft_b200_genome_name.restype = C.c_char_p
    lib.bft_b200_stream.argtypes = [vp]
    lib.bft_b200_stream.restype = vp
    lib.bft_b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.bft_b200_host_alloc.argtypes = [sz]
    lib.bft_b200_host_alloc.restype = vp
    lib.bft_b200_host_free.argtypes = [vp]
    lib.bft_b200_host_free.restype = None
    lib.bft_b200_query_kmers.argtypes = [vp, u64p, sz, u8p, u32p, u32p]
    lib.bft_b200_query_kmers_device.argtypes = [vp, u64p, sz, u8p, u32p, u32p]
    lib.bft_b200_query_kmers_device_counted.argtypes = [vp, u64p, sz, u8p, u32p, u64p]
    lib.bft_b200_query_kmers_device_accumulate.argtypes = [vp, u64p, sz, u8p, u32p, u64p]
    lib.bft_b200_query_kmers_ascii.argtypes = [vp, C.c_void_p, sz, u8p, u8p, u32p, u32p]
    lib.bft_b200_class_rows.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    lib.bft_b200_class_counts.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint64)]
    lib.bft_b200_query_sequences.argtypes = [vp, C.c_void_p, u64p, sz, C.c_double, C.c_int, u32p, u8p]
    lib.bft_b200_query_sequences_device.argtypes = [vp, C.c_void_p, u64p, sz, C.c_double, C.c_int, u32p, u8p]
    lib.bft_b200_query_branching.argtypes = [vp, u64p, sz, u8p, u8p, C.POINTER(C.c_uint64)]
    lib.bft_b200_query_branching_device.argtypes = [vp, u64p, sz, u8p, u8p, vp]
    lib.bft_b200_query_neighbors.argtypes = [vp, u64p, sz, u32p]
    lib.bft_b200_set_reference_exact_branching.argtypes = [vp, C.c_int]
    lib.bft_b200_query_kmers_file.argtypes = [vp, C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_uint64)]
    lib.bft_b200_query_branching_file.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_uint64)]
    lib.bft_b200_query_sequences_file.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_double, C.c_int]
    lib.bft_b200_kmer_walk_stats_device.argtypes = [vp, u64p, sz, C.POINTER(C.c_uint64 * 8)]
    lib.bft_b200_random_gather_probe.argtypes = [vp, sz, sz, C.POINTER(C.c_double)]
    lib.bft_b200_extract_kmers.argtypes = [vp, u64p, u32p, u32p, sz, C.POINTER(C.c_uint64)]
    lib.bft_b200_extract_kmers_device.argtypes = [vp, u64p, u32p, sz]
    lib.bft_b200_extract_kmers_file.argtypes = [vp, C.c_char_p, C.c_int]
    lib.bft_b200_device_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    lib.bft_b200_device_free.argtypes = [vp, vp]
    lib.bft_b200_peer_export.argtypes = [vp, vp, C.c_char_p]
    lib.bft_b200_peer_import.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    lib.bft_b200_peer_close.argtypes = [vp, vp]
    lib.bft_b200_record_bytes.argtypes = [vp]
    lib.bft_b200_row_bytes.argtypes = [vp]
    lib.bft_b200_query_records.argtypes = [vp, u8p, sz, u8p, u8p, C.POINTER(C.c_uint64)]
    lib.bft_b200_query_records_device.argtypes = [vp, u8p, sz, u8p, u8p, u64p]
    lib.bft_b200_query_records_compact.argtypes = [vp, u8p, sz, u8p, u8p, C.POINTER(C.c_uint64)]
    lib.bft_b200_annotation_setop.argtypes = [vp, C.c_int, u32p, u64p, sz, u32p, u32p]
    lib.bft_b200_annotation_setop_device.argtypes = [vp, C.c_int, u32p, u64p, sz, u32p, u32p]
    lib.bft_b200_graph_prepare.argtypes = [vp]
    lib.bft_b200_graph_release.argtypes = [vp]
    lib.bft_b200_graph_adjacency.argtypes = [vp, u32p, sz]
    lib.bft_b200_query_vertex_ids.argtypes = [vp, u64p, sz, u32p]
    lib.bft_b200_connected_components.argtypes = [vp, u32p, C.c_int, C.POINTER(C.c_uint64), u32p]
    lib.bft_b200_simple_paths.argtypes = [vp, C.c_double, C.POINTER(vp), C.POINTER(sz), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.bft_b200_simple_paths_file.argtypes = [vp, C.c_double, C.c_char_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.bft_b200_free.argtypes = [vp]
    lib.bft_b200_free.restype = None
    lib.bft_b200_sync.argtypes = [vp]
    lib.bft_b200_launch_count.argtypes = [vp]
    lib.bft_b200_launch_count.restype = C.c_uint64
    _lib = lib
    return lib


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if isinstance(a, int):
        return a              # raw device address (e.g. a peer mapping)
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a.data_ptr())  # torch tensor


class PinnedBuffer:
    """Pinned host memory from bft_b200_host_alloc, viewed as a NumPy array."""

    def __init__(self, shape, dtype):
        self.lib = load_library()
        self.shape = tuple(np.atleast_1d(shape).tolist())
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = self.lib.bft_b200_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise BFTError(self.lib.bft_b200_last_error().decode())
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.bft_b200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class BFTEngine:
    """One .bft file flattened onto one GPU. Mirrors the query subset of the reference's bft.h."""

    def __init__(self, bft_path: str, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.bft_b200_open(os.fsencode(bft_path), device, C.byref(h))
        if rc != 0:
            raise BFTError(f"bft_b200_open({bft_path!r}) = {rc}: {self.lib.bft_b200_last_error().decode()}")
        self.h = h
        self.k = self.lib.bft_b200_k(h)
        self.n_genomes = self.lib.bft_b200_n_genomes(h)
        self.W = self.lib.bft_b200_kmer_words(h)
        self.RW = self.lib.bft_b200_row_words(h)
        self.device = self.lib.bft_b200_device(h)
        self.genome_names = [self.lib.bft_b200_genome_name(h, i).decode() for i in range(self.n_genomes)]

    # -- life cycle
    def close(self):
        if getattr(self, "h", None):
            self.lib.bft_b200_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise BFTError(f"{what} = {rc}: {self.lib.bft_b200_last_error().decode()}")

    @property
    def stream(self) -> int:
        return int(self.lib.bft_b200_stream(self.h) or 0)

    def stats(self) -> dict:
        s = Stats()
        self._ck(self.lib.bft_b200_get_stats(self.h, C.byref(s)), "bft_b200_get_stats")
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def sync(self):
        self._ck(self.lib.bft_b200_sync(self.h), "bft_b200_sync")

    def launch_count(self) -> int:
        return int(self.lib.bft_b200_launch_count(self.h))

    # -- k-mers
    def query_kmers(self, kmers: np.ndarray, want_rows: bool = True, want_classes: bool = False,
                    out_present: Optional[np.ndarray] = None, out_rows: Optional[np.ndarray] = None,
                    out_classes: Optional[np.ndarray] = None):
        """kmers: uint64 [n, W] (host). Returns (present uint8 [n], rows uint32 [n, RW] | None, classes | None)."""
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64).reshape(-1, self.W)
        n = kmers.shape[0]
        present = out_present if out_present is not None else np.empty(n, dtype=np.uint8)
        rows = out_rows if out_rows is not None else (np.empty((n, self.RW), dtype=np.uint32) if want_rows else None)
        cls = out_classes if out_classes is not None else (np.empty(n, dtype=np.uint32) if want_classes else None)
        self._ck(self.lib.bft_b200_query_kmers(self.h, _ptr(kmers), n, _ptr(present), _ptr(rows), _ptr(cls)),
                 "bft_b200_query_kmers")
        return present, rows, cls

    def query_records(self, records: np.ndarray, want_present: bool = True, out_rows: Optional[np.ndarray] = None,
                      out_present: Optional[np.ndarray] = None):
        """The reference's record format: uint8 [n, ceil(2k/8)] in -> (present uint8 [n] | None, rows uint8 [n, ceil(G/8)],
        number of k-mers present)."""
        nb, rb = self.lib.bft_b200_record_bytes(self.h), self.lib.bft_b200_row_bytes(self.h)
        records = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1, nb)
        n = records.shape[0]
        rows = out_rows if out_rows is not None else np.empty((n, rb), dtype=np.uint8)
        present = out_present if out_present is not None else (np.empty(n, dtype=np.uint8) if want_present else None)
        cnt = C.c_uint64()
        self._ck(self.lib.bft_b200_query_records(self.h, _ptr(records), n, _ptr(present), _ptr(rows), C.byref(cnt)), "bft_b200_query_records")
        return present, rows, int(cnt.value)

    def query_records_compact(self, records: np.ndarray, out_bits: Optional[np.ndarray] = None, out_rows: Optional[np.ndarray] = None):
        """(presence bits uint8 [(n+7)//8], rows of the present k-mers uint8 [n_present, ceil(G/8)] in query order, n_present)."""
        nb, rb = self.lib.bft_b200_record_bytes(self.h), self.lib.bft_b200_row_bytes(self.h)
        records = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1, nb)
        n = records.shape[0]
        bits = out_bits if out_bits is not None else np.empty((n + 7) // 8, dtype=np.uint8)
        rows = out_rows if out_rows is not None else np.empty((n, rb), dtype=np.uint8)
        cnt = C.c_uint64()
        self._ck(self.lib.bft_b200_query_records_compact(self.h, _ptr(records), n, _ptr(bits), _ptr(rows), C.byref(cnt)),
                 "bft_b200_query_records_compact")
        return bits, rows[:cnt.value], int(cnt.value)

    def query_records_device(self, d_records, n: int, d_present, d_rows, d_n_present=None):
        self._ck(self.lib.bft_b200_query_records_device(self.h, _ptr(d_records), n, _ptr(d_present), _ptr(d_rows), _ptr(d_n_present)),
                 "bft_b200_query_records_device")

    def query_kmers_ascii(self, ascii_kmers: bytes, want_rows: bool = True):
        n = len(ascii_kmers) // self.k
        buf = np.frombuffer(ascii_kmers, dtype=np.uint8)
        valid = np.empty(n, dtype=np.uint8)
        present = np.empty(n, dtype=np.uint8)
        rows = np.empty((n, self.RW), dtype=np.uint32) if want_rows else None
        self._ck(self.lib.bft_b200_query_kmers_ascii(self.h, _ptr(buf), n, _ptr(valid), _ptr(present), _ptr(rows), None),
                 "bft_b200_query_kmers_ascii")
        return valid, present, rows

    def query_kmers_device(self, d_kmers, n: int, d_present=None, d_rows=None, d_classes=None):
        """Device-resident variant: torch CUDA tensors (or raw device addresses); enqueues on self.stream."""
        self._ck(self.lib.bft_b200_query_kmers_device(self.h, _ptr(d_kmers), n, _ptr(d_present), _ptr(d_rows),
                                                      _ptr(d_classes)), "bft_b200_query_kmers_device")

    def query_kmers_device_counted(self, d_kmers, n: int, d_present, d_rows, d_n_present):
        """Device-resident k-mers -> presence, rows and the hit count in d_n_present (int64/uint64 tensor of 1)."""
        self._ck(self.lib.bft_b200_query_kmers_device_counted(self.h, _ptr(d_kmers), n, _ptr(d_present), _ptr(d_rows),
                                                              _ptr(d_n_present)), "bft_b200_query_kmers_device_counted")

    def query_kmers_device_accumulate(self, d_kmers, n: int, d_present, d_rows, d_counter):
        """As query_kmers_device_counted, but ADDS the hit count to d_counter (not zeroed; may be a peer mapping of a
        counter on another GPU — the sharded path's reduction fused into the query kernel)."""
        self._ck(self.lib.bft_b200_query_kmers_device_accumulate(self.h, _ptr(d_kmers), n, _ptr(d_present), _ptr(d_rows),
                                                                 _ptr(d_counter)), "bft_b200_query_kmers_device_accumulate")

    def class_rows(self) -> np.ndarray:
        p = C.c_void_p()
        n = C.c_uint64()
        self._ck(self.lib.bft_b200_class_rows(self.h, C.byref(p), C.byref(n)), "bft_b200_class_rows")
        buf = (C.c_uint32 * (n.value * self.RW)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.uint32).reshape(n.value, self.RW).copy()

    def class_counts(self) -> np.ndarray:
        p = C.c_void_p()
        n = C.c_uint64()
        self._ck(self.lib.bft_b200_class_counts(self.h, C.byref(p), C.byref(n)), "bft_b200_class_counts")
        buf = (C.c_uint32 * n.value).from_address(p.value)
        return np.frombuffer(buf, dtype=np.uint32).copy()

    SET_INTERSECTION, SET_UNION, SET_SYM_DIFFERENCE = 0, 1, 2

    def annotation_setop(self, op: int, class_ids: np.ndarray, group_offs: np.ndarray):
        """intersection / union / symmetric difference of the colour sets of groups of classes (host arrays):
        (rows uint32 [n_groups, RW], genome counts uint32 [n_groups])."""
        class_ids = np.ascontiguousarray(class_ids, dtype=np.uint32)
        group_offs = np.ascontiguousarray(group_offs, dtype=np.uint64)
        n = len(group_offs) - 1
        rows = np.empty((n, self.RW), dtype=np.uint32)
        counts = np.empty(n, dtype=np.uint32)
        self._ck(self.lib.bft_b200_annotation_setop(self.h, int(op), _ptr(class_ids), _ptr(group_offs), n, _ptr(rows), _ptr(counts)),
                 "bft_b200_annotation_setop")
        return rows, counts

    def annotation_setop_device(self, op: int, d_class_ids, d_group_offs, n_groups: int, d_rows=None, d_counts=None):
        self._ck(self.lib.bft_b200_annotation_setop_device(self.h, int(op), _ptr(d_class_ids), _ptr(d_group_offs), n_groups, _ptr(d_rows),
                                                           _ptr(d_counts)), "bft_b200_annotation_setop_device")

    def kmer_walk_stats_device(self, d_kmers, n: int) -> dict:
        out = (C.c_uint64 * 8)()
        self._ck(self.lib.bft_b200_kmer_walk_stats_device(self.h, _ptr(d_kmers), n, C.byref(out)),
                 "bft_b200_kmer_walk_stats_device")
        return dict(nodes=int(out[0]), search_depth=int(out[1]), found=int(out[2]), cc_probed=int(out[3]),
                    block_lines=int(out[4]), bucket_searches=int(out[5]), filter_rejects=int(out[6]), n=n)

    def random_gather_probe(self, table_bytes: int, n_loads: int) -> float:
        out = C.c_double()
        self._ck(self.lib.bft_b200_random_gather_probe(self.h, table_bytes, n_loads, C.byref(out)),
                 "bft_b200_random_gather_probe")
        return float(out.value)

    # -- sequences
    def query_sequences(self, chars: np.ndarray, offs: np.ndarray, threshold: float, canonical: bool,
                        out_rows: Optional[np.ndarray] = None, out_status: Optional[np.ndarray] = None):
        chars = np.ascontiguousarray(chars, dtype=np.uint8)
        offs = np.ascontiguousarray(offs, dtype=np.uint64)
        n = len(offs) - 1
        rows = out_rows if out_rows is not None else np.empty((n, self.RW), dtype=np.uint32)
        status = out_status if out_status is not None else np.empty(n, dtype=np.uint8)
        self._ck(self.lib.bft_b200_query_sequences(self.h, _ptr(chars) if len(chars) else None, _ptr(offs), n,
                                                   float(threshold), int(bool(canonical)), _ptr(rows), _ptr(status)),
                 "bft_b200_query_sequences")
        return rows, status

    def query_sequence_list(self, seqs, threshold: float, canonical: bool):
        chars, offs = pack_sequences(seqs)
        return self.query_sequences(chars, offs, threshold, canonical)

    def query_sequences_device(self, d_chars, d_offs, n: int, threshold: float, canonical: bool, d_rows, d_status=None):
        self._ck(self.lib.bft_b200_query_sequences_device(self.h, _ptr(d_chars), _ptr(d_offs), n, float(threshold),
                                                          int(bool(canonical)), _ptr(d_rows), _ptr(d_status)),
                 "bft_b200_query_sequences_device")

    # -- branching
    def query_branching(self, kmers: np.ndarray, out_succ: Optional[np.ndarray] = None,
                        out_pred: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray, int]:
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64).reshape(-1, self.W)
        n = kmers.shape[0]
        succ = out_succ if out_succ is not None else np.empty(n, dtype=np.uint8)
        pred = out_pred if out_pred is not None else np.empty(n, dtype=np.uint8)
        cnt = C.c_uint64()
        self._ck(self.lib.bft_b200_query_branching(self.h, _ptr(kmers), n, _ptr(succ), _ptr(pred), C.byref(cnt)),
                 "bft_b200_query_branching")
        return succ, pred, int(cnt.value)

    def query_neighbors(self, kmers: np.ndarray) -> np.ndarray:
        """uint32 [n, 8] colour class per neighbour (0xffffffff = absent): 0-3 predecessors, 4-7 successors."""
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64).reshape(-1, self.W)
        out = np.empty((kmers.shape[0], 8), dtype=np.uint32)
        self._ck(self.lib.bft_b200_query_neighbors(self.h, _ptr(kmers), kmers.shape[0], _ptr(out)), "bft_b200_query_neighbors")
        return out

    def set_reference_exact_branching(self, exact: bool):
        self._ck(self.lib.bft_b200_set_reference_exact_branching(self.h, int(bool(exact))),
                 "bft_b200_set_reference_exact_branching")

    def query_branching_device(self, d_kmers, n: int, d_succ=None, d_pred=None, d_count=None):
        self._ck(self.lib.bft_b200_query_branching_device(self.h, _ptr(d_kmers), n, _ptr(d_succ), _ptr(d_pred),
                                                          _ptr(d_count)), "bft_b200_query_branching_device")

    # -- peer (NVLink) result buffers
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._ck(self.lib.bft_b200_device_alloc(self.h, nbytes, C.byref(p)), "bft_b200_device_alloc")
        return int(p.value)

    def device_free(self, ptr: int):
        self._ck(self.lib.bft_b200_device_free(self.h, ptr), "bft_b200_device_free")

    def peer_export(self, ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.lib.bft_b200_peer_export(self.h, ptr, buf), "bft_b200_peer_export")
        return buf.raw

    def peer_import(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._ck(self.lib.bft_b200_peer_import(self.h, handle, C.byref(p)), "bft_b200_peer_import")
        return int(p.value)

    def peer_close(self, ptr: int):
        self._ck(self.lib.bft_b200_peer_close(self.h, ptr), "bft_b200_peer_close")

    def copy_from_device(self, ptr: int, out: np.ndarray):
        """Blocking device -> host copy of a raw device buffer into a NumPy array (test / gather helper)."""
        rt = C.CDLL("libcudart.so.12")
        rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        rc = rt.cudaMemcpy(out.ctypes.data, ptr, out.nbytes, 2)
        if rc != 0:
            raise BFTError(f"cudaMemcpy failed ({rc})")
        return out

    # -- enumeration
    def extract_kmers(self, want_classes: bool = True, want_rows: bool = False):
        """Every stored k-mer: (uint64 [n, W], class ids uint32 [n] | None, rows uint32 [n, RW] | None), arena order."""
        n = int(self.stats()["n_kmers"])
        kmers = np.empty((n, self.W), dtype=np.uint64)
        cls = np.empty(n, dtype=np.uint32) if want_classes else None
        rows = np.empty((n, self.RW), dtype=np.uint32) if want_rows else None
        out = C.c_uint64()
        self._ck(self.lib.bft_b200_extract_kmers(self.h, _ptr(kmers), _ptr(cls), _ptr(rows), n, C.byref(out)),
                 "bft_b200_extract_kmers")
        assert out.value == n
        return kmers, cls, rows

    def extract_kmers_file(self, path: str, compressed_output: bool = True):
        self._ck(self.lib.bft_b200_extract_kmers_file(self.h, os.fsencode(path), int(bool(compressed_output))),
                 "bft_b200_extract_kmers_file")

    # -- graph traversals (reference src/snippets.c)
    def graph_prepare(self):
        self._ck(self.lib.bft_b200_graph_prepare(self.h), "bft_b200_graph_prepare")

    def graph_release(self):
        self._ck(self.lib.bft_b200_graph_release(self.h), "bft_b200_graph_release")

    def query_vertex_ids(self, kmers: np.ndarray) -> np.ndarray:
        """Index of each k-mer in the order of extract_kmers (0xffffffff if absent): the key for caller-side marks."""
        kmers = np.ascontiguousarray(kmers, dtype=np.uint64).reshape(-1, self.W)
        out = np.empty(kmers.shape[0], dtype=np.uint32)
        self._ck(self.lib.bft_b200_query_vertex_ids(self.h, _ptr(kmers), kmers.shape[0], _ptr(out)), "bft_b200_query_vertex_ids")
        return out

    def graph_adjacency(self) -> np.ndarray:
        """uint32 [n_kmers, 8]: vertex index of each possible neighbour (0-3 predecessors, 4-7 successors) or 0xffffffff."""
        n = int(self.stats()["n_kmers"])
        adj = np.empty((n, 8), dtype=np.uint32)
        self._ck(self.lib.bft_b200_graph_adjacency(self.h, _ptr(adj), n), "bft_b200_graph_adjacency")
        return adj

    def connected_components(self, genome_ids=(), want_labels: bool = False):
        """Number of connected components of the graph (or of the colour subgraph); optionally the per-k-mer labels."""
        ids = np.asarray(list(genome_ids), dtype=np.uint32)
        n = int(self.stats()["n_kmers"])
        labels = np.empty(n, dtype=np.uint32) if want_labels else None
        out = C.c_uint64()
        self._ck(self.lib.bft_b200_connected_components(self.h, _ptr(ids) if len(ids) else None, len(ids), C.byref(out), _ptr(labels)),
                 "bft_b200_connected_components")
        return (int(out.value), labels) if want_labels else int(out.value)

    def simple_paths_raw(self, core_ratio: float = 0.0, copy: bool = True):
        """(newline-terminated path lines as one bytes object, number of paths, longest length)."""
        buf, nb, n_paths, longest = C.c_void_p(), C.c_size_t(), C.c_uint64(), C.c_uint64()
        self._ck(self.lib.bft_b200_simple_paths(self.h, float(core_ratio), C.byref(buf), C.byref(nb), C.byref(n_paths), C.byref(longest)),
                 "bft_b200_simple_paths")
        try:
            raw = C.string_at(buf.value, nb.value) if (nb.value and copy) else b""
        finally:
            self.lib.bft_b200_free(buf)
        return raw, int(n_paths.value), int(longest.value), int(nb.value)

    def simple_paths(self, core_ratio: float = 0.0):
        """(list of path strings as bytes, longest length). core_ratio 0: all simple paths."""
        raw, n_paths, longest, _ = self.simple_paths_raw(core_ratio)
        lines = raw.split(b"\n")[:-1] if raw else []
        assert len(lines) == n_paths
        return lines, longest

    def simple_paths_file(self, path: str, core_ratio: float = 0.0):
        n_paths, longest = C.c_uint64(), C.c_uint64()
        self._ck(self.lib.bft_b200_simple_paths_file(self.h, float(core_ratio), os.fsencode(path), C.byref(n_paths), C.byref(longest)),
                 "bft_b200_simple_paths_file")
        return int(n_paths.value), int(longest.value)

    # -- file-level drivers
    def query_kmers_file(self, query_path: str, binary: bool, csv_path: str) -> int:
        n = C.c_uint64()
        self._ck(self.lib.bft_b200_query_kmers_file(self.h, os.fsencode(query_path), int(binary), os.fsencode(csv_path),
                                                    C.byref(n)), "bft_b200_query_kmers_file")
        return int(n.value)

    def query_branching_file(self, query_path: str, binary: bool) -> int:
        n = C.c_uint64()
        self._ck(self.lib.bft_b200_query_branching_file(self.h, os.fsencode(query_path), int(binary), C.byref(n)),
                 "bft_b200_query_branching_file")
        return int(n.value)

    def query_sequences_file(self, query_path: str, csv_path: str, threshold: float, canonical: bool):
        self._ck(self.lib.bft_b200_query_sequences_file(self.h, os.fsencode(query_path), os.fsencode(csv_path),
                                                        float(threshold), int(bool(canonical))),
                 "bft_b200_query_sequences_file")


def pack_sequences(seqs) -> Tuple[np.ndarray, np.ndarray]:
    """List of bytes -> (concatenated uint8 chars, uint64 offsets [n+1])."""
    offs = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        offs[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    chars = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, dtype=np.uint8)
    return chars, offs


def rows_to_bool(rows: np.ndarray, n_genomes: int) -> np.ndarray:
    """uint32 bitmap rows [n, RW] -> bool [n, n_genomes]."""
    bits = np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little")
    return bits[:, :n_genomes].astype(bool)
