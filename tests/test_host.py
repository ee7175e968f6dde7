"""CPU suite, part 2: host logic — codecs, the serializer + arena walk (run on the host by tests/tools), the C-ABI
library's exports, and the multi-process sharding/gather logic under gloo (world_size 2)."""
import ctypes
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

import refutil
import cases
from bloomfiltertrie_b200 import shard, synth

ROOT = refutil.ROOT
CSRC = os.path.join(ROOT, "bloomfiltertrie_b200", "csrc")
NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(refutil.GOLDEN, "golden_*.npz")))


def _gcc(args, **kw):
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    return subprocess.run(["gcc"] + args, env=env, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, **kw)


# ---- codecs ---------------------------------------------------------------------------------------------------
def test_readme_encoding_kat():
    # reference README.md:172: ACTTGTCTG -> 11110100 11011110 00000010
    b = synth.words_to_bytes(synth.pack_windows(synth.ascii_to_codes(b"ACTTGTCTG"), 9), 9)[0]
    assert [format(x, "08b") for x in b] == ["11110100", "11011110", "00000010"]


@pytest.mark.parametrize("k", [9, 27, 36, 63])
def test_revcomp_and_canonical_match_string_semantics(k):
    rng = np.random.default_rng(k)
    codes = rng.integers(0, 4, size=(300, k), dtype=np.uint8)
    codes[0] = 0                      # poly-A
    codes[1, : k // 2] = 0            # low half
    comp = {65: 84, 67: 71, 71: 67, 84: 65}
    words = np.concatenate([synth.pack_windows(c, k) for c in codes])
    rc = synth.revcomp_words(words, k)
    canon = synth.canonical_words(words, k)
    asc = synth.words_to_ascii(words, k)
    for i in range(len(codes)):
        s = bytes(asc[i])
        r = bytes(comp[c] for c in reversed(s))
        assert bytes(synth.words_to_ascii(rc[i:i + 1], k)[0]) == r
        want = r if s >= r else s      # strcmp(fwd, rc) >= 0 -> rc (reference src/bft.c:1291)
        assert bytes(synth.words_to_ascii(canon[i:i + 1], k)[0]) == want
    np.testing.assert_array_equal(synth.bytes_to_words(synth.words_to_bytes(words, k), k), words)


def test_xxh64_known_answers_and_reference(tmp_path):
    src = tmp_path / "x.c"
    src.write_text('#include <stdio.h>\n#include <stdlib.h>\n#include "bft_xxh64.h"\n'
                   'int main(int c, char** v){ unsigned long long seed = strtoull(v[2], 0, 10);'
                   ' printf("%016llx\\n", (unsigned long long)bft_xxh64(v[1], strlen(v[1]), seed)); return 0; }\n')
    exe = str(tmp_path / "x")
    _gcc(["-O2", "-I", CSRC, str(src), "-o", exe])

    def mine(s, seed=0):
        return int(subprocess.run([exe, s, str(seed)], stdout=subprocess.PIPE, check=True).stdout, 16)

    assert mine("") == 0xEF46DB3751D8E999      # published xxHash64 test vectors
    assert mine("abc") == 0x44BC2CF5AD770999
    lib_path = os.path.join(refutil.REF_DIR, "libbft_ref.so")
    if os.path.exists(lib_path):               # the reference's vendored XXH64 (src/xxhash.c), all length classes
        ref = ctypes.CDLL(lib_path).BFT_HASH_XXH64
        ref.restype = ctypes.c_uint64
        ref.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64]
        for s, seed in [("a", 1), ("abc", 1804289383), ("0123456789abcdef", 7), ("x" * 31, 3), ("y" * 32, 5), ("z" * 77, 846930886)]:
            assert mine(s, seed) == ref(s.encode(), len(s), seed)


# ---- serializer + walk + colour decoder on the host --------------------------------------------------------------
@pytest.fixture(scope="module")
def host_tool(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("tool") / "arena_host_query")
    _gcc(["-O2", "-std=c11", "-I", CSRC, os.path.join(ROOT, "tests", "tools", "arena_host_query.c"),
          os.path.join(CSRC, "bft_flatten.c"), os.path.join(CSRC, "bft_io.c"), "-o", exe])
    return exe


def _csv_bytes(names, rows, n_genomes):
    out = [",".join(names).encode() + b"\n"]
    bits = np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little")[:, :n_genomes]
    for r in bits:
        out.append(b",".join(b"1" if x else b"0" for x in r) + b"\n")
    data = b"".join(out)
    return data[:-1] + b"\0"           # the reference overwrites the last byte with NUL (src/file_io.c:873-876)


@pytest.mark.parametrize("name", NAMES)
def test_flattened_arena_walk_reproduces_golden(name, host_tool, tmp_path):
    z = np.load(os.path.join(refutil.GOLDEN, name + ".npz"))
    k, G = int(z["k"]), int(z["n_genomes"])
    q = str(tmp_path / "q.kc")
    synth.write_kmers_comp(q, z["queries"], k)
    out = str(tmp_path / "out.csv")
    p = subprocess.run([host_tool, os.path.join(refutil.GOLDEN, name + ".bft"), "kmers_comp", q, out], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, check=True)
    assert int(p.stdout.split(b"=")[1]) == int(z["present"].sum())
    names = [f"genome_{i:04d}.kc" for i in range(G)]
    assert open(out, "rb").read() == _csv_bytes(names, z["rows"], G)


@pytest.mark.parametrize("name", NAMES)
def test_storage_locations_are_unique_per_stored_kmer(name, tmp_path):
    """The traversal kernels key their vertex table on the storage location a look-up ends in (bft_lookup_loc): every
    stored k-mer must end in its own location below n_loc, absent k-mers in none, and asking for the location must
    not change the answer. Host build of the same walk the kernels compile."""
    import re
    exe = str(tmp_path / "arena_loc_check")
    _gcc(["-O2", "-std=c11", "-I", CSRC, os.path.join(ROOT, "tests", "tools", "arena_loc_check.c"),
          os.path.join(CSRC, "bft_flatten.c"), os.path.join(CSRC, "bft_io.c"), "-o", exe])
    c = cases.make_golden_case(name)
    k = c["k"]
    stored = np.unique(np.concatenate(c["genome_words"]), axis=0)
    z = np.load(os.path.join(refutil.GOLDEN, name + ".npz"))
    absent = z["queries"][z["present"] == 0]
    q = str(tmp_path / "all.kc")
    synth.write_kmers_comp(q, np.concatenate([stored, absent]), k)
    out = subprocess.run([exe, os.path.join(refutil.GOLDEN, name + ".bft"), q], stdout=subprocess.PIPE, check=True).stdout.decode()
    f = {m.group(1): int(m.group(2)) for m in re.finditer(r"(\w+)=(\d+)", out)}
    assert f["queries"] == len(stored) + len(absent)
    assert f["found"] == f["distinct_locs"] == f["stored"] == len(stored)
    assert f["max_loc"] < f["n_loc"] and f["answer_mismatch"] == 0


@pytest.mark.parametrize("name", NAMES)
def test_accelerated_lookup_equals_the_structural_walk_on_the_host(name, tmp_path):
    """The accelerator tables the engine builds on the device (stored-k-mer filter, fused root directory + filter, collapsed
    subtrees) rebuilt sequentially on the host with the same hash functions and slot formats; the look-up code that reads them
    (bft_arena.h, shared with the kernels) must answer like the walk over the structure: every stored k-mer found with its own
    class, every golden query the same, with and without filters, presence-only, and with the collapsed blocks at a load of 1."""
    import re
    exe = str(tmp_path / "arena_accel_check")
    _gcc(["-O2", "-std=c11", "-I", CSRC, os.path.join(ROOT, "tests", "tools", "arena_accel_check.c"),
          os.path.join(CSRC, "bft_flatten.c"), os.path.join(CSRC, "bft_io.c"), "-o", exe])
    z = np.load(os.path.join(refutil.GOLDEN, name + ".npz"))
    k = int(z["k"])
    q = str(tmp_path / "q.kc")
    synth.write_kmers_comp(q, z["queries"], k)
    for sectors, tight in ((0, 0), (1, 0), (3, 1), (4, 1)):
        out = subprocess.run([exe, os.path.join(refutil.GOLDEN, name + ".bft"), q, str(sectors), str(tight)], stdout=subprocess.PIPE, check=True).stdout.decode()
        f = {m.group(1): int(m.group(2)) for m in re.finditer(r"(\w+)=(\d+)", out)}
        assert f["stored_bad"] == 0 and f["mismatch"] == 0, out
        assert f["queries"] == len(z["queries"]) and f["present"] == int(z["present"].sum()), out
        if name == "golden_deep_k63_g12":
            assert f["deep_kmers"] > 0 and f["deep_prefixes"] > 0, out
            if tight:
                assert f["moved_by_probing"] > 0, out     # the linear probing was actually exercised
        assert f["filter_rejects"] > 0, out                # the filters do reject absent queries


def test_accelerated_lookup_on_a_bench_size_bft(tmp_path):
    """The same check on the 4-genome x 5 Mbp BFT of BASELINE config[0] (8.6 M k-mers, paired buckets at load 0.6) when the bench
    data is present in this checkout: all stored k-mers through the accelerated view, plus 200 000 random queries."""
    import re
    import bench_workloads as wl
    bft = wl.bft_path(wl.C1, 27, 5_000_000)
    if not os.path.exists(bft):
        pytest.skip("data/ BFT of config c1 not built in this checkout (tools/build_bench_data.py c1)")
    exe = str(tmp_path / "arena_accel_check")
    _gcc(["-O2", "-std=c11", "-I", CSRC, os.path.join(ROOT, "tests", "tools", "arena_accel_check.c"),
          os.path.join(CSRC, "bft_flatten.c"), os.path.join(CSRC, "bft_io.c"), "-o", exe])
    rng = np.random.default_rng(7)
    qw = rng.integers(0, 1 << 54, size=(200_000, 1), dtype=np.uint64)
    q = str(tmp_path / "q.kc")
    synth.write_kmers_comp(q, qw, 27)
    out = subprocess.run([exe, bft, q, "1", "0"], stdout=subprocess.PIPE, check=True).stdout.decode()
    f = {m.group(1): int(m.group(2)) for m in re.finditer(r"(\w+)=(\d+)", out)}
    assert f["stored"] > 8_000_000 and f["stored_bad"] == 0 and f["mismatch"] == 0, out
    assert f["filter_rejects"] > 150_000, out     # random 27-mers are absent and nearly all rejected by the fused filter


@pytest.mark.parametrize("name", NAMES)
def test_bucket_block_invariants(name, tmp_path):
    """Occupancy and shape of the hashed suffix blocks: every inline prefix's lines add up to its declared count (slots + overflow
    runs), slots fill from the front (the look-up reads "last slot holds a key" as "bucket full"), the overflow area holds exactly
    what the descriptors announce, and for one-word keys the blocks of more than one bucket have an even number of buckets and
    start on a 64-byte line (buckets pair up inside a DRAM fetch) at a load of at most 0.6 + rounding."""
    import re
    exe = str(tmp_path / "arena_stats")
    _gcc(["-O2", "-std=c11", "-I", CSRC, os.path.join(ROOT, "tests", "tools", "arena_stats.c"), os.path.join(CSRC, "bft_flatten.c"), "-o", exe])
    out = subprocess.run([exe, os.path.join(refutil.GOLDEN, name + ".bft")], stdout=subprocess.PIPE, check=True).stdout.decode()
    f = {m.group(1): int(m.group(2)) for m in re.finditer(r"(\w+)=(\d+)", out)}
    assert f["count_mismatch"] == 0 and f["holes"] == 0 and f["odd_blocks"] == 0 and f["misaligned"] == 0, out
    assert f["in_ovf"] == f["n_ovf"], out
    assert f["block_buckets"] <= f["buckets"] <= f["block_buckets"] + f["blocks"], out      # at most one padding bucket per block
    if f["blocks"]:
        assert f["load_permille"] <= 1000 and (f["W"] > 1 or f["load_permille"] <= 700), out


@pytest.mark.parametrize("name", NAMES)
def test_arena_enumeration_order_is_the_reference_order(name, tmp_path):
    """The serializer's enumeration tables (depth-first pref_out, uc_rank, memcmp rank inside a prefix's lines) walked on
    the host exactly as the extraction kernels walk them: the k-mers must come out in the reference's
    iterate_over_kmers order — SHA-256 of the reference's own `-extract_kmers` output (tests/golden/extract_sha256.json)."""
    import hashlib
    import json
    exe = str(tmp_path / "arena_enum_check")
    _gcc(["-O2", "-std=c11", "-I", CSRC, os.path.join(ROOT, "tests", "tools", "arena_enum_check.c"),
          os.path.join(CSRC, "bft_flatten.c"), "-o", exe])
    out = str(tmp_path / "enum.txt")
    msg = subprocess.run([exe, os.path.join(refutil.GOLDEN, name + ".bft"), out], stdout=subprocess.PIPE, check=True).stdout.decode()
    want = json.load(open(os.path.join(refutil.GOLDEN, "extract_sha256.json")))[name]
    assert f"kmers={want['n_kmers']} duplicates=0 missing=0" in msg
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == want["sha256"]


def test_sequence_file_reader_fasta_fastq_and_lines(tmp_path):
    """bft_read_sequence_file: FASTA (multi-line records, CRLF, ';' comments, empty records, no trailing newline), FASTQ
    and the reference's own line-per-sequence layout all yield the same list of sequences; bft_read_kmer_text_file keeps
    the first k characters of every line long enough."""
    exe = str(tmp_path / "seqfile_dump")
    _gcc(["-O2", "-std=c11", "-I", CSRC, os.path.join(ROOT, "tests", "tools", "seqfile_dump.c"), os.path.join(CSRC, "bft_io.c"), "-o", exe])
    seqs = [b"ACGTACGTAC", b"", b"GGGTTTNNNACGT", b"acgu", b"T" * 200]

    def dump(*a):
        out = subprocess.run([exe, *a], stdout=subprocess.PIPE, check=True).stdout.split(b"\n")
        return int(out[0]), out[1:-1]

    flat = tmp_path / "flat.txt"
    flat.write_bytes(b"\n".join(seqs) + b"\n")
    assert dump("seq", str(flat)) == (len(seqs), seqs)
    flat.write_bytes(b"\r\n".join(seqs))                       # CRLF, no newline at the end
    assert dump("seq", str(flat)) == (len(seqs), seqs)
    fa = tmp_path / "x.fa"
    fa.write_bytes(b";comment\n>r0 desc\nACGTA\nCGTAC\n>r1\n>r2\r\nGGGTTT\r\n\r\nNNNACGT\n>r3\nacgu\n>r4\n" + b"T" * 120 + b"\n" + b"T" * 80)
    assert dump("seq", str(fa)) == (len(seqs), seqs)
    fq = tmp_path / "x.fq"
    fq.write_bytes(b"".join(b"@r%d\n" % i + s + b"\n+\n" + b"#" * len(s) + b"\n" for i, s in enumerate(seqs)))
    assert dump("seq", str(fq)) == (len(seqs), seqs)
    empty = tmp_path / "empty.txt"
    empty.write_bytes(b"")
    assert dump("seq", str(empty)) == (0, [])
    km = tmp_path / "k.txt"
    km.write_bytes(b"ACGTACGTA\nACG\nACGTACGTACGT\r\n\nNNNNNNNNN")
    assert dump("kmers", "9", str(km)) == (3, [b"ACGTACGTA", b"ACGTACGTA", b"NNNNNNNNN"])


def test_flattener_rejects_garbage(host_tool, tmp_path):
    bad = tmp_path / "bad.bft"
    bad.write_bytes(b"\x01\x02\x03")
    p = subprocess.run([host_tool, str(bad), "kmers_comp", "/dev/null", "/dev/null"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 and b"bft_flatten" in p.stderr
    g = os.path.join(refutil.GOLDEN, NAMES[0] + ".bft")
    trunc = tmp_path / "trunc.bft"
    trunc.write_bytes(open(g, "rb").read()[:5000])
    p = subprocess.run([host_tool, str(trunc), "kmers_comp", "/dev/null", "/dev/null"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0 and b"truncated" in p.stderr


# ---- the C-ABI library -----------------------------------------------------------------------------------------
def test_abi_library_loads_and_exports_every_declared_symbol():
    from bloomfiltertrie_b200 import engine
    lib = engine.load_library()
    header = open(os.path.join(ROOT, "include", "bft_b200.h")).read()
    import re
    declared = sorted(set(re.findall(r"\b(bft_b200_[a-z0-9_]+)\s*\(", header)))
    assert declared and set(declared) == set(engine.ABI_SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    compat = open(os.path.join(ROOT, "include", "bft_compat.h")).read()
    for s in ["load_BFT", "free_cdbg", "get_kmer", "is_kmer_in_cdbg", "get_annotation", "get_list_id_genomes",
              "get_count_id_genomes", "presence_genome", "query_sequence", "get_neighbors", "get_predecessors",
              "get_successors", "queryBFT_kmerPresences_from_KmerFiles", "queryBFT_kmerBranching_from_KmerFiles",
              "query_sequences_outputCSV", "free_BFT_kmer", "free_BFT_annotation", "create_kmer", "iterate_over_kmers",
              "v_iterate_over_kmers", "extract_kmers_to_disk", "write_kmer_ascii_to_disk", "write_kmer_comp_to_disk",
              "prefix_matching", "intersection_annotations", "union_annotations", "sym_difference_annotations",
              "intersection_list_id_genomes", "set_marking", "unset_marking", "set_flag_kmer", "get_flag_kmer",
              "extract_core_kmers", "extract_dispensable_kmers", "extract_singleton_kmers", "extract_pangenome_kmers_to_disk",
              "extract_simple_paths_to_disk", "extract_simple_core_paths_to_disk", "BFS", "DFS", "BFS_subgraph", "DFS_subgraph",
              "cdbg_traversal", "get_nb_connected_component"]:
        assert s in compat and hasattr(lib, s), s


def test_graph_normal_forms():
    """The order-independent normal forms the traversal tests compare in (tests/graphutil.py)."""
    import graphutil
    k = 4
    circ = b"ACGTTGCA"                                   # a closed loop of 8 k-mers, written from two different starts
    a = circ + circ[:k - 1]
    b = circ[3:] + circ[:3] + (circ[3:] + circ[:3])[:k - 1]
    assert graphutil.canon_cycle(a, k) == graphutil.canon_cycle(b, k) and len(graphutil.canon_cycle(a, k)) == len(a)
    assert graphutil.canon_cycle(b"ACGTAC", k) == b"ACGTAC"       # an open path is left alone
    short = b"ACAC" + b"A"                                # 2-k-mer loop ACAC <-> CACA, shorter than k - 1
    other = b"CACA" + b"C"
    assert graphutil.canon_cycle(short, k) == graphutil.canon_cycle(other, k)
    kset = {b"AAAC", b"AACG", b"ACGT", b"ACGA", b"CGTT"}  # AACG -> ACGT / ACGA would need 2 successors of AACG? no: ACGT, ACGA follow AACG
    assert graphutil.out_degree(b"AACG", kset) == 2 and graphutil.in_degree(b"ACGT", kset) == 1
    # a reference line that ends in a branching k-mer is cut back to the order-independent path
    assert graphutil.trim_branching_ends([b"AAACG"], kset, k) == [b"AAAC"]
    assert graphutil.trim_branching_ends([b"ACGTT"], kset, k) == [b"ACGTT"]
    labels = np.array([0, 0, 5, 0xFFFFFFFF], dtype=np.uint32)
    assert graphutil.partition(labels, [b"a", b"b", b"c", b"d"]) == {frozenset([b"a", b"b"]), frozenset([b"c"])}


def test_engine_fails_loudly_without_a_gpu():
    import torch
    from bloomfiltertrie_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.BFTError, match="no CPU fallback"):
        engine.BFTEngine(os.path.join(refutil.GOLDEN, NAMES[0] + ".bft"))


# ---- sharding ---------------------------------------------------------------------------------------------------
def test_shard_ranges_cover_everything_in_order():
    for n in [0, 1, 7, 8, 1000, 1001]:
        for world in [1, 2, 3, 8]:
            r = [shard.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    offs = np.concatenate([[0], np.cumsum(np.random.default_rng(0).integers(0, 400, size=257))]).astype(np.uint64)
    for world in [1, 2, 4, 8]:
        r = [shard.shard_sequences(offs, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == 257 and all(r[i][1] == r[i + 1][0] for i in range(world - 1))


_WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
import refutil
from bloomfiltertrie_b200 import shard
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[3], RANK=str(rank), WORLD_SIZE=str(world))
dist.init_process_group("gloo", rank=rank, world_size=world)
z = np.load(os.path.join(refutil.GOLDEN, {name!r} + ".npz"))
bft = os.path.join(refutil.GOLDEN, {name!r} + ".bft")
k, G = int(z["k"]), int(z["n_genomes"])
wd = sys.argv[4]

class OracleEngine:          # stands in for BFTEngine on the CPU: same methods, answers from the C restatement
    device = None
    def query_kmers(self, q):
        p, r = refutil.oracle_kmers(bft, q, k, G, wd); return p, r, None
    def query_sequences(self, chars, offs, thr, canonical):
        return refutil.oracle_sequences(bft, refutil.split_seqs(chars, offs), thr, canonical, G, wd), None
    def query_branching(self, q):
        s, p = refutil.oracle_branching(bft, q, k, wd); return s, p, int(((s > 1) | (p > 1)).sum())

sq = shard.ShardedQuery(OracleEngine())
q = z["queries"][:601]
present, rows = sq.query_kmers(q)
srows = sq.query_sequences(z["seq_chars"], z["seq_offs"], 0.8, False)
cnt = sq.query_branching_count(q[:301])
if rank == 0:
    assert np.array_equal(present, z["present"][:601]) and np.array_equal(rows, z["rows"][:601])
    assert np.array_equal(srows, z["seqrows_c0_t0.8"])
    assert cnt == int(((z["succ"][:301] > 1) | (z["pred"][:301] > 1)).sum())
    print("SHARD_OK")
dist.destroy_process_group()
'''


def test_sharded_query_gathers_in_order_gloo_world2(tmp_path):
    refutil.ensure_oracle()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, name="golden_shallow_k27_g4"))
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(2):
        wd = tmp_path / f"r{r}"
        wd.mkdir()
        procs.append(subprocess.Popen([sys.executable, str(script), str(r), "2", port, str(wd)], stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "SHARD_OK" in outs[0]


def test_flattener_survives_mutated_files_under_asan(tmp_path):
    """Corrupted .bft files (byte flips, truncations, header damage) must be rejected or parsed without any memory
    error: the serializer and the arena walk are built with AddressSanitizer + UBSan and run over the mutations."""
    exe = str(tmp_path / "ahq_asan")
    try:
        _gcc(["-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=c11", "-I", CSRC,
              os.path.join(ROOT, "tests", "tools", "arena_host_query.c"), os.path.join(CSRC, "bft_flatten.c"),
              os.path.join(CSRC, "bft_io.c"), "-o", exe])
    except subprocess.CalledProcessError:
        pytest.skip("sanitizer runtime not available")
    rng = np.random.default_rng(11)
    name = "golden_deep_k63_g12"
    z = np.load(os.path.join(refutil.GOLDEN, name + ".npz"))
    q = str(tmp_path / "q.kc")
    synth.write_kmers_comp(q, z["queries"][:300], int(z["k"]))
    data = open(os.path.join(refutil.GOLDEN, name + ".bft"), "rb").read()
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1")
    outcomes = {"ok": 0, "rejected": 0}
    for t in range(45):
        d = bytearray(data)
        if t % 3 == 0:
            for _ in range(int(rng.integers(1, 8))):
                d[int(rng.integers(0, len(d)))] = int(rng.integers(0, 256))
        elif t % 3 == 1:
            d = d[: int(rng.integers(1, len(d)))]
        else:
            d[int(rng.integers(0, 600))] = int(rng.integers(0, 256))
        m = tmp_path / "m.bft"
        m.write_bytes(bytes(d))
        r = subprocess.run([exe, str(m), "kmers_comp", q, str(tmp_path / "o.csv")], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           env=env, timeout=120)
        assert b"AddressSanitizer" not in r.stderr and b"LeakSanitizer" not in r.stderr and b"runtime error" not in r.stderr \
            and r.returncode >= 0, r.stderr.decode(errors="replace")[:1500]
        outcomes["ok" if r.returncode == 0 else "rejected"] += 1
    assert outcomes["rejected"] > 0
