"""Parity of the CUDA path (through the C-ABI) against the reference's own query path (oracle/_ref), bit-exact:
presence bits, colour rows, branching answers and threshold calls on the same .bft and the same inputs."""
import os

import numpy as np
import pytest

from bloomfiltertrie_b200 import synth
import cases
import refutil

pytestmark = pytest.mark.gpu

THRESHOLDS = [0.8, 0.5, 1.0, 0.3, 1e-6]


@pytest.fixture(scope="module")
def engine_mod():
    from bloomfiltertrie_b200 import engine
    return engine


@pytest.fixture(scope="module", params=list(cases.CASES))
def built(request, workdir, engine_mod):
    if not refutil.have_ref():
        pytest.skip("oracle/_ref (compiled reference) not present; golden-fixture tests cover parity")
    c = cases.make_case(request.param)
    path = refutil.build_bft(workdir, c["name"], c["genome_words"], c["k"])
    eng = engine_mod.BFTEngine(path)
    yield c, path, eng
    eng.close()


def test_kmer_presence_and_colours(built):
    c, path, eng = built
    q = c["queries"]
    ref_present, ref_rows = refutil.ref_kmers(path, q, c["k"], c["n_genomes"])
    present, rows, cls = eng.query_kmers(q, want_rows=True, want_classes=True)
    assert ref_present.sum() > 0 and (ref_present == 0).sum() > 0
    np.testing.assert_array_equal(present, ref_present)
    np.testing.assert_array_equal(rows, ref_rows)
    # the reference's record format in, byte rows out: same answers
    rb = (c["n_genomes"] + 7) // 8
    p_rec, r_rec, n_rec = eng.query_records(synth.words_to_bytes(q, c["k"]))
    np.testing.assert_array_equal(p_rec, ref_present)
    np.testing.assert_array_equal(r_rec, np.ascontiguousarray(ref_rows).view(np.uint8).reshape(len(q), -1)[:, :rb])
    assert n_rec == int(ref_present.sum())
    bits, crows, n_c = eng.query_records_compact(synth.words_to_bytes(q, c["k"]))
    np.testing.assert_array_equal(np.unpackbits(bits, bitorder="little")[:len(q)], ref_present)
    np.testing.assert_array_equal(crows, r_rec[ref_present.astype(bool)])
    assert n_c == n_rec
    # class ids are consistent with rows and counts
    table = eng.class_rows()
    counts = eng.class_counts()
    hit = present.astype(bool)
    assert (cls[~hit] == 0xFFFFFFFF).all()
    np.testing.assert_array_equal(table[cls[hit]], rows[hit])
    pop = np.unpackbits(rows[hit].view(np.uint8), axis=1).sum(axis=1)
    np.testing.assert_array_equal(counts[cls[hit]], pop)
    # device-resident entry points with the in-kernel hit counter, at batch sizes that are not multiples of a warp (every row
    # width: k_query_kmers_rows for 1/2/4 words, k_query_kmers_wide — 32 rows written per warp — for the rest)
    import torch
    dev = torch.device("cuda", eng.device)
    dq = torch.from_numpy(np.ascontiguousarray(q).view(np.int64)).to(dev)
    for n in (len(q), len(q) - 1, 33, 31, 1):
        dp = torch.full((n + 64,), 7, dtype=torch.uint8, device=dev)
        dr = torch.full((n + 64, eng.RW), -1, dtype=torch.int32, device=dev)
        dc = torch.zeros(2, dtype=torch.int64, device=dev)
        eng.query_kmers_device_accumulate(dq, n, dp, dr, dc)
        eng.query_kmers_device_accumulate(dq, n, dp, dr, dc)      # accumulates: twice the hits
        eng.sync()
        np.testing.assert_array_equal(dp[:n].cpu().numpy(), ref_present[:n])
        np.testing.assert_array_equal(dr[:n].cpu().numpy().view(np.uint32), ref_rows[:n])
        assert int(dc[0].item()) == 2 * int(ref_present[:n].sum()) and int(dc[1].item()) == 0
        assert bool((dp[n:] == 7).all()) and bool((dr[n:] == -1).all()), "stores past the end of the batch"


def test_kmer_ascii_input(built):
    c, path, eng = built
    q = c["queries"][:2000]
    asc = synth.words_to_ascii(q, c["k"]).copy()
    asc[5, 3] = ord("N")        # invalid k-mer: dropped by the reference, flagged here
    asc[7] = np.char.lower(asc[7].view("S1")).view(np.uint8)
    valid, present, rows = eng.query_kmers_ascii(asc.tobytes())
    p2, r2, _ = eng.query_kmers(q)
    assert valid[5] == 0 and present[5] == 0 and not rows[5].any()
    ok = np.ones(len(q), bool)
    ok[5] = False
    assert valid[ok].all()
    np.testing.assert_array_equal(present[ok], p2[ok])
    np.testing.assert_array_equal(rows[ok], r2[ok])


def test_branching(built):
    c, path, eng = built
    if c["k"] == 9:
        pytest.skip("the reference itself segfaults on -query_branching at k=9 (root level == leaf level)")
    q = c["queries"][:4000]
    ref_succ, ref_pred = refutil.ref_branching(path, q, c["k"])
    succ, pred, count = eng.query_branching(q)
    np.testing.assert_array_equal(succ, ref_succ)
    np.testing.assert_array_equal(pred, ref_pred)
    assert count == int(((ref_succ > 1) | (ref_pred > 1)).sum())
    # get_neighbors order: 0-3 predecessors, 4-7 successors; counts must agree with the per-neighbour classes
    nb = eng.query_neighbors(q[:500])
    np.testing.assert_array_equal((nb[:, :4] != 0xFFFFFFFF).sum(axis=1), ref_pred[:500])
    np.testing.assert_array_equal((nb[:, 4:] != 0xFFFFFFFF).sum(axis=1), ref_succ[:500])


def test_branching_set_semantics_mode(built):
    """exact=False: successors/predecessors by plain membership of the 8 neighbour k-mers (what the reference
    computes everywhere except at the leaf level of deep tries)."""
    c, path, eng = built
    k = c["k"]
    q = c["queries"][:1500]
    allw = np.unique(np.concatenate(c["genome_words"]), axis=0)

    def as_int(r):
        return sum(int(r[w]) << (64 * w) for w in range(len(r)))

    members = set(as_int(r) for r in allw)
    mask = (1 << (2 * k)) - 1
    ints = [as_int(r) for r in q]

    def has(x):
        return x in members

    exp_succ = np.array([sum(has((x >> 2) | (cc << (2 * (k - 1)))) for cc in range(4)) for x in ints], dtype=np.uint8)
    exp_pred = np.array([sum(has(((x << 2) & mask) | cc) for cc in range(4)) for x in ints], dtype=np.uint8)
    eng.set_reference_exact_branching(False)
    try:
        succ, pred, _ = eng.query_branching(q)
    finally:
        eng.set_reference_exact_branching(True)
    np.testing.assert_array_equal(succ, exp_succ)
    np.testing.assert_array_equal(pred, exp_pred)


@pytest.mark.parametrize("canonical", [False, True])
def test_sequences_threshold(built, canonical):
    c, path, eng = built
    seqs = c["seqs"]
    for thr in THRESHOLDS:
        ref_rows = refutil.ref_sequences(path, seqs, thr, canonical, c["n_genomes"])
        rows, status = eng.query_sequence_list(seqs, thr, canonical)
        assert (status != 2).all()
        short = np.array([len(s) < c["k"] for s in seqs])
        np.testing.assert_array_equal(status == 1, short)
        np.testing.assert_array_equal(rows, ref_rows, err_msg=f"thr={thr} canonical={canonical}")


def test_file_drivers_match_reference_cli(built, workdir):
    c, path, eng = built
    d = os.path.join(workdir, c["name"], "cli")
    os.makedirs(d, exist_ok=True)
    k = c["k"]
    qk = os.path.join(d, "queries.kc")
    synth.write_kmers_comp(qk, c["queries"][:3000], k)
    qt = os.path.join(d, "queries_txt.txt")
    synth.write_kmers_text(qt, c["queries"][:1000], k)
    qs = os.path.join(d, "reads.txt")
    with open(qs, "wb") as f:
        f.write(b"\n".join(c["seqs"]) + b"\n")
    for name, p in (("lk", qk), ("lt", qt), ("ls", qs)):
        with open(os.path.join(d, name), "w") as f:
            f.write(p + "\n")
    out = refutil.ref_cli(path, ["-query_kmers", "kmers_comp", os.path.join(d, "lk"),
                                 "-query_kmers", "kmers", os.path.join(d, "lt"),
                                 "-query_sequences", "0.8", "canonical" if c["canonical"] else "non_canonical", os.path.join(d, "ls")]
                          + (["-query_branching", "kmers_comp", os.path.join(d, "lk")] if k > 9 else []), cwd=d)
    n_present = eng.query_kmers_file(qk, True, os.path.join(d, "mine_k.csv"))
    eng.query_kmers_file(qt, False, os.path.join(d, "mine_t.csv"))
    eng.query_sequences_file(qs, os.path.join(d, "mine_s.csv"), 0.8, c["canonical"])
    n_branching = eng.query_branching_file(qk, True)
    assert open(os.path.join(d, "mine_k.csv"), "rb").read() == open(os.path.join(d, "queries.csv"), "rb").read()
    assert open(os.path.join(d, "mine_t.csv"), "rb").read() == open(os.path.join(d, "queries_txt.csv"), "rb").read()
    assert open(os.path.join(d, "mine_s.csv"), "rb").read() == open(os.path.join(d, "reads.csv"), "rb").read()
    counts = [int(x) for x in __import__("re").findall(r"Nb k-mers present = (\d+)", out)]
    assert counts[0] == n_present
    if k > 9:
        assert refutil.parse_count(out, "Nb branching k-mers") == n_branching


def test_fasta_and_fastq_input_equal_the_flattened_file(built, workdir):
    """north_star (2): sequences from FASTA. The reference reads one sequence per line (src/file_io.c:1519-1524), so
    parity = the CSV of a FASTA (multi-line records, CRLF, comments, an empty record) or FASTQ file must be byte-identical
    to what the REFERENCE CLI writes for the line-per-sequence flattening of the same records."""
    c, path, eng = built
    d = os.path.join(workdir, c["name"], "fasta")
    os.makedirs(d, exist_ok=True)
    seqs = [s for s in c["seqs"] if b">" not in s and b"@" not in s][:60]
    seqs.insert(3, b"")                                   # a record without sequence lines
    flat = os.path.join(d, "flat.txt")
    with open(flat, "wb") as f:
        f.write(b"\n".join(seqs) + b"\n")
    with open(os.path.join(d, "ls"), "w") as f:
        f.write(flat + "\n")
    mode = "canonical" if c["canonical"] else "non_canonical"
    refutil.ref_cli(path, ["-query_sequences", "0.8", mode, os.path.join(d, "ls")], cwd=d)
    want = open(os.path.join(d, "flat.csv"), "rb").read()
    rng = np.random.default_rng(3)
    fa = os.path.join(d, "reads.fa")
    with open(fa, "wb") as f:
        f.write(b"; a comment line\n")
        for i, s in enumerate(seqs):
            f.write(b">read_%d some description\r\n" % i if i % 2 else b">read_%d\n" % i)
            pos = 0
            while pos < len(s):                           # ragged line lengths, some CRLF, a blank line now and then
                w = int(rng.integers(1, 71))
                f.write(s[pos:pos + w] + (b"\r\n" if rng.random() < 0.3 else b"\n"))
                if rng.random() < 0.1:
                    f.write(b"\n")
                pos += w
    eng.query_sequences_file(fa, os.path.join(d, "fa.csv"), 0.8, c["canonical"])
    assert open(os.path.join(d, "fa.csv"), "rb").read() == want
    fq = os.path.join(d, "reads.fq")
    with open(fq, "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b"@read_%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
    eng.query_sequences_file(fq, os.path.join(d, "fq.csv"), 0.8, c["canonical"])
    assert open(os.path.join(d, "fq.csv"), "rb").read() == want


def _sort_rows(words):
    order = np.lexsort([words[:, w] for w in range(words.shape[1])])
    return order


def test_enumeration_matches_inserted_sets_and_reference_extract(built, workdir):
    """iterate_over_kmers / -extract_kmers on the device: the set of (k-mer, colour set) pairs must equal what was
    inserted, and the files the engine writes must be BYTE-IDENTICAL to the reference's own -extract_kmers output
    (same k-mers in the same iterate_over_kmers order, src/extract_kmers.c:3-597), binary and text form."""
    c, path, eng = built
    k, G = c["k"], c["n_genomes"]
    kmers, cls, rows = eng.extract_kmers(want_classes=True, want_rows=True)
    # expected: union of the genome k-mer sets, colours by membership
    allw = np.concatenate(c["genome_words"])
    gid = np.concatenate([np.full(len(w), g, dtype=np.int64) for g, w in enumerate(c["genome_words"])])
    uniq, inv = np.unique(allw, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    assert len(kmers) == len(uniq) == int(eng.stats()["n_kmers"])
    want_rows = np.zeros((len(uniq), eng.RW), dtype=np.uint32)
    np.bitwise_or.at(want_rows, (inv, gid >> 5), (np.uint32(1) << (gid & 31).astype(np.uint32)))
    o = _sort_rows(kmers)
    ou = _sort_rows(uniq)
    np.testing.assert_array_equal(kmers[o], uniq[ou])
    np.testing.assert_array_equal(rows[o], want_rows[ou])
    np.testing.assert_array_equal(eng.class_rows()[cls], rows)
    # the reference's own extraction, and the engine's file writer
    d = os.path.join(workdir, c["name"], "extract")
    os.makedirs(d, exist_ok=True)
    refutil.ref_cli(path, ["-extract_kmers", "kmers_comp", os.path.join(d, "ref.kc")], cwd=d)
    eng.extract_kmers_file(os.path.join(d, "mine.kc"), True)
    nb = synth.kmer_nbytes(k)

    def records(p):
        raw = open(p, "rb").read()
        l1 = raw.index(b"\n")
        l2 = raw.index(b"\n", l1 + 1)
        assert int(raw[:l1]) == k and int(raw[l1 + 1:l2]) == len(uniq)
        return np.frombuffer(raw[l2 + 1:], dtype=np.uint8).reshape(-1, nb)

    mine, ref = records(os.path.join(d, "mine.kc")), records(os.path.join(d, "ref.kc"))
    if not np.array_equal(mine, ref):
        bad = np.nonzero((mine != ref).any(axis=1))[0]
        raise AssertionError(f"{c['name']}: -extract_kmers order differs from the reference at {len(bad)} of {len(ref)} records, first {bad[:5]}")
    assert open(os.path.join(d, "mine.kc"), "rb").read() == open(os.path.join(d, "ref.kc"), "rb").read()
    refutil.ref_cli(path, ["-extract_kmers", "kmers", os.path.join(d, "ref.txt")], cwd=d)
    eng.extract_kmers_file(os.path.join(d, "mine.txt"), False)
    assert open(os.path.join(d, "mine.txt"), "rb").read() == open(os.path.join(d, "ref.txt"), "rb").read()
    # the in-memory enumeration is that same order
    np.testing.assert_array_equal(np.ascontiguousarray(kmers).view(np.uint8).reshape(len(kmers), -1)[:, :nb], ref)


def test_duplicate_kmers_are_refused(workdir, engine_mod):
    """For k > 63 the reference's insertion can store one k-mer several times (its -extract_kmers lists more records
    than distinct k-mers), which makes its own query answers depend on the copy a search meets. The serializer
    detects that and refuses the file."""
    if not refutil.have_ref():
        pytest.skip("needs the compiled reference to build the BFT")
    c = cases.case_deep(k=99, n_genomes=4, n_kmers=100_000, pools=(20, 3, 3, 3, 3, 3), seed=211)
    path = refutil.build_bft(workdir, "dup_k99", c["genome_words"], 99)
    d = os.path.join(workdir, "dup_k99")
    refutil.ref_cli(path, ["-extract_kmers", "kmers_comp", os.path.join(d, "ref.kc")], cwd=d)
    raw = open(os.path.join(d, "ref.kc"), "rb").read()
    l1 = raw.index(b"\n")
    l2 = raw.index(b"\n", l1 + 1)
    rec = np.frombuffer(raw[l2 + 1:], dtype=np.uint8).reshape(-1, synth.kmer_nbytes(99))
    if len(np.unique(rec, axis=0)) == len(rec):
        pytest.skip("this reference build produced no duplicates")
    with pytest.raises(engine_mod.BFTError, match="same k-mer twice"):
        engine_mod.BFTEngine(path)
