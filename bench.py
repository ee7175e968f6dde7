#!/usr/bin/env python
"""bench.py — k-mers queried/sec on the 100-genome synthetic pan-genome BFT (BASELINE.json metric, config[2];
k=27 stands in for "k=31": the reference only accepts k divisible by 9).

  python bench.py [--gpus N] [--steps K] [--warmup W]          engine arm (one process per GPU; torchrun for N>1)
  python bench.py --impl reference [...]                       reference arm: the unmodified reference's CPU query
                                                               path (oracle/_ref) on the box's host cores
A step = one pass of the hot path (k-mer membership + colour rows) over one batch of synthetic queries per GPU.
`value` is measured with the batch already resident in HBM; `e2e` goes through the host C-ABI call with pinned host
buffers, host<->device copies inside the timed region. Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "kmers_queried_per_sec"
UNIT = "k-mers/s"
K = 27
DEFAULT_LEN = 5_000_000
FALLBACK_LEN = 200_000  # built on the fly with the reference when no prebuilt BFT travelled with the repo
MIX = (0.5, 0.25, 0.25)


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# stdout carries the one JSON line and nothing else: libraries that write to fd 1 on their own (NCCL prints its version
# banner there) are pointed at stderr for the whole run, and the line goes to the saved descriptor.
_JSON_FD = None


def claim_stdout():
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except Exception:
                continue
            for nme, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            for t, line in self.rows[-3:]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except Exception:
                    pass
        if not sm:
            return None
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def pick_workload(args):
    import bench_workloads as wl
    if getattr(args, "pangenome", "c3") == "c5":  # informational: 1000 colours, wide rows (RW = 32)
        return wl.C5, (args.genome_len or 100_000)
    cfg = wl.C3
    if args.genome_len:
        L = args.genome_len
    else:
        have = wl.available_lengths(cfg, K)
        L = DEFAULT_LEN if DEFAULT_LEN in have else (max(have) if have else FALLBACK_LEN)
    return cfg, L


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_capture():
    """Summary of the committed ncu --set full capture of the dominant kernel (tools/ncu_summary.py writes it), if any."""
    p = os.path.join(ROOT, "profiles", "ncu_k_query_kmers.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            return None
    return None


def kmer_roofline(eng, ws, n, k_ms, W, RW, args, traffic_ok, kernel=None):
    """Roofline record of the k-mer query kernel (HBM-bound, random access).

    achieved = ALGORITHMIC bytes of the ARENA's walk per launch / CUDA-event time: per k-mer the key words in (8*W), the
    presence byte + colour row out (1 + 4*RW) and one 32*W-byte bucket for every k-mer whose walk reaches a bucket
    (measured on the timed batch by k_kmer_walk_stats; k-mers cut short by the root directory or by the L2-resident
    stored-k-mer filter touch no HBM). The root directory entry, the filter sector and the class row come from L2 and
    are listed as l2_bytes_per_kmer. traffic = ncu dram__bytes of the committed --set full capture of this kernel on
    this workload (profiles/ncu_k_query_kmers.json, written by tools/ncu_summary.py); traffic / algorithmic = wasted
    DRAM traffic. A_min / A_ref (the sectors the REFERENCE layout's walk would dereference) are context:
    the arena does not move them, so they are not the roofline."""
    nodes_pk, depth_pk, found_pk = ws["nodes"] / n, ws["search_depth"] / n, ws["found"] / n
    cc_pk = ws["cc_probed"] / n
    bucket_pk = ws["bucket_searches"] / n
    reject_pk = ws["filter_rejects"] / n
    a_min = 8 * W + (1 + 4 * RW) + 32.0 * (6 * nodes_pk + depth_pk + found_pk)
    a_ref = 8 * W + (1 + 4 * RW) + 32.0 * (5 * nodes_pk + 2 * cc_pk + depth_pk + found_pk)
    a_arena = 8 * W + (1 + 4 * RW) + 32.0 * W * bucket_pk
    achieved = a_arena * n / (k_ms / 1e3) / 1e9
    peak, peak_src = peaks()
    cap = ncu_capture() if traffic_ok else None
    traffic = cap.get("dram_bytes_per_kmer") if cap else None
    probe = None
    if not args.no_probe:
        probe = eng.random_gather_probe(4 << 30, 1 << 28)
    st = eng.stats()
    return {"bound": "hbm", "kernel": kernel or ("k_query_kmers_rows" if RW in (1, 2, 4) else "k_query_kmers+k_expand_rows"),
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": (traffic * n if traffic else None), "peak_source": peak_src, "kernel_ms": k_ms,
            "kmers_per_sec_kernel": n / (k_ms / 1e3),
            "algorithmic_bytes_per_kmer": a_arena,
            "algorithmic_bytes_formula": "8*W in + (1 + 4*RW) out + 32*W * P(walk reaches a bucket)",
            "bucket_accesses_per_kmer": bucket_pk, "filter_rejects_per_kmer": reject_pk, "found_frac": found_pk,
            "l2_bytes_per_kmer": 8 + (32 if st.get("filter_bytes") else 0) + 4 * RW * found_pk,
            "filter_mb": st.get("filter_bytes", 0) / 1e6,
            "dram_bytes_per_kmer_ncu": traffic,
            "wasted_traffic_ratio": (traffic / a_arena) if traffic else None,
            "dram_gbps_physical": (traffic * n / (k_ms / 1e3) / 1e9) if traffic else None,
            "dram_frac_physical": (traffic * n / (k_ms / 1e3) / 1e9 / peak) if traffic else None,
            "ncu_capture": ({"file": cap.get("source"), "kernel": cap.get("kernel"), "filter_mb": cap.get("filter_mb")} if cap else None),
            "random_gather_probe_loads_per_s": probe,
            "random_access_frac": (bucket_pk * n / (k_ms / 1e3) / probe) if probe else None,
            "context_reference_layout": {"a_min_bytes_per_kmer": a_min, "a_ref_bytes_per_kmer": a_ref, "nodes_per_kmer": nodes_pk,
                                         "search_depth_per_kmer": depth_pk,
                                         "cc_probed_per_node": cc_pk / max(nodes_pk, 1e-9),
                                         "mean_suffix_block_lines": ws["block_lines"] / max(1, n),
                                         "note": "sectors the REFERENCE layout's walk dereferences (SURVEY.md 8d); the arena replaces them by one "
                                                 "L2-resident directory load + one bucket, so they are context, not the roofline"},
            "note": "frac = algorithmic HBM bytes of the arena walk / measured copy bandwidth; the kernel is bound by the RATE of random "
                    "64-byte HBM accesses (random_access_frac: bucket accesses/s over the measured rate of independent random loads), "
                    "not by bytes"}


def write_query_file(path, q_np, k):
    from bloomfiltertrie_b200 import synth
    import numpy as np
    synth.write_kmers_comp(path, q_np.view(np.uint64), k)


def run_reference_harness(bft, qfile, threads, passes):
    import bench_workloads as wl
    out = qfile + ".out"
    p = subprocess.run([wl.REF_HARNESS, "kmers", bft, qfile, out, str(threads), str(passes)], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if os.path.exists(out):
        os.remove(out)
    if p.returncode != 0:
        raise RuntimeError("ref_harness failed: " + p.stdout[-1000:])
    return [float(x) for x in re.findall(r"REF_PASS \d+ seconds=([0-9.]+)", p.stdout)]


def reference_arm(args):
    """The reference's own CPU implementation of the path (isKmerPresent + get_annotation + get_list_id_genomes under
    OpenMP, one copy_BFT_Root per thread) on all host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import bench_workloads as wl
    cfg, L = pick_workload(args)
    genomes = wl.pangenome(cfg, L)
    bft = wl.ensure_bft(cfg, K, L, genomes)
    cores = os.cpu_count() or 1
    n = args.ref_sample
    cat, starts, lens = wl.genomes_to_torch(genomes, torch.device("cpu"))
    q = wl.gen_kmer_queries(cat, starts, lens, K, n, seed=777, mix=MIX)[0].numpy()
    qfile = os.path.join("/tmp", f"bft_bench_ref_{os.getpid()}.kc")
    write_query_file(qfile, q, K)
    secs = run_reference_harness(bft, qfile, cores, args.warmup + args.steps)
    os.remove(qfile)
    timed = secs[args.warmup:]
    ms = 1e3 * sum(timed) / len(timed)
    value = n / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"{cfg['name']}_k{K}: -query_kmers on a {cfg['n_genomes']}-genome synthetic pan-genome BFT",
                       "k": K, "n_genomes": cfg["n_genomes"], "genome_len": L, "query_mix_present_mismatch_random": MIX},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"{n} k-mers per step (same generator and mix as the GPU batch), all {cores} host threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def engine_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from bloomfiltertrie_b200 import engine as E
    import bench_workloads as wl

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL writes its banner ("NCCL version ...", when NCCL_DEBUG is set on the box) to stdout; stdout carries the
        # one JSON line only
        claim_stdout()
        dist.init_process_group("nccl", device_id=dev)

    cfg, L = pick_workload(args)
    t0 = time.time()
    genomes = wl.pangenome(cfg, L)
    if rank == 0:
        bft = wl.ensure_bft(cfg, K, L, genomes)
    if world > 1:
        dist.barrier()
    bft = wl.bft_path(cfg, K, L)
    eng = E.BFTEngine(bft, device=local)
    st = eng.stats()
    if rank == 0:
        log(f"arena: {st['n_kmers']} k-mers, {st['n_nodes']} nodes, {st['n_ccs']} CCs, {st['n_classes']} colour classes, "
            f"{st['arena_bytes'] / 1e6:.0f} MB (+{st['class_row_bytes'] / 1e6:.0f} MB class rows); flatten {st['flatten_seconds']:.1f}s "
            f"upload {st['upload_seconds']:.1f}s decode {st['decode_seconds']:.3f}s; setup {time.time() - t0:.1f}s")
    n = args.queries_per_gpu
    cat, starts, lens = wl.genomes_to_torch(genomes, dev)
    q, q_kind = wl.gen_kmer_queries(cat, starts, lens, K, n, seed=1000 + rank, mix=MIX)
    del cat
    torch.cuda.empty_cache()
    RW, W = eng.RW, eng.W
    d_present = torch.empty(n, dtype=torch.uint8, device=dev)
    d_rows = torch.empty((n, RW), dtype=torch.int32, device=dev)
    es = torch.cuda.ExternalStream(eng.stream, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The hit count of every step ("Nb k-mers present" of the reference driver) is the one number the sharded path
    # reduces. It is accumulated by the query kernels themselves: one uint64 slot per step in rank 0's HBM, mapped into
    # every rank through CUDA IPC, and each CTA adds its share with one system-scope atomic (over NVLink from the other
    # ranks) — compute + reduction in one kernel, no NCCL call and no extra launch in the steady state.
    counted = RW in (1, 2, 4)
    n_slots = args.warmup + args.steps + args.warmup + args.steps + 8
    ctr_local = eng.device_alloc(8 * n_slots) if rank == 0 else 0   # zero-filled by the call
    ctr_base = ctr_local
    if world > 1:
        box = [eng.peer_export(ctr_local) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            ctr_base = eng.peer_import(box[0])
    slot_i = [0]

    def step():
        if counted:
            eng.query_kmers_device_accumulate(q, n, d_present, d_rows, ctr_base + 8 * slot_i[0])
            slot_i[0] += 1
        else:
            eng.query_kmers_device(q, n, d_present, d_rows, None)

    # ---- timed region: K steps, inputs resident in HBM (batch of n*8*W bytes >> 126 MB L2, so no L2 flush needed)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.time()
    ev0.record(es)
    for _ in range(args.steps):
        step()
    ev1.record(es)
    barrier()
    tw1 = time.time()
    ms_total = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = n * world / (ms_step / 1e3)
    n_present = int(d_present.sum().item())
    if counted:  # every step's slot must hold the sum of all ranks' hits (checked outside the timed region)
        tot = torch.tensor([n_present], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        if rank == 0:
            slots = eng.copy_from_device(ctr_local, np.zeros(n_slots, dtype=np.uint64))
            used = slots[:slot_i[0]]
            assert (used == np.uint64(int(tot.item()))).all(), f"in-kernel hit counters {used.tolist()} != {int(tot.item())}"
    # size-independent property at full size: every window sampled from an inserted genome must be found, and a
    # found k-mer must carry at least one colour
    assert bool(d_present[q_kind == 0].all()), "a k-mer window of an inserted genome was reported absent"
    assert bool((d_rows[d_present.bool()] != 0).any(dim=1).all()), "a present k-mer came back without colours"
    assert not bool((d_rows[~d_present.bool()] != 0).any()), "an absent k-mer came back with colours"

    d_hits = torch.zeros(1, dtype=torch.int64, device=dev)

    def step_kernel():
        if counted:
            eng.query_kmers_device_accumulate(q, n, d_present, d_rows, d_hits)
        else:
            eng.query_kmers_device(q, n, d_present, d_rows, None)

    # ---- dominant kernel alone: the fused walk + colour-row kernel (k_query_kmers_rows for RW in {1,2,4}; for wider
    # rows k_query_kmers followed by k_expand_rows), CUDA events on the stream it is launched on
    for _ in range(args.warmup):
        step_kernel()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(es)
    for _ in range(args.steps):
        step_kernel()
    b.record(es)
    b.synchronize()
    k_ms = a.elapsed_time(b) / args.steps
    # clocks sampled from the start of the timed region to the end of the per-kernel timing loop (GPU busy throughout)
    clocks = sampler.stop(tw0, time.time()) if rank == 0 else None
    ws = eng.kmer_walk_stats_device(q, n)
    nodes_pk, depth_pk, found_pk = ws["nodes"] / n, ws["search_depth"] / n, ws["found"] / n
    roofline = kmer_roofline(eng, ws, n, k_ms, W, RW, args, traffic_ok=(cfg["name"] == wl.C3["name"] and L == 5_000_000 and K == 27))

    # ---- e2e: host C-ABI call with pinned host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        hq = E.PinnedBuffer((n, W), np.uint64)
        hp = E.PinnedBuffer((n,), np.uint8)
        hr = E.PinnedBuffer((n, RW), np.uint32)
        hq.array[:] = q.cpu().numpy().view(np.uint64)
        for _ in range(max(1, args.warmup - 1)):
            eng.query_kmers(hq.array, out_present=hp.array, out_rows=hr.array)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.query_kmers(hq.array, out_present=hp.array, out_rows=hr.array)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item()) / args.steps * 1e3
        assert int(hp.array.sum()) == n_present, "e2e and device-resident paths disagree"
        word_api = {"value": n * world / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": n * W * 8 * world,
                    "d2h_bytes_per_step": n * (1 + 4 * RW) * world, "ms_per_step": e2e_ms,
                    "api": "bft_b200_query_kmers (host pointers, pinned): 8*W-byte words in, presence byte + 4*RW-byte colour row out"}
        # the same call asking for colour-class ids instead of rows (4 B instead of 4*RW B back per k-mer; the class ->
        # row table is downloaded once per context): information for link-bound deployments, not the headline
        hc = E.PinnedBuffer((n,), np.uint32)
        eng.query_kmers(hq.array, want_rows=False, out_present=hp.array, out_classes=hc.array)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.query_kmers(hq.array, want_rows=False, out_present=hp.array, out_classes=hc.array)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        c_ms = float(tt.item()) / args.steps * 1e3
        class_id_mode = {"value": n * world / (c_ms / 1e3), "unit": UNIT, "ms_per_step": c_ms,
                                "d2h_bytes_per_step": n * 5 * world}
        # the headline e2e: the same batch in the reference's own record format (ceil(2k/8)-byte kmers_comp records in,
        # ceil(G/8)-byte colour rows + the present count out) — fewer bytes over the link that bounds this call
        nb, rb = (2 * K + 7) // 8, (cfg["n_genomes"] + 7) // 8
        hrec = E.PinnedBuffer((n, nb), np.uint8)
        hrec.array[:] = hq.array.view(np.uint8).reshape(n, 8 * W)[:, :nb]
        ns = min(n, 1 << 20)
        want_sample = hr.array[:ns].view(np.uint8).reshape(ns, 4 * RW)[:, :rb].copy()
        hq.free(); hp.free(); hr.free(); hc.free()          # keep the pinned footprint per rank small
        hrow = E.PinnedBuffer((n, rb), np.uint8)
        _, _, cnt = eng.query_records(hrec.array, want_present=False, out_rows=hrow.array)
        assert cnt == n_present, "record-format and device-resident paths disagree"
        assert np.array_equal(hrow.array[:ns], want_sample), "record-format rows differ"
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.query_records(hrec.array, want_present=False, out_rows=hrow.array)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        r_ms = float(tt.item()) / args.steps * 1e3
        fixed = {"value": n * world / (r_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": n * nb * world,
                 "d2h_bytes_per_step": (n * rb + 8) * world, "ms_per_step": r_ms,
                 "api": "bft_b200_query_records (host pointers, pinned): the reference's ceil(2k/8)-byte k-mer records in, "
                        "ceil(G/8)-byte colour row per k-mer (fixed stride) + the number of k-mers present out"}
        # the same answers without the zeros: one presence bit per k-mer + the rows of the present k-mers only, in query
        # order (the row of k-mer i is found with one running index, the way a CSV writer walks the batch)
        hbits = E.PinnedBuffer(((n + 7) // 8,), np.uint8)
        _, crows, cnt = eng.query_records_compact(hrec.array, out_bits=hbits.array, out_rows=hrow.array)
        assert cnt == n_present, "compact and device-resident paths disagree"
        bits_s = np.unpackbits(hbits.array[: ns // 8], bitorder="little").astype(bool)
        assert np.array_equal(crows[: int(bits_s.sum())], want_sample[: len(bits_s)][bits_s]), "compact rows differ"
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.query_records_compact(hrec.array, out_bits=hbits.array, out_rows=hrow.array)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        k_ms2 = float(tt.item()) / args.steps * 1e3
        e2e = {"value": n * world / (k_ms2 / 1e3), "unit": UNIT, "h2d_bytes_per_step": n * nb * world,
               "d2h_bytes_per_step": ((n + 7) // 8 + n_present * rb + 4 * ((n + (1 << 22) - 1) >> 22)) * world, "ms_per_step": k_ms2,
               "api": "bft_b200_query_records_compact (host pointers, pinned): the reference's ceil(2k/8)-byte k-mer records in; "
                      "one presence bit per k-mer + the ceil(G/8)-byte colour rows of the present k-mers (query order) + their count out — "
                      "every answer of the fixed-stride call, without the all-zero rows of absent k-mers",
               "present_frac": n_present / n,
               "fixed_stride_records": fixed, "word_api": word_api, "class_id_mode": class_id_mode}
        hbits.free()
        hrec.free(); hrow.free()

    # ---- CPU baseline beside it: the unmodified reference on a bounded sample (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.access(wl.REF_HARNESS, os.X_OK):
        cores = os.cpu_count() or 1
        ns = min(args.ref_sample, n)
        qfile = os.path.join("/tmp", f"bft_bench_cpu_{os.getpid()}.kc")
        write_query_file(qfile, q[:ns].cpu().numpy(), K)
        secs = run_reference_harness(bft, qfile, cores, 2)
        os.remove(qfile)
        cpu = {"value": ns / min(secs), "unit": UNIT, "cores": cores, "kind": "reference",
               "sample": f"first {ns} k-mers of the GPU batch; isKmerPresent+get_annotation+get_list_id_genomes, OpenMP over "
                         f"{cores} threads with one copy_BFT_Root each; best of 2 passes"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
                "data": "synthetic",
                "config": {"workload": f"{cfg['name']}_k{K}: -query_kmers (presence + colour rows) on a {cfg['n_genomes']}-genome "
                                       f"synthetic pan-genome BFT built by the reference",
                           "k": K, "k_note": "reference accepts only k % 9 == 0; 27 stands in for 31", "n_genomes": cfg["n_genomes"],
                           "genome_len": L, "kmers_in_bft": st["n_kmers"], "colour_classes": st["n_classes"],
                           "arena_mb": round(st["arena_bytes"] / 1e6, 1), "queries_per_gpu": n,
                           "query_mix_present_mismatch_random": MIX, "present_frac": n_present / n,
                           "l2": "inputs larger than L2 (no flush needed)", "sharding": f"arena replicated, queries sharded x{world}"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": launches * world, "roofline": roofline, "cpu_baseline": cpu}
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def side_workload(args):
    """Informational single-GPU lines for the other two query kinds (not the headline metric): BASELINE config[1]
    (-query_sequences 0.8 canonical, 150 bp reads vs the 16-genome canonical BFT) and config[3] (-query_branching at
    k=63 on the 100-genome BFT). Same JSON keys; `metric` stays k-mers (windows / k-mers) per second."""
    import numpy as np
    import torch
    from bloomfiltertrie_b200 import engine as E
    import bench_workloads as wl
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    seq = args.workload == "sequences"
    cfg, k, L = (wl.C2, 27, args.genome_len or 5_000_000) if seq else (wl.C3, 63, args.genome_len or 1_000_000)
    genomes = wl.pangenome(cfg, L)
    bft = wl.ensure_bft(cfg, k, L, genomes)
    eng = E.BFTEngine(bft, device=0)
    st = eng.stats()
    cat, starts, lens = wl.genomes_to_torch(genomes, dev)
    es = torch.cuda.ExternalStream(eng.stream, device=dev)
    cores = os.cpu_count() or 1
    if seq:
        n, rl = args.reads, 150
        chars, offs = wl.gen_reads(cat, starts, lens, n, rl, seed=4242)
        units = n * (rl - k + 1)
        d_rows = torch.empty((n, eng.RW), dtype=torch.int32, device=dev)
        d_stat = torch.empty(n, dtype=torch.uint8, device=dev)
        run = lambda: eng.query_sequences_device(chars, offs, n, 0.8, True, d_rows, d_stat)
        h_chars = E.PinnedBuffer((n * rl,), np.uint8); h_chars.array[:] = chars.cpu().numpy()
        h_offs = E.PinnedBuffer((n + 1,), np.uint64); h_offs.array[:] = offs.cpu().numpy().view(np.uint64)
        h_rows = E.PinnedBuffer((n, eng.RW), np.uint32); h_stat = E.PinnedBuffer((n,), np.uint8)
        run_e2e = lambda: eng.query_sequences(h_chars.array, h_offs.array, 0.8, True, out_rows=h_rows.array, out_status=h_stat.array)
        h2d, d2h = n * rl + 8 * (n + 1), n * (4 * eng.RW + 1)
    else:
        n = args.queries_per_gpu
        q, _ = wl.gen_kmer_queries(cat, starts, lens, k, n, seed=99, mix=MIX)
        units = n
        d_succ = torch.empty(n, dtype=torch.uint8, device=dev)
        d_pred = torch.empty(n, dtype=torch.uint8, device=dev)
        d_cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        run = lambda: eng.query_branching_device(q, n, d_succ, d_pred, d_cnt)
        hq = E.PinnedBuffer((n, eng.W), np.uint64); hq.array[:] = q.cpu().numpy().view(np.uint64)
        hs = E.PinnedBuffer((n,), np.uint8); hp = E.PinnedBuffer((n,), np.uint8)
        run_e2e = lambda: eng.query_branching(hq.array, out_succ=hs.array, out_pred=hp.array)
        h2d, d2h = n * 8 * eng.W, 2 * n + 8
    for _ in range(args.warmup):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count()
    a.record(es)
    for _ in range(args.steps):
        run()
    b.record(es)
    b.synchronize()
    ms = a.elapsed_time(b) / args.steps
    launches = eng.launch_count() - l0
    run_e2e()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_e2e()
    e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
    cpu = None
    if not args.no_cpu_baseline and os.access(wl.REF_HARNESS, os.X_OK):
        if seq:
            ns = min(n, 200_000)
            p = f"/tmp/bft_bench_reads_{os.getpid()}.txt"
            arr = chars[: ns * 150].cpu().numpy().reshape(ns, 150)
            with open(p, "wb") as f:
                f.write(np.concatenate([arr, np.full((ns, 1), 10, np.uint8)], axis=1).tobytes())
            out = subprocess.run([wl.REF_HARNESS, "sequences", bft, p, "0.8", "canonical", p + ".out", str(cores), "2"],
                                 stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
            secs = [float(x) for x in re.findall(r"REF_PASS \d+ seconds=([0-9.]+)", out)]
            ref_rows = np.fromfile(p + ".out", dtype=np.uint32).reshape(ns, eng.RW)
            assert np.array_equal(ref_rows, d_rows[:ns].cpu().numpy().view(np.uint32)), "GPU and reference disagree on the sample"
            os.remove(p); os.remove(p + ".out")
            cpu = {"value": ns * (150 - k + 1) / min(secs), "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"first {ns} reads; query_sequence under OpenMP, one copy_BFT_Root per thread; rows identical to the GPU's"}
        else:
            ns = min(n, 1 << 22)
            p = f"/tmp/bft_bench_br_{os.getpid()}.kc"
            write_query_file(p, q[:ns].cpu().numpy(), k)
            out = subprocess.run([wl.REF_HARNESS, "branching", bft, p, p + ".out", str(cores), "2"], stdout=subprocess.PIPE,
                                 stderr=subprocess.STDOUT, text=True).stdout
            secs = [float(x) for x in re.findall(r"REF_PASS \d+ seconds=([0-9.]+)", out)]
            raw = np.fromfile(p + ".out", dtype=np.uint8)
            assert np.array_equal(raw[:ns], d_succ[:ns].cpu().numpy()) and np.array_equal(raw[ns:2 * ns], d_pred[:ns].cpu().numpy()), \
                "GPU and reference disagree on the sample"
            os.remove(p); os.remove(p + ".out")
            cpu = {"value": ns / min(secs), "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"first {ns} k-mers; isBranchingRight + isBranchingLeft under OpenMP; counts identical to the GPU's"}
    line = {"metric": METRIC, "value": units / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": ("c2: -query_sequences 0.8 canonical, 150 bp reads vs 16-genome canonical BFT (value counts k-mer windows)"
                                    if seq else "c4: -query_branching at k=63 on the 100-genome BFT (value counts query k-mers; 8 neighbour lookups each)"),
                       "k": k, "n_genomes": cfg["n_genomes"], "genome_len": L, "kmers_in_bft": st["n_kmers"], "items_per_step": n,
                       "reads_per_sec": (n / (ms / 1e3)) if seq else None},
            "e2e": {"value": units / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
            "gpu_launches": launches, "cpu_baseline": cpu}
    emit(line)
    eng.close()


def graph_workload(args):
    """Informational single-GPU line for the traversal snippets: build the device graph of the
    100-genome BFT, count its connected components, extract its simple paths; beside it the reference's own
    get_nb_connected_component(BFS) / extract_simple_core_paths_to_disk on one host core (they cannot thread)."""
    import numpy as np
    import torch
    from bloomfiltertrie_b200 import engine as E
    import bench_workloads as wl
    torch.cuda.set_device(0)
    cfg, k, L = wl.C3, 27, args.genome_len or 1_000_000
    genomes = wl.pangenome(cfg, L)
    bft = wl.ensure_bft(cfg, k, L, genomes)
    eng = E.BFTEngine(bft, device=0)
    st = eng.stats()
    n = int(st["n_kmers"])
    t = {"build": [], "components": [], "paths": []}
    n_comp = n_paths = longest = path_bytes = 0
    l0 = 0
    for it in range(args.warmup + args.steps):
        if it == args.warmup:
            t = {kk: [] for kk in t}
            l0 = eng.launch_count()
        eng.graph_release()
        eng.sync()
        t0 = time.perf_counter()
        eng.graph_prepare()
        t1 = time.perf_counter()
        n_comp = eng.connected_components()
        t2 = time.perf_counter()
        _, n_paths, longest, path_bytes = eng.simple_paths_raw(0.0, copy=False)   # the C call: kernels + copy to a host buffer
        t3 = time.perf_counter()
        t["build"].append(t1 - t0); t["components"].append(t2 - t1); t["paths"].append(t3 - t2)
    launches = eng.launch_count() - l0
    ms = {kk: 1e3 * sum(v) / len(v) for kk, v in t.items()}
    step_ms = ms["build"] + ms["components"]
    cpu = None
    ref_graph = os.path.join(os.path.dirname(wl.REF_HARNESS), "ref_graph")
    if not args.no_cpu_baseline and os.access(ref_graph, os.X_OK):
        t0 = time.perf_counter()
        out = subprocess.run([ref_graph, "components", bft, "bfs"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        dt = time.perf_counter() - t0 - st["flatten_seconds"] * 0  # includes the reference's own load of the file
        m = re.search(r"REF_COMPONENTS (\d+)", out)
        assert m and int(m.group(1)) == n_comp, f"GPU and reference disagree: {n_comp} vs {out[-200:]}"
        cpu = {"value": n / dt, "unit": "k-mers/s", "cores": 1, "kind": "reference",
               "sample": f"get_nb_connected_component(BFS) over the whole BFT incl. load_BFT: {dt:.1f} s; same count as the GPU ({n_comp})"}
    line = {"metric": "k-mers traversed/sec (graph build + connected components)", "value": n / (step_ms / 1e3), "unit": "k-mers/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "graph: get_nb_connected_component + extract_simple_paths_to_disk on the 100-genome BFT",
                       "k": k, "n_genomes": cfg["n_genomes"], "genome_len": L, "kmers_in_bft": n, "graph_build_ms": ms["build"],
                       "components_ms": ms["components"], "simple_paths_ms": ms["paths"], "n_components": n_comp, "n_paths": n_paths,
                       "longest_path": longest, "path_bytes": path_bytes,
                       "timing": "host clock around the blocking C-ABI calls (each ends with a stream synchronize)"},
            "e2e": {"value": n / ((step_ms + ms["paths"]) / 1e3), "unit": "k-mers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": path_bytes + 8,
                    "ms_per_step": step_ms + ms["paths"], "note": "build + components + simple paths copied to the host"},
            "gpu_launches": launches, "cpu_baseline": cpu}
    emit(line)
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--genome-len", type=int, default=0, help="override the genome length of the 100-genome pan-genome")
    ap.add_argument("--queries-per-gpu", type=int, default=125_000_000)
    ap.add_argument("--ref-sample", type=int, default=1 << 24, help="k-mers per step of the CPU reference legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--workload", default="kmers", choices=["kmers", "sequences", "branching", "graph"],
                    help="kmers = the headline metric (default); the other two print informational single-GPU lines")
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--pangenome", default="c3", choices=["c3", "c5"],
                    help="c3 = 100 genomes (headline); c5 = 1000 colours, 100 kbp genomes (informational)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.workload == "graph":
        graph_workload(args)
    elif args.workload != "kmers":
        if args.queries_per_gpu == 125_000_000:
            args.queries_per_gpu = 20_000_000
        side_workload(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        engine_arm(args)


if __name__ == "__main__":
    main()
