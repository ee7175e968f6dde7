#!/usr/bin/env python
"""bench.py — k-mers queried/sec on the 100-genome synthetic pan-genome BFT (BASELINE.json metric, config[2];
k=27 stands in for "k=31": the reference only accepts k divisible by 9), with the other four BASELINE
configs measured in the same run as sub-records (`configs`: c1, c2, c4, c5) plus one forced-deep trie (`deep`).

  python bench.py [--gpus N] [--steps K] [--warmup W]          engine arm (one process per GPU; torchrun for N>1)
  python bench.py --impl reference [...]                       reference arm: the unmodified reference's CPU query
                                                               path (oracle/_ref) on the box's host cores
  python bench.py --config c4 --sub ""                         one config as the main line (profiling)
A step = one pass of the hot path over one batch of synthetic queries per GPU. `value` is measured with the batch
already resident in HBM; `e2e` goes through the host C-ABI call with pinned host buffers, host<->device copies inside the
timed region. Prints ONE JSON line on rank 0.
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
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions (B200_PROFILING.md clocks line)."""
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

    def window(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in list(self.rows):
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
            for t, line in list(self.rows)[-3:]:
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

    def stop(self):
        if self.proc:
            self.proc.terminate()
            self.proc = None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_capture(tag):
    """Summary of the committed ncu --set full capture of a config's dominant kernel (tools/ncu_summary.py writes
    profiles/ncu_summary.json from the .ncu-rep files), if any."""
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(tag)
        except Exception:
            return None
    return None


def write_query_file(path, q_np, k):
    from bloomfiltertrie_b200 import synth
    import numpy as np
    synth.write_kmers_comp(path, q_np.view(np.uint64), k)


def run_ref(argv):
    import bench_workloads as wl
    p = subprocess.run([wl.REF_HARNESS, *argv], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("ref_harness failed: " + p.stdout[-1000:])
    return [float(x) for x in re.findall(r"REF_PASS \d+ seconds=([0-9.]+)", p.stdout)]


# ---------------------------------------------------------------------------------------------------------------------
# the BASELINE configs. queries = per GPU per step. fallback_L: a smaller BFT of the same pan-genome that
# is used — and flagged `degraded` in the record — when the full-size file did not travel with the repo snapshot.
def specs():
    import bench_workloads as wl
    return {
        "c3": dict(kind="kmers", cfg=wl.C3, k=27, L=5_000_000, queries=125_000_000,
                   workload="c3: -query_kmers (presence + colour rows) on the 100-genome synthetic pan-genome BFT built by the reference; "
                            "1 B k-mers per step over 8 GPUs (125 M per GPU)"),
        "c1": dict(kind="kmers", cfg=wl.C1, k=27, L=5_000_000, queries=10_000_000,
                   workload="c1: -query_kmers, 10 M random+present k-mers on the BFT of 4 synthetic 5 Mbp genomes"),
        "c2": dict(kind="sequences", cfg=wl.C2, k=27, L=5_000_000, reads=1_000_000,
                   workload="c2: -query_sequences threshold 0.8 canonical, 1 M synthetic 150 bp reads vs the 16-genome pan-genome BFT "
                            "(value counts k-mer windows)"),
        "c4": dict(kind="branching", cfg=wl.C3, k=63, L=5_000_000, fallback_L=1_000_000, queries=100_000_000,
                   workload="c4: -query_branching at k=63 on the 100-genome BFT (value counts query k-mers; 8 neighbour look-ups each)"),
        "c5": dict(kind="kmers", cfg=wl.C5, k=27, L=500_000, fallback_L=100_000, queries=100_000_000,
                   workload="c5: colour-set retrieval for 100 M k-mers on the 1000-colour pan-genome BFT (compressed annotations)"),
        # not a BASELINE config: the per-level path of the walk, 3-4 Nodes per look-up
        "deep": dict(kind="kmers", cfg=wl.DEEP, k=63, L=0, queries=30_000_000,
                     workload="deep: -query_kmers on a forced-deep trie (6 M random 63-mers with pooled 9-nt blocks, 16 921 Nodes, 4 levels); "
                              "queries 1/3 members, 1/3 one-mismatch, 1/3 random"),
    }


class Run:
    """torch.distributed plumbing shared by every record: barrier, max over ranks, device."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU reference)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            claim_stdout()  # NCCL writes its banner to stdout when NCCL_DEBUG is set on the box
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.args = args
        self.sampler = ClockSampler(self.local)
        if self.rank == 0:
            self.sampler.start()
        self.cores = os.cpu_count() or 1

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, x: float) -> float:
        if not self.dist:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x: int) -> int:
        if not self.dist:
            return int(x)
        t = self.torch.tensor([x], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t)
        return int(t.item())

    def open(self, spec, headline):
        """The BFT of a config: rank 0 makes sure the .bft is there (decompressing the shipped .xz); everyone opens it.
        Returns (engine, stats, path, genome length used, degraded note or None)."""
        import bench_workloads as wl
        from bloomfiltertrie_b200 import engine as E
        cfg, k = spec["cfg"], spec["k"]

        def have(L):
            p = wl.bft_path(cfg, k, L)
            return os.path.exists(p) or os.path.exists(p + ".xz")

        L, note, err = spec["L"], None, None
        if self.rank == 0:
            if not have(L):
                fb = spec.get("fallback_L")
                if self.args.allow_build:
                    pass
                elif fb and have(fb):
                    note = (f"genome length {fb} instead of {L}: {os.path.basename(wl.bft_path(cfg, k, L))}[.xz] did not travel with the "
                            f"repo snapshot (see DESIGN.md, data shipping); same pan-genome generator, smaller BFT")
                    L = fb
                else:
                    err = (f"{os.path.relpath(wl.bft_path(cfg, k, L), ROOT)}[.xz] is absent: the BFT of this config is built by the reference "
                           f"(tools/build_bench_data.py, tens of minutes) and shipped under data/; refusing to substitute another one "
                           f"(--allow-build builds it here)")
            if not err:
                try:
                    wl.ensure_bft(cfg, k, L, None)
                except Exception as e:  # noqa: BLE001
                    err = repr(e)
        if self.dist:
            box = [L, note, err]
            self.dist.broadcast_object_list(box, src=0)
            L, note, err = box
        if err:
            raise FileNotFoundError(err)
        path = wl.bft_path(cfg, k, L)
        t0 = time.time()
        eng = E.BFTEngine(path, device=self.local)
        st = eng.stats()
        if self.rank == 0:
            log(f"{os.path.basename(path)}: {st['n_kmers']} k-mers, {st['n_nodes']} nodes, {st['n_ccs']} CCs, {st['n_classes']} colour classes, "
                f"arena {st['arena_bytes'] / 1e6:.0f} MB + class rows {st['class_row_bytes'] / 1e6:.0f} MB + filter {st['filter_bytes'] / 1e6:.0f} MB "
                f"+ fused root/filter {st['rootkf_bytes'] / 1e6:.0f} MB + collapsed subtrees {st['deep_bytes'] / 1e6:.0f} MB; "
                f"flatten {st['flatten_seconds']:.1f}s upload {st['upload_seconds']:.1f}s decode+filter {st['decode_seconds']:.3f}s; open {time.time() - t0:.1f}s")
        return eng, st, path, L, note

    def timed(self, eng, step, steps, warmup, flush_l2=False):
        """W untimed steps, then K steps between barrier + synchronize on both sides, CUDA events on the engine's stream,
        max over ranks. flush_l2: the batch is not larger than L2, so a 256 MB buffer is overwritten between steps and
        every step gets its own event pair (the flushes are outside the pairs)."""
        torch = self.torch
        es = torch.cuda.ExternalStream(eng.stream, device=self.dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev) if flush_l2 else None

        def do_flush():
            with torch.cuda.stream(es):
                flush.fill_(1)

        for _ in range(warmup):
            step()
        self.barrier()
        l0 = eng.launch_count()
        tw0 = time.time()
        if flush_l2:
            pairs = []
            for _ in range(steps):
                do_flush()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(es)
                step()
                b.record(es)
                pairs.append((a, b))
            self.barrier()
            ms_total = sum(a.elapsed_time(b) for a, b in pairs)
        else:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(es)
            for _ in range(steps):
                step()
            ev1.record(es)
            self.barrier()
            ms_total = ev0.elapsed_time(ev1)
        tw1 = time.time()
        launches = eng.launch_count() - l0
        ms_step = self.allmax(ms_total) / steps
        del flush
        return ms_step, launches, (tw0, tw1)

    def clocks(self, win):
        if self.rank != 0:
            return None
        time.sleep(0.05)
        return self.sampler.window(*win)

    def host_timed(self, fn, steps, warmup=1):
        """e2e: the blocking host C-ABI call (copies inside), wall clock, max over ranks."""
        for _ in range(warmup):
            fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        self.torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return self.allmax(dt) / steps * 1e3


def base_roofline(kernel, a_bytes, formula, units, k_ms, cap, extra):
    peak, peak_src = peaks()
    achieved = a_bytes * units / (k_ms / 1e3) / 1e9
    traffic_per_unit = (cap["dram_bytes_per_launch"] / cap["units_per_launch"]) if cap else None
    r = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": (traffic_per_unit * units if traffic_per_unit else None), "peak_source": peak_src, "kernel_ms": k_ms,
         "units_per_launch": units, "algorithmic_bytes_per_unit": a_bytes, "algorithmic_bytes_formula": formula,
         "dram_bytes_per_unit_ncu": traffic_per_unit,
         "wasted_traffic_ratio": (traffic_per_unit / a_bytes) if traffic_per_unit else None,
         "dram_frac_physical": (traffic_per_unit * units / (k_ms / 1e3) / 1e9 / peak) if traffic_per_unit else None,
         "ncu_capture": ({k: cap.get(k) for k in ("source", "kernel", "units_per_launch", "duration_ms_under_ncu", "lts_hit_rate_pct",
                                                  "dram_pct_of_peak", "sm_pct_of_peak")} if cap else None)}
    r.update(extra)
    return r


# ---------------------------------------------------------------------------------------------------------------------
def kmers_record(run: Run, tag: str, spec: dict, headline: bool):
    """-query_kmers: presence + colour rows of a batch of packed k-mers (k_query_kmers_rows, or k_query_kmers_wide for more
    than 128 genomes)."""
    import numpy as np
    torch = run.torch
    from bloomfiltertrie_b200 import engine as E
    import bench_workloads as wl
    args = run.args
    cfg, k = spec["cfg"], spec["k"]
    n = args.queries_per_gpu if (headline and args.queries_per_gpu) else spec["queries"]
    steps, warmup = (args.steps, args.warmup) if headline else (args.sub_steps, 3)
    eng, st, bft, L, degraded = run.open(spec, headline)
    if cfg.get("pools"):
        words, _ = wl.kmer_sets(cfg, k)
        q, q_kind = wl.gen_set_queries(words, k, n, 1000 + run.rank, run.dev)
        del words
    else:
        genomes = wl.pangenome(cfg, L)
        cat, starts, lens = wl.genomes_to_torch(genomes, run.dev)
        q, q_kind = wl.gen_kmer_queries(cat, starts, lens, k, n, seed=1000 + run.rank, mix=MIX)
        del cat
    torch.cuda.empty_cache()
    RW, W = eng.RW, eng.W
    dev = run.dev
    d_present = torch.empty(n, dtype=torch.uint8, device=dev)
    d_rows = torch.empty((n, RW), dtype=torch.int32, device=dev)
    counted = True   # every row width accumulates the hit count in the kernel (k_query_kmers_rows / k_query_kmers_wide)
    # The hit count of every step ("Nb k-mers present" of the reference driver, src/file_io.c:813) is the one number the
    # sharded path reduces. The query kernels accumulate it themselves: one uint64 slot per step in rank 0's HBM, mapped
    # into every rank through CUDA IPC; each CTA adds its share with one system-scope atomic (over NVLink from the other
    # ranks) — compute + reduction in one kernel, no NCCL call and no extra launch in the steady state.
    n_slots = 2 * (warmup + steps) + 8
    ctr_local = eng.device_alloc(8 * n_slots) if run.rank == 0 else 0   # zero-filled by the call
    ctr_base = ctr_local
    reduction = "none (one GPU)"
    if run.dist and counted:
        box = [eng.peer_export(ctr_local) if run.rank == 0 else None]
        run.dist.broadcast_object_list(box, src=0)
        failed = 0
        if run.rank != 0:
            try:
                ctr_base = eng.peer_import(box[0])
            except Exception as e:  # noqa: BLE001 — a box without CUDA IPC between its GPUs
                log(f"rank {run.rank}: peer mapping of the hit counter failed ({e})")
                failed = 1
        if run.allsum(failed):
            # Fallback, named in the record: every rank counts into its own slots and the slots are summed ONCE after the
            # timed region (still no collective per step).
            if run.rank != 0:
                if not failed:
                    eng.peer_close(ctr_base)
                ctr_local = ctr_base = eng.device_alloc(8 * n_slots)
            reduction = "per-rank counters, one NCCL all_reduce after the timed region (CUDA IPC unavailable on this box)"
        else:
            reduction = "inside the kernel: system-scope atomics into rank 0's peer-mapped counter (CUDA IPC over NVLink)"
    peer_mapped = reduction.startswith("inside")
    slot_i = [0]

    def step():
        if counted:
            eng.query_kmers_device_accumulate(q, n, d_present, d_rows, ctr_base + 8 * slot_i[0])
            slot_i[0] += 1
        else:
            eng.query_kmers_device(q, n, d_present, d_rows, None)

    flush = n * 8 * W < (200 << 20)
    ms_step, launches, win = run.timed(eng, step, steps, warmup, flush_l2=flush)
    value = n * run.world / (ms_step / 1e3)
    n_present = int(d_present.sum().item())
    tot = run.allsum(n_present)
    if counted:  # every step's slot must hold the sum of all ranks' hits
        if run.dist and not peer_mapped:
            slots_t = torch.from_numpy(eng.copy_from_device(ctr_local, np.zeros(n_slots, dtype=np.uint64)).view(np.int64)).to(dev)
            run.dist.all_reduce(slots_t)
            slots = slots_t.cpu().numpy().view(np.uint64)
        elif run.rank == 0:
            slots = eng.copy_from_device(ctr_local, np.zeros(n_slots, dtype=np.uint64))
        if run.rank == 0:
            used = slots[:slot_i[0]]
            assert (used == np.uint64(tot)).all(), f"in-kernel hit counters {used.tolist()} != {tot}"
    # size-independent properties at full size: every window sampled from an inserted genome is found, a found k-mer
    # carries at least one colour, an absent one none
    assert bool(d_present[q_kind == 0].all()), "a k-mer window of an inserted genome was reported absent"
    pb = d_present.bool()
    if RW <= 4:
        assert bool((d_rows[pb] != 0).any(dim=1).all()), "a present k-mer came back without colours"
        assert not bool((d_rows[~pb] != 0).any()), "an absent k-mer came back with colours"
    else:  # wide rows: check a slice (the boolean gathers above would need several GB)
        m = min(n, 4_000_000)
        assert bool((d_rows[:m][pb[:m]] != 0).any(dim=1).all()) and not bool((d_rows[:m][~pb[:m]] != 0).any())
    del pb

    # ---- dominant kernel alone (CUDA events on the stream it is launched on)
    d_hits = torch.zeros(1, dtype=torch.int64, device=dev)

    def step_kernel():
        if counted:
            eng.query_kmers_device_accumulate(q, n, d_present, d_rows, d_hits)
        else:
            eng.query_kmers_device(q, n, d_present, d_rows, None)

    k_ms, _, win2 = run.timed(eng, step_kernel, steps, warmup, flush_l2=flush)
    clocks = run.clocks((win[0], win2[1]))
    ws = eng.kmer_walk_stats_device(q, n)
    found_pk, bucket_pk, reject_pk = ws["found"] / n, ws["bucket_searches"] / n, ws["filter_rejects"] / n
    nodes_pk, depth_pk, cc_pk = ws["nodes"] / n, ws["search_depth"] / n, ws["cc_probed"] / n
    a_arena = 8 * W + (1 + 4 * RW) + 32.0 * W * bucket_pk
    cap = ncu_capture(tag if not degraded else tag + "_fb")
    probe = eng.random_gather_probe(4 << 30, 1 << 28) if (headline and not args.no_probe) else getattr(run, "probe", None)
    run.probe = probe   # the sub-records quote their random-access fraction against the same in-process measurement
    roofline = base_roofline(
        "k_query_kmers_rows" if RW <= 4 else "k_query_kmers_wide", a_arena,
        "8*W in + (1 + 4*RW) out + 32*W * P(walk reaches a bucket)", n, k_ms, cap,
        {"kmers_per_sec_kernel": n / (k_ms / 1e3), "bucket_accesses_per_kmer": bucket_pk, "filter_rejects_per_kmer": reject_pk,
         "found_frac": found_pk, "filter_mb": st["filter_bytes"] / 1e6, "rootkf_mb": st["rootkf_bytes"] / 1e6,
         "l2_random_requests_per_kmer": (1 if st["rootkf_bytes"] else (2 if st["filter_bytes"] else 1)) + bucket_pk + found_pk,
         "l2_bytes_per_kmer": (32 if st["rootkf_bytes"] else 8 + (32 if st["filter_bytes"] else 0)) + 4 * RW * found_pk,
         "random_gather_probe_loads_per_s": probe,
         "random_access_frac": (bucket_pk * n / (k_ms / 1e3) / probe) if probe else None,
         "context_reference_layout": {
             "a_min_bytes_per_kmer": 8 * W + (1 + 4 * RW) + 32.0 * (6 * nodes_pk + depth_pk + found_pk),
             "a_ref_bytes_per_kmer": 8 * W + (1 + 4 * RW) + 32.0 * (5 * nodes_pk + 2 * cc_pk + depth_pk + found_pk),
             "nodes_per_kmer": nodes_pk, "search_depth_per_kmer": depth_pk, "cc_probed_per_node": cc_pk / max(nodes_pk, 1e-9),
             "note": "sectors the REFERENCE layout's walk dereferences (SURVEY.md 8d); the arena replaces them by one L2-resident "
                     "directory load + one bucket, so they are context, not the roofline"},
         "note": "achieved = algorithmic HBM bytes of the arena's walk / kernel time; the root directory entry + filter bits (one fused "
                 "sector when rootkf_mb > 0) and the class rows are served by L2 (l2_bytes_per_kmer). The kernel is bound by the RATE of "
                 "random accesses — 64-byte HBM accesses (random_access_frac) and random L2 sectors (l2_random_requests_per_kmer; "
                 "about 300 G/s chip-wide, tools/mix_probe.cu) — on top of the streamed batch, not by bytes"})

    # ---- e2e: host C-ABI calls with pinned host buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        nb, rb = (2 * k + 7) // 8, (cfg["n_genomes"] + 7) // 8
        ne = n if RW <= 4 else min(n, 10_000_000)   # wide rows: 125 B per present k-mer back — bound the pinned footprint
        qh = q[:ne].cpu().numpy().view(np.uint64)
        hrec = E.PinnedBuffer((ne, nb), np.uint8)
        hrec.array[:] = qh.view(np.uint8).reshape(ne, 8 * W)[:, :nb]
        hrow = E.PinnedBuffer((ne, rb), np.uint8)
        hbits = E.PinnedBuffer(((ne + 7) // 8,), np.uint8)
        _, crows, cnt = eng.query_records_compact(hrec.array, out_bits=hbits.array, out_rows=hrow.array)
        ne_present = int(d_present[:ne].sum().item())
        assert cnt == ne_present, "compact and device-resident paths disagree"
        ns = min(ne, 1 << 20)
        want = d_rows[:ns].cpu().numpy().view(np.uint8).reshape(ns, 4 * RW)[:, :rb]
        bits_s = np.unpackbits(hbits.array[: ns // 8], bitorder="little").astype(bool)
        assert np.array_equal(crows[: int(bits_s.sum())], want[: len(bits_s)][bits_s]), "compact rows differ"
        es = max(2, steps // 2)
        c_ms = run.host_timed(lambda: eng.query_records_compact(hrec.array, out_bits=hbits.array, out_rows=hrow.array), es)
        e2e = {"value": ne * run.world / (c_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": ne * nb * run.world,
               "d2h_bytes_per_step": ((ne + 7) // 8 + ne_present * rb + 4 * ((ne + (1 << 22) - 1) >> 22)) * run.world, "ms_per_step": c_ms,
               "kmers_per_step_per_gpu": ne, "steps": es,
               "api": "bft_b200_query_records_compact (host pointers, pinned): the reference's ceil(2k/8)-byte k-mer records in; one presence bit "
                      "per k-mer + the ceil(G/8)-byte colour rows of the present k-mers (query order) + their count out",
               "present_frac": ne_present / ne}
        if headline:
            r_ms = run.host_timed(lambda: eng.query_records(hrec.array, want_present=False, out_rows=hrow.array), es)
            e2e["fixed_stride_records"] = {"value": ne * run.world / (r_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": ne * nb * run.world,
                                           "d2h_bytes_per_step": (ne * rb + 8) * run.world, "ms_per_step": r_ms,
                                           "api": "bft_b200_query_records: same records in, one colour row per k-mer (fixed stride) out"}
        hbits.free()
        hrow.free()
        hrec.free()
        if headline or RW > 4:  # class ids instead of rows: what a link-bound deployment (or 1000 colours) asks for
            hq = E.PinnedBuffer((ne, W), np.uint64)
            hq.array[:] = qh
            hp = E.PinnedBuffer((ne,), np.uint8)
            hc = E.PinnedBuffer((ne,), np.uint32)
            i_ms = run.host_timed(lambda: eng.query_kmers(hq.array, want_rows=False, out_present=hp.array, out_classes=hc.array), es)
            assert int(hp.array.sum()) == ne_present
            e2e["class_id_mode"] = {"value": ne * run.world / (i_ms / 1e3), "unit": UNIT, "ms_per_step": i_ms, "h2d_bytes_per_step": ne * 8 * W * run.world,
                                    "d2h_bytes_per_step": ne * 5 * run.world,
                                    "api": "bft_b200_query_kmers with class ids instead of rows (the class -> row table is fetched once per context)"}
            hq.free()
            hp.free()
            hc.free()

    # ---- CPU baseline beside it: the unmodified reference on a bounded sample (rank 0)
    cpu = None
    if run.rank == 0 and not args.no_cpu_baseline and os.access(wl.REF_HARNESS, os.X_OK):
        ns = min(args.ref_sample if headline else (1 << 22), n)
        qfile = os.path.join("/tmp", f"bft_bench_cpu_{os.getpid()}.kc")
        write_query_file(qfile, q[:ns].cpu().numpy(), k)
        secs = run_ref(["kmers", bft, qfile, qfile + ".out", str(run.cores), "2"])
        raw = np.fromfile(qfile + ".out", dtype=np.uint8)
        assert np.array_equal(raw[:ns], d_present[:ns].cpu().numpy()), "GPU and reference disagree on the sample (presence)"
        assert np.array_equal(raw[ns:].view(np.uint32).reshape(ns, RW), d_rows[:ns].cpu().numpy().view(np.uint32)), \
            "GPU and reference disagree on the sample (rows)"
        os.remove(qfile)
        os.remove(qfile + ".out")
        cpu = {"value": ns / min(secs), "unit": UNIT, "cores": run.cores, "kind": "reference",
               "sample": f"first {ns} k-mers of rank 0's batch; isKmerPresent+get_annotation+get_list_id_genomes, OpenMP over {run.cores} threads "
                         f"with one copy_BFT_Root each; best of 2 passes; answers identical to the GPU's"}
    if run.dist:
        run.dist.barrier()
    rec = {"value": value, "unit": UNIT, "ms_per_step": ms_step, "steps": steps, "warmup": warmup, "n_gpus": run.world,
           "config": {"workload": spec["workload"], "k": k, "k_note": "reference accepts only k % 9 == 0; 27 stands in for 31" if k == 27 else None,
                      "n_genomes": cfg["n_genomes"], "genome_len": L, "degraded": degraded, "kmers_in_bft": st["n_kmers"], "nodes": st["n_nodes"],
                      "colour_classes": st["n_classes"], "arena_mb": round(st["arena_bytes"] / 1e6, 1), "filter_mb": round(st["filter_bytes"] / 1e6, 1),
                      "rootkf_mb": round(st["rootkf_bytes"] / 1e6, 1), "collapsed_subtrees_mb": round(st["deep_bytes"] / 1e6, 1),
                      "arena_bytes_per_kmer": round(st["arena_bytes"] / max(1, st["n_kmers"]), 1), "nodes_per_lookup": nodes_pk,
                      "queries_per_gpu": n, "query_mix_present_mismatch_random": (1 / 3, 1 / 3, 1 / 3) if cfg.get("pools") else MIX,
                      "present_frac": n_present / n,
                      "l2": ("256 MB written between steps (batch smaller than L2); one event pair per step" if flush
                             else "inputs larger than L2 (no flush needed)"),
                      "sharding": f"arena replicated, queries sharded x{run.world}", "hit_count_reduction": reduction},
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches * run.world, "roofline": roofline, "cpu_baseline": cpu}
    if run.dist and counted and run.rank != 0 and peer_mapped:
        eng.peer_close(ctr_base)
    if run.dist:
        run.dist.barrier()
    if run.rank == 0 or (run.dist and not peer_mapped):
        eng.device_free(ctr_local)
    eng.close()
    del q, d_present, d_rows
    torch.cuda.empty_cache()
    return rec


def sequences_record(run: Run, tag: str, spec: dict, headline: bool):
    """-query_sequences 0.8 canonical: 150 bp reads, per-read per-genome hit counts and threshold (k_query_sequences)."""
    import numpy as np
    torch = run.torch
    from bloomfiltertrie_b200 import engine as E, synth
    import bench_workloads as wl
    args = run.args
    cfg, k = spec["cfg"], spec["k"]
    steps, warmup = (args.steps, args.warmup) if headline else (args.sub_steps, 3)
    eng, st, bft, L, degraded = run.open(spec, headline)
    genomes = wl.pangenome(cfg, L)
    cat, starts, lens = wl.genomes_to_torch(genomes, run.dev)
    n, rl = spec["reads"], 150
    chars, offs = wl.gen_reads(cat, starts, lens, n, rl, seed=4242 + run.rank)
    del cat
    units = n * (rl - k + 1)
    d_rows = torch.empty((n, eng.RW), dtype=torch.int32, device=run.dev)
    d_stat = torch.empty(n, dtype=torch.uint8, device=run.dev)
    step = lambda: eng.query_sequences_device(chars, offs, n, 0.8, True, d_rows, d_stat)  # noqa: E731
    ms_step, launches, win = run.timed(eng, step, steps, warmup)
    clocks = run.clocks(win)
    assert int((d_stat != 0).sum().item()) == 0
    frac_hit = float((d_rows != 0).any(dim=1).float().mean().item())
    # one substitution error removes up to k of the 124 windows (> 20 %), so only the error-free reads (~47 % at 0.5 % per base)
    # are sure to reach the 0.8 threshold
    assert frac_hit > 0.3, "error-free reads sampled from the genomes must reach the 0.8 threshold for some genome"
    # walk statistics of the canonical windows of a sample of reads (host-side packing, device-side walk)
    codes = synth._CODE[chars[: 2000 * rl].cpu().numpy()].reshape(2000, rl)
    wins = np.concatenate([synth.canonical_words(synth.pack_windows(c, k), k) for c in codes])
    ws = eng.kmer_walk_stats_device(torch.from_numpy(wins.view(np.int64)).to(run.dev), len(wins))
    bucket_pw = ws["bucket_searches"] / len(wins)
    W, RW, G = eng.W, eng.RW, cfg["n_genomes"]
    a_win = 1.0 + 32.0 * W * bucket_pw + (4 * RW + 1 + 8) / (rl - k + 1)
    roofline = base_roofline("k_query_sequences", a_win,
                             "per k-mer window: 1 char streamed + 32*W * P(walk reaches a bucket) + (row + status + offset) / windows per read",
                             units, ms_step, ncu_capture(tag if not degraded else tag + "_fb"),
                             {"bucket_accesses_per_window": bucket_pw, "filter_rejects_per_window": ws["filter_rejects"] / len(wins),
                              "found_frac_windows": ws["found"] / len(wins),
                              "random_gather_probe_loads_per_s": getattr(run, "probe", None),
                              "random_access_frac_upper": (bucket_pw * units / (ms_step / 1e3) / run.probe) if getattr(run, "probe", None) else None,
                              "note": "one launch per step, so kernel time = step time. Bound by random HBM accesses (one bucket per window "
                                      "looked up; random_access_frac_upper assumes every window is looked up — the early exit of reads that "
                                      "cannot reach the threshold skips some) and, second, by instruction issue (see the ncu capture)"})
    e2e = None
    if not args.no_e2e:
        h_chars = E.PinnedBuffer((n * rl,), np.uint8)
        h_chars.array[:] = chars.cpu().numpy()
        h_offs = E.PinnedBuffer((n + 1,), np.uint64)
        h_offs.array[:] = offs.cpu().numpy().view(np.uint64)
        h_rows = E.PinnedBuffer((n, RW), np.uint32)
        h_stat = E.PinnedBuffer((n,), np.uint8)
        e_ms = run.host_timed(lambda: eng.query_sequences(h_chars.array, h_offs.array, 0.8, True, out_rows=h_rows.array, out_status=h_stat.array),
                              max(2, steps // 2))
        assert np.array_equal(h_rows.array, d_rows.cpu().numpy().view(np.uint32))
        e2e = {"value": units * run.world / (e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": (n * rl + 8 * (n + 1)) * run.world,
               "d2h_bytes_per_step": n * (4 * RW + 1) * run.world, "ms_per_step": e_ms, "reads_per_sec": n * run.world / (e_ms / 1e3),
               "api": "bft_b200_query_sequences (host pointers, pinned): characters + offsets in, one colour row + status per read out"}
        for b in (h_chars, h_offs, h_rows, h_stat):
            b.free()
    cpu = None
    if run.rank == 0 and not args.no_cpu_baseline and os.access(wl.REF_HARNESS, os.X_OK):
        nr = min(n, 200_000)
        p = f"/tmp/bft_bench_reads_{os.getpid()}.txt"
        arr = chars[: nr * rl].cpu().numpy().reshape(nr, rl)
        with open(p, "wb") as f:
            f.write(np.concatenate([arr, np.full((nr, 1), 10, np.uint8)], axis=1).tobytes())
        secs = run_ref(["sequences", bft, p, "0.8", "canonical", p + ".out", str(run.cores), "2"])
        ref_rows = np.fromfile(p + ".out", dtype=np.uint32).reshape(nr, RW)
        assert np.array_equal(ref_rows, d_rows[:nr].cpu().numpy().view(np.uint32)), "GPU and reference disagree on the sample"
        os.remove(p)
        os.remove(p + ".out")
        cpu = {"value": nr * (rl - k + 1) / min(secs), "unit": UNIT, "cores": run.cores, "kind": "reference",
               "sample": f"first {nr} reads of rank 0's batch; query_sequence under OpenMP, one copy_BFT_Root per thread; rows identical to the GPU's"}
    if run.dist:
        run.dist.barrier()
    rec = {"value": units * run.world / (ms_step / 1e3), "unit": UNIT, "ms_per_step": ms_step, "steps": steps, "warmup": warmup, "n_gpus": run.world,
           "config": {"workload": spec["workload"], "k": k, "n_genomes": G, "genome_len": L, "degraded": degraded, "kmers_in_bft": st["n_kmers"],
                      "reads_per_gpu": n, "read_len": rl, "windows_per_read": rl - k + 1, "reads_per_sec": n * run.world / (ms_step / 1e3),
                      "reads_reaching_threshold": frac_hit, "l2": "inputs larger than L2 (150 MB of characters per step)",
                      "sharding": f"arena replicated, reads sharded x{run.world}"},
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches * run.world, "roofline": roofline, "cpu_baseline": cpu}
    eng.close()
    del chars, offs, d_rows, d_stat
    torch.cuda.empty_cache()
    return rec


def neighbours_np(q, k):
    """The 8 neighbours of packed k-mers (uint64 [n, W]) as the branching kernel forms them: [n*8, W]."""
    import numpy as np
    n, W = q.shape
    out = np.zeros((n, 8, W), dtype=np.uint64)
    mask = [np.uint64((1 << min(64, max(0, 2 * k - 64 * w))) - 1) for w in range(W)]
    for c in range(4):
        s = np.zeros((n, W), dtype=np.uint64)   # successor: (x >> 2) | c << 2(k-1)
        p = np.zeros((n, W), dtype=np.uint64)   # predecessor: (x << 2 | c) masked
        for w in range(W):
            s[:, w] = q[:, w] >> np.uint64(2)
            p[:, w] = (q[:, w] << np.uint64(2)) & mask[w]
            if w + 1 < W:
                s[:, w] |= (q[:, w + 1] & np.uint64(3)) << np.uint64(62)
            if w > 0:
                p[:, w] |= q[:, w - 1] >> np.uint64(62)
                p[:, w] &= mask[w]
        top = 2 * (k - 1)
        s[:, top // 64] |= np.uint64(c) << np.uint64(top % 64)
        p[:, 0] |= np.uint64(c)
        out[:, c] = s
        out[:, 4 + c] = p
    return out.reshape(n * 8, W)


def branching_record(run: Run, tag: str, spec: dict, headline: bool):
    """-query_branching: successors / predecessors present in the graph for every k-mer (k_query_branching, 8 look-ups each)."""
    import numpy as np
    torch = run.torch
    from bloomfiltertrie_b200 import engine as E
    import bench_workloads as wl
    args = run.args
    cfg, k = spec["cfg"], spec["k"]
    steps, warmup = (args.steps, args.warmup) if headline else (args.sub_steps, 3)
    eng, st, bft, L, degraded = run.open(spec, headline)
    genomes = wl.pangenome(cfg, L)
    cat, starts, lens = wl.genomes_to_torch(genomes, run.dev)
    n = args.queries_per_gpu if (headline and args.queries_per_gpu) else spec["queries"]
    q, _ = wl.gen_kmer_queries(cat, starts, lens, k, n, seed=99 + run.rank, mix=MIX)
    del cat
    W = eng.W
    d_succ = torch.empty(n, dtype=torch.uint8, device=run.dev)
    d_pred = torch.empty(n, dtype=torch.uint8, device=run.dev)
    d_cnt = torch.zeros(1, dtype=torch.int64, device=run.dev)
    step = lambda: eng.query_branching_device(q, n, d_succ, d_pred, d_cnt)  # noqa: E731
    ms_step, launches, win = run.timed(eng, step, steps, warmup)
    clocks = run.clocks(win)
    n_br = int(d_cnt.item())
    assert n_br == int(((d_succ > 1) | (d_pred > 1)).sum().item()), "in-kernel branching counter disagrees with the per-k-mer counts"
    ns = 500_000
    nb8 = neighbours_np(q[:ns].cpu().numpy().view(np.uint64), k)
    ws = eng.kmer_walk_stats_device(torch.from_numpy(nb8.view(np.int64)).to(run.dev), len(nb8))
    bucket_pl = ws["bucket_searches"] / len(nb8)
    # the walk statistics should count exactly the neighbours the kernel found (same k-mers, same look-up)
    stats_consistent = ws["found"] == int(d_succ[:ns].sum().item()) + int(d_pred[:ns].sum().item())
    a_q = 8.0 * W + 2 + 8 * 32.0 * W * bucket_pl
    roofline = base_roofline("k_query_branching", a_q, "per query k-mer: 8*W in + 2 out + 8 look-ups * 32*W * P(walk reaches a bucket)", n, ms_step,
                             ncu_capture(tag if not degraded else tag + "_fb"),
                             {"lookups_per_sec": 8 * n / (ms_step / 1e3), "bucket_accesses_per_lookup": bucket_pl,
                              "filter_rejects_per_lookup": ws["filter_rejects"] / len(nb8), "neighbours_present_frac": ws["found"] / len(nb8),
                              "filter_mb": st["filter_bytes"] / 1e6, "walk_stats_match_kernel_counts": bool(stats_consistent),
                              "random_gather_probe_loads_per_s": getattr(run, "probe", None),
                              "random_access_frac": (8 * n * bucket_pl / (ms_step / 1e3) / run.probe) if getattr(run, "probe", None) else None,
                              "note": "one launch per step, so kernel time = step time; most of the 8 neighbours of a k-mer are absent and are "
                                      "answered by the L2-resident stored-k-mer filter"})
    e2e = None
    if not args.no_e2e:
        hq = E.PinnedBuffer((n, W), np.uint64)
        hq.array[:] = q.cpu().numpy().view(np.uint64)
        hs = E.PinnedBuffer((n,), np.uint8)
        hp = E.PinnedBuffer((n,), np.uint8)
        e_ms = run.host_timed(lambda: eng.query_branching(hq.array, out_succ=hs.array, out_pred=hp.array), max(2, steps // 2))
        assert np.array_equal(hs.array, d_succ.cpu().numpy()) and np.array_equal(hp.array, d_pred.cpu().numpy())
        e2e = {"value": n * run.world / (e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": n * 8 * W * run.world, "d2h_bytes_per_step": (2 * n + 8) * run.world,
               "ms_per_step": e_ms, "api": "bft_b200_query_branching (host pointers, pinned): packed k-mers in, successor / predecessor counts + branching total out"}
        hq.free()
        hs.free()
        hp.free()
    cpu = None
    if run.rank == 0 and not args.no_cpu_baseline and os.access(wl.REF_HARNESS, os.X_OK):
        nsr = min(n, 1 << 22)
        p = f"/tmp/bft_bench_br_{os.getpid()}.kc"
        write_query_file(p, q[:nsr].cpu().numpy(), k)
        secs = run_ref(["branching", bft, p, p + ".out", str(run.cores), "2"])
        raw = np.fromfile(p + ".out", dtype=np.uint8)
        assert np.array_equal(raw[:nsr], d_succ[:nsr].cpu().numpy()) and np.array_equal(raw[nsr:2 * nsr], d_pred[:nsr].cpu().numpy()), \
            "GPU and reference disagree on the sample"
        os.remove(p)
        os.remove(p + ".out")
        cpu = {"value": nsr / min(secs), "unit": UNIT, "cores": run.cores, "kind": "reference",
               "sample": f"first {nsr} k-mers of rank 0's batch; isBranchingRight + isBranchingLeft under OpenMP; counts identical to the GPU's"}
    if run.dist:
        run.dist.barrier()
    rec = {"value": n * run.world / (ms_step / 1e3), "unit": UNIT, "ms_per_step": ms_step, "steps": steps, "warmup": warmup, "n_gpus": run.world,
           "config": {"workload": spec["workload"], "k": k, "n_genomes": cfg["n_genomes"], "genome_len": L, "degraded": degraded,
                      "kmers_in_bft": st["n_kmers"], "nodes": st["n_nodes"], "arena_mb": round(st["arena_bytes"] / 1e6, 1),
                      "filter_mb": round(st["filter_bytes"] / 1e6, 1), "collapsed_subtrees_mb": round(st["deep_bytes"] / 1e6, 1),
                      "queries_per_gpu": n, "branching_frac": n_br / n,
                      "query_mix_present_mismatch_random": MIX, "l2": "inputs larger than L2 (no flush needed)",
                      "sharding": f"arena replicated, queries sharded x{run.world}"},
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches * run.world, "roofline": roofline, "cpu_baseline": cpu}
    eng.close()
    del q, d_succ, d_pred
    torch.cuda.empty_cache()
    return rec


RECORD = {"kmers": kmers_record, "sequences": sequences_record, "branching": branching_record}


def engine_arm(args):
    run = Run(args)
    sp = specs()
    main_tag = args.config
    if args.genome_len:  # explicit override (development only): named in the config, never silent
        sp[main_tag] = dict(sp[main_tag], L=args.genome_len,
                            workload=sp[main_tag]["workload"] + f" [genome length overridden to {args.genome_len}]")
    t0 = time.time()
    main = RECORD[sp[main_tag]["kind"]](run, main_tag, sp[main_tag], True)
    subs = {}
    for tag in [t for t in args.sub.split(",") if t]:
        if tag == main_tag or tag not in sp:
            continue
        t1 = time.time()
        try:
            subs[tag] = RECORD[sp[tag]["kind"]](run, tag, sp[tag], False)
            subs[tag]["wall_s"] = round(time.time() - t1, 1)
        except FileNotFoundError as e:  # the BFT of this config did not travel: say so, never substitute silently
            subs[tag] = {"unavailable": str(e)}
    if run.rank == 0:
        line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": run.world, "steps": main["steps"], "warmup": main["warmup"],
                "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic", "config": main["config"], "clocks": main["clocks"], "e2e": main["e2e"],
                "gpu_launches": main["gpu_launches"], "roofline": main["roofline"], "cpu_baseline": main["cpu_baseline"],
                "configs": subs, "wall_s": round(time.time() - t0, 1)}
        emit(line)
    run.sampler.stop()
    if run.dist:
        run.dist.destroy_process_group()


def reference_arm(args):
    """The reference's own CPU implementation of the path (isKmerPresent + get_annotation + get_list_id_genomes under
    OpenMP, one copy_BFT_Root per thread) on all host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import bench_workloads as wl
    spec = specs()["c3"]
    cfg, k, L = spec["cfg"], spec["k"], args.genome_len or spec["L"]
    path = wl.bft_path(cfg, k, L)
    if not (os.path.exists(path) or os.path.exists(path + ".xz")) and not args.allow_build:
        raise SystemExit(f"bench.py --impl reference: {path}[.xz] is absent (tools/build_bench_data.py c3)")
    genomes = wl.pangenome(cfg, L)
    bft = wl.ensure_bft(cfg, k, L, genomes)
    cores = os.cpu_count() or 1
    n = args.ref_sample
    cat, starts, lens = wl.genomes_to_torch(genomes, torch.device("cpu"))
    q = wl.gen_kmer_queries(cat, starts, lens, k, n, seed=777, mix=MIX)[0].numpy()
    qfile = os.path.join("/tmp", f"bft_bench_ref_{os.getpid()}.kc")
    write_query_file(qfile, q, k)
    secs = run_ref(["kmers", bft, qfile, qfile + ".out", str(cores), str(args.warmup + args.steps)])
    os.remove(qfile)
    os.remove(qfile + ".out")
    timed = secs[args.warmup:]
    ms = 1e3 * sum(timed) / len(timed)
    value = n / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": spec["workload"], "k": k, "n_genomes": cfg["n_genomes"], "genome_len": L, "query_mix_present_mismatch_random": MIX,
                       "sample": f"{n} k-mers per step (same generator and mix as the GPU batch)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"{n} k-mers per step (same generator and mix as the GPU batch), all {cores} host threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def graph_workload(args):
    """Informational single-GPU line for the traversal snippets: build the device graph of the
    100-genome BFT, count its connected components, extract its simple paths; beside it the reference's own
    get_nb_connected_component(BFS) / extract_simple_core_paths_to_disk on one host core (they cannot thread)."""
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
        t["build"].append(t1 - t0)
        t["components"].append(t2 - t1)
        t["paths"].append(t3 - t2)
    launches = eng.launch_count() - l0
    ms = {kk: 1e3 * sum(v) / len(v) for kk, v in t.items()}
    step_ms = ms["build"] + ms["components"]
    cpu = None
    ref_graph = os.path.join(os.path.dirname(wl.REF_HARNESS), "ref_graph")
    if not args.no_cpu_baseline and os.access(ref_graph, os.X_OK):
        t0 = time.perf_counter()
        out = subprocess.run([ref_graph, "components", bft, "bfs"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        dt = time.perf_counter() - t0
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
    ap.add_argument("--config", default="c3", choices=["c1", "c2", "c3", "c4", "c5", "deep"],
                    help="the config reported as the main line (default c3 = BASELINE config[2], the headline metric)")
    ap.add_argument("--sub", default=None, help="comma-separated configs measured as sub-records (default: c1,c2,c4,c5,deep when --config c3)")
    ap.add_argument("--sub-steps", type=int, default=5, help="timed steps of every sub-record")
    ap.add_argument("--genome-len", type=int, default=0, help="development only: override the genome length of the main config")
    ap.add_argument("--queries-per-gpu", type=int, default=0, help="override the k-mers per GPU per step of the main config")
    ap.add_argument("--ref-sample", type=int, default=1 << 24, help="k-mers per step of the CPU reference legs")
    ap.add_argument("--allow-build", action="store_true", help="build a missing BFT with the reference instead of failing")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-probe", action="store_true")
    ap.add_argument("--workload", default="kmers", choices=["kmers", "graph"], help="graph = informational traversal line")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.sub is None:
        args.sub = "c1,c2,c4,c5,deep" if args.config == "c3" else ""
    if args.workload == "graph":
        graph_workload(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        engine_arm(args)


if __name__ == "__main__":
    main()
