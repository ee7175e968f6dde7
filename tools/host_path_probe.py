#!/usr/bin/env python
"""Host-path probe: what the PCIe / host-memory side of the box sustains when 1, 2, 4 or 8 ranks copy at
once — the bound of the host-facing (e2e) calls, which move 7 B in and ~7 B out per k-mer while the kernels need 0.02 ns.

  torchrun --nproc-per-node N tools/host_path_probe.py      (or plain python for N = 1)

Every rank copies a 1 GiB pinned host buffer to its GPU and a 1 GiB device buffer back to pinned memory, (a) one direction
at a time, (b) both directions at once on two streams, ITER times, all ranks started together. Pinned buffers are
allocated twice: as the process comes up, and after the process pinned itself to the CPUs `nvidia-smi topo -m` lists for its
GPU (first touch then lands on the GPU's NUMA node). Prints one JSON line on rank 0 with per-rank and aggregate GB/s.
"""
import json
import os
import re
import subprocess
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GIB = 1 << 30
ITER = 8


def gpu_cpu_affinity(index):
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, text=True, timeout=30).stdout
    except Exception:
        return None, None
    for line in out.splitlines():
        f = line.split()
        if f and f[0] == f"GPU{index}":
            m = [x for x in f[1:] if re.fullmatch(r"[0-9,\-]+", x) and ("-" in x or "," in x)]
            numa = [x for x in f[1:] if re.fullmatch(r"[0-9]+", x)]
            if m:
                cpus = set()
                for part in m[0].split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
                return sorted(cpus), (numa[0] if numa else None)
    return None, None


def measure(dev, world, label):
    h_in = torch.empty(GIB, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(GIB, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(0)   # touch
    d_in = torch.empty(GIB, dtype=torch.uint8, device=dev)
    d_out = torch.ones(GIB, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    res = {}
    for mode in ("h2d", "d2h", "both"):
        for _ in range(2):
            d_in.copy_(h_in, non_blocking=True)
            h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(ITER):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        nbytes = ITER * GIB * (2 if mode == "both" else 1)
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = {"per_rank_gbs": nbytes / dt / 1e9, "aggregate_gbs": world * nbytes / float(t.item()) / 1e9}
    del h_in, h_out, d_in, d_out
    return {label: res}


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {"n_ranks": world, "bytes_per_copy": GIB, "iterations": ITER, "host_cpus": os.cpu_count()}
    out.update(measure(dev, world, "default_placement"))
    cpus, numa = gpu_cpu_affinity(local)
    out["rank0_gpu_cpu_affinity"] = f"{cpus[0]}-{cpus[-1]}" if cpus else None
    out["rank0_gpu_numa"] = numa
    if cpus:
        try:
            os.sched_setaffinity(0, cpus)
            out.update(measure(dev, world, "numa_local_placement"))
        except OSError as e:
            out["numa_local_placement"] = {"error": str(e)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
