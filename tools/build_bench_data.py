#!/usr/bin/env python
"""Build (with the UNMODIFIED reference, oracle/_ref/bft) and cache under data/ the BFTs of the BASELINE configs.
Bench/test infrastructure; run in this container only (the GPU box receives the cached files).

  python tools/build_bench_data.py c1 c2 c3 c4 c5 deep c4_fb c5_fb c3_1mbp     # any subset

Every .bft is also compressed to the .bft.xz that travels with the repo snapshot (the raw files are gpurun-ignored).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_workloads as wl

SPECS = {
    "c1": (wl.C1, 27, 5_000_000),
    "c2": (wl.C2, 27, 5_000_000),
    "c3": (wl.C3, 27, 5_000_000),
    "c4": (wl.C3, 63, 5_000_000),
    "c5": (wl.C5, 27, 500_000),
    "deep": (wl.DEEP, 63, 0),
    # the smaller BFTs of the same generators that the default snapshot carries for c4 / c5 (bench.py flags them `degraded`),
    # and the 1 Mbp pan-genome of the informational graph workload
    "c4_fb": (wl.C3, 63, 1_000_000),
    "c5_fb": (wl.C5, 27, 100_000),
    "c3_1mbp": (wl.C3, 27, 1_000_000),
}

if __name__ == "__main__":
    for name in sys.argv[1:]:
        cfg, k, L = SPECS[name]
        path = wl.ensure_bft(cfg, k, L)
        if not os.path.exists(path + ".xz"):
            subprocess.run(["xz", "-T0", "-6", "-k", path], check=True)
        print(name, path, os.path.getsize(path + ".xz") >> 20, "MiB compressed", flush=True)
