#!/usr/bin/env python
"""Build (with the UNMODIFIED reference, oracle/_ref/bft) and cache under data/ the BFTs of the BASELINE configs.
Bench/test infrastructure; run in this container only (the GPU box receives the cached files).

  python tools/build_bench_data.py c1 c2 c3 c4 c5 deep      # any subset
"""
import os
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
}

if __name__ == "__main__":
    for name in sys.argv[1:]:
        cfg, k, L = SPECS[name]
        print(name, wl.ensure_bft(cfg, k, L), flush=True)
