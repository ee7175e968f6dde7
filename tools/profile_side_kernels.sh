#!/bin/bash
# ncu --set full captures of the kernels that are not the headline one; run under gpurun, outputs in gpurun_out/.
set -x
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:k_query_sequences -s 3 -c 1 -o gpurun_out/prof_sequences python bench.py --workload sequences --steps 2 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:k_query_branching -s 3 -c 1 -o gpurun_out/prof_branching python bench.py --workload branching --steps 2 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:k_expand_rows_v4 -s 3 -c 1 -o gpurun_out/prof_expand_rows python bench.py --pangenome c5 --queries-per-gpu 20000000 --steps 2 --no-e2e --no-cpu-baseline --no-probe > /dev/null 2>&1
$NCU -k regex:k_query_kmers -s 3 -c 1 -o gpurun_out/prof_kmers_c5 python bench.py --pangenome c5 --queries-per-gpu 20000000 --steps 2 --no-e2e --no-cpu-baseline --no-probe > /dev/null 2>&1
$NCU -k "regex:k_extract|k_decode" -c 3 -o gpurun_out/prof_extract python - <<'PY'
import sys
sys.path.insert(0, ".")
from bloomfiltertrie_b200 import engine
import bench_workloads as wl
eng = engine.BFTEngine(wl.ensure_bft(wl.C5, 27, 100_000))
km, cls, _ = eng.extract_kmers()
print(len(km))
PY
python bench.py --workload branching > gpurun_out/bench_br2.json 2>/dev/null
ls -la gpurun_out/*.ncu-rep
