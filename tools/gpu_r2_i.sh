#!/bin/bash
# Round-2 (second session) pass 2, one GPU: traffic-mix probe variants (which of the L2 / HBM accesses of the headline kernel
# cost what), gather probe by table size, and the fused wide-row kernel on the 1000-colour BFT (100 kbp fallback).
set -x
O=gpurun_out/r2i
mkdir -p $O
timeout 300 tools/mix_probe > $O/mix_probe_c3.txt 2>&1
timeout 300 tools/mix_probe 125000000 771 25 0.54 0.51 > /dev/null 2>&1
timeout 300 tools/mix_probe 125000000 385 25 > $O/mix_probe_half_arena.txt 2>&1
for g in 0.25 0.75 4; do timeout 300 tools/gather_probe $g > $O/gather_probe_${g}GiB.txt 2>&1; done
timeout 900 python bench.py --config c5 --steps 10 --sub "" > $O/bench_c5_fb.json 2> $O/bench_c5_fb.err
tail -3 $O/bench_c5_fb.err | cut -c1-300
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_query_kmers_wide -s 8 -c 1 -o $O/prof_c5_wide python bench.py --config c5 --steps 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > /dev/null 2> $O/ncu_c5.err
cat $O/mix_probe_c3.txt
ls -la $O
