#!/bin/bash
# Round-2 (second session) pass 1, one GPU, no bench data shipped: the whole -m gpu suite on the current build (fused wide-row
# kernel included), the memory-system probes (random-access rate by table size / width; the traffic mix of the headline kernel
# replayed without the walk) and ncu captures of the probes so the ceilings quoted in DESIGN.md §5 are evidence.
set -x
O=gpurun_out/r2h
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw --format=csv > $O/smi.txt
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 300 tools/mix_probe > $O/mix_probe_c3.txt 2>&1
timeout 300 tools/mix_probe 125000000 771 0 > $O/mix_probe_c3_nofilter_tables.txt 2>&1
timeout 300 tools/mix_probe 125000000 385 25 > $O/mix_probe_half_arena.txt 2>&1
for g in 0.25 0.75 4; do timeout 300 tools/gather_probe $g > $O/gather_probe_${g}GiB.txt 2>&1; done
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_mix -s 30 -c 1 -o $O/prof_mix_all tools/mix_probe > /dev/null 2> $O/ncu_mix.err
timeout 600 $NCU -k regex:k_gather256 -s 6 -c 1 -o $O/prof_gather256 tools/gather_probe 4 > /dev/null 2> $O/ncu_gather.err
timeout 600 $NCU -k regex:k_gather -s 0 -c 1 -o $O/prof_gather8 tools/gather_probe 4 > /dev/null 2> $O/ncu_gather8.err
cat $O/mix_probe_c3.txt
ls -la $O
