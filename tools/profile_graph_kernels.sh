#!/bin/bash
# ncu --set full captures of the traversal kernels and the record-format query kernel on the 100-genome, 1 Mbp BFT;
# run under gpurun, outputs in gpurun_out/ (text summaries next to the .ncu-rep files).
set -x
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k "regex:k_graph_adjacency|k_graph_hook|k_graph_labels" -s 6 -c 3 -o gpurun_out/prof_graph python bench.py --workload graph --steps 1 --no-cpu-baseline > /dev/null 2>&1
$NCU -k "regex:k_paths_rank_step|k_paths_write|k_paths_link" -s 12 -c 4 -o gpurun_out/prof_paths python bench.py --workload graph --steps 1 --no-cpu-baseline > /dev/null 2>&1
$NCU -k regex:k_query_records -s 2 -c 1 -o gpurun_out/prof_records python bench.py --genome-len 1000000 --queries-per-gpu 20000000 --steps 2 --no-cpu-baseline --no-probe > /dev/null 2>&1
for f in graph paths records; do
  ncu -i gpurun_out/prof_$f.ncu-rep --page details --print-units base > gpurun_out/prof_$f.txt 2>&1
done
ls -la gpurun_out/prof_graph* gpurun_out/prof_paths* gpurun_out/prof_records*
