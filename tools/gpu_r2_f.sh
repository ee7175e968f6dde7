#!/bin/bash
# Round-2 sixth GPU pass (1 GPU): two-phase neighbour engine — parity, c4 / c2 / c3 timings, ncu capture of k_query_branching.
set -x
O=gpurun_out/r2f
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -k "branching or graph or neighbor or golden or dropin or snippets or kmer_presence" > $O/pytest_sub.log 2>&1; echo "pytest rc=$?" >> $O/pytest_sub.log
tail -4 $O/pytest_sub.log
timeout 600 python bench.py --config c4 --steps 10 --sub "" > $O/bench_c4.json 2> $O/bench_c4.err
timeout 600 python bench.py --config c2 --steps 10 --sub "" --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
timeout 600 python bench.py --config c3 --steps 10 --sub "" --no-cpu-baseline --no-e2e > $O/bench_c3.json 2> $O/bench_c3.err
timeout 600 python bench.py --workload graph --genome-len 5000000 --steps 3 --no-cpu-baseline > $O/bench_graph.json 2> $O/bench_graph.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_query_branching -s 4 -c 1 -o $O/prof_c4 python bench.py --config c4 --steps 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > /dev/null 2> $O/ncu_c4.err
ls -la $O
