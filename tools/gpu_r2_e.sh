#!/bin/bash
# Round-2 fifth GPU pass (1 GPU): full parity suite, full bench (all configs + deep), ncu captures of every config's dominant kernel.
set -x
O=gpurun_out/r2e
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
timeout 1200 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
tail -4 $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
NCU="ncu --set full --clock-control none --import-source on"
Q="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-probe --sub ''"
timeout 600 $NCU -k regex:k_query_kmers_rows -s 4 -c 1 -o $O/prof_c3 python bench.py --config c3 --steps 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > /dev/null 2> $O/ncu_c3.err
timeout 600 $NCU -k regex:k_query_kmers_rows -s 4 -c 1 -o $O/prof_c1 python bench.py --config c1 --steps 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > /dev/null 2> $O/ncu_c1.err
timeout 600 $NCU -k regex:k_query_sequences -s 4 -c 1 -o $O/prof_c2 python bench.py --config c2 --steps 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > /dev/null 2> $O/ncu_c2.err
timeout 600 $NCU -k regex:k_query_branching -s 4 -c 1 -o $O/prof_c4 python bench.py --config c4 --steps 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > /dev/null 2> $O/ncu_c4.err
timeout 600 $NCU -k "regex:k_query_kmers|k_expand_rows" -s 8 -c 2 -o $O/prof_c5 python bench.py --config c5 --steps 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > /dev/null 2> $O/ncu_c5.err
timeout 600 $NCU -k regex:k_query_kmers_rows -s 4 -c 1 -o $O/prof_deep python bench.py --config deep --steps 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > /dev/null 2> $O/ncu_deep.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-probe > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
ls -la $O
