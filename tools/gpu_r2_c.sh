#!/bin/bash
# Round-2 third GPU pass (1 GPU): full bench with all configs; racecheck/synccheck of the TMA-staged records kernel; deep parity.
set -x
O=gpurun_out/r2c
mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
tail -12 $O/bench_n1.err
timeout 600 python -m pytest tests -m gpu -x -q -k "deep or lowcomplex or golden_kmers or golden_branching" > $O/pytest_deep.log 2>&1; echo "pytest rc=$?" >> $O/pytest_deep.log
tail -3 $O/pytest_deep.log
for tool in racecheck synccheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_golden.py -x -q -k "test_reference_record_format and pan_k27" > $O/sanitizer_records_$tool.log 2>&1
  tail -2 $O/sanitizer_records_$tool.log
done
BFT_B200_KF_BITS=0 timeout 300 python bench.py --config c4 --steps 5 --no-e2e --no-cpu-baseline --sub "" > $O/ab_c4_nofilter.json 2> $O/ab_c4_nofilter.err
BFT_B200_KF_BITS=0 timeout 300 python bench.py --config c2 --steps 5 --no-e2e --no-cpu-baseline --sub "" > $O/ab_c2_nofilter.json 2> $O/ab_c2_nofilter.err
BFT_B200_KF_BITS=0 timeout 300 python bench.py --config c1 --steps 5 --no-e2e --no-cpu-baseline --sub "" > $O/ab_c1_nofilter.json 2> $O/ab_c1_nofilter.err
ls -la $O
