#!/bin/bash
# Round-2 second GPU pass (1 GPU): new bench with all configs, L2 fetch-granularity A/B, new parity tests.
set -x
O=gpurun_out/r2b
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "enumeration or fasta or file_drivers or extract or sharded or peer" > $O/pytest_new.log 2>&1; echo "pytest rc=$?" >> $O/pytest_new.log
tail -5 $O/pytest_new.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
tail -3 $O/bench_n1.err
Q="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-probe --sub ''"
for g in 32 64 128; do
  BFT_B200_L2_FETCH=$g timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > $O/ab_l2fetch$g.json 2> $O/ab_l2fetch$g.err
done
BFT_B200_KF_BITS=5 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > $O/ab_filter5.json 2> $O/ab_f5.err
BFT_B200_KF_BITS=0 timeout 300 python bench.py --config c4 --steps 5 --no-e2e --no-cpu-baseline --sub "" > $O/ab_c4_nofilter.json 2> $O/ab_c4_nofilter.err
BFT_B200_KF_BITS=0 timeout 300 python bench.py --config c2 --steps 5 --no-e2e --no-cpu-baseline --sub "" > $O/ab_c2_nofilter.json 2> $O/ab_c2_nofilter.err
ls -la $O
