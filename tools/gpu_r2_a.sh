#!/bin/bash
# Round-2 first GPU pass: parity suite, headline bench, filter / L2-persistence A/B, ncu captures. Run under gpurun (1 GPU).
set -x
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
Q="--steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-probe"
BFT_B200_KF_BITS=0 timeout 300 python bench.py $Q > $O/ab_nofilter.json 2> $O/ab_nofilter.err
BFT_B200_KF_BITS=0 BFT_B200_NO_L2_PERSIST=1 timeout 300 python bench.py $Q > $O/ab_nofilter_nopersist.json 2>> $O/ab_nofilter.err
BFT_B200_NO_L2_PERSIST=1 timeout 300 python bench.py $Q > $O/ab_filter8_nopersist.json 2> $O/ab_f8np.err
BFT_B200_KF_BITS=6 timeout 300 python bench.py $Q > $O/ab_filter6.json 2> $O/ab_f6.err
BFT_B200_KF_BITS=11 timeout 300 python bench.py $Q > $O/ab_filter11.json 2> $O/ab_f11.err
BFT_B200_KF_BITS=14 BFT_B200_KF_MAX_MB=64 timeout 300 python bench.py $Q > $O/ab_filter14.json 2> $O/ab_f14.err
# ncu: launch list of the bench command, then one --set full capture of the dominant kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-probe > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_query_kmers_rows -s 4 -c 2 -o $O/prof_kmers_rows python bench.py $Q > /dev/null 2> $O/ncu_full.err
ls -la $O
