#!/bin/bash
# Round-2 (second session) pass 3, one GPU: the fused root directory + filter table (rootkf) — parity, then A/B on c3 / c1 / c2
# (BFT_B200_RKF_SECTORS=0 is the two-table path), and the sequence kernel's early exit.
set -x
O=gpurun_out/r2j
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
B="python bench.py --steps 10 --sub '' --no-cpu-baseline --no-e2e --no-probe"
for s in 0 4 3; do
  BFT_B200_RKF_SECTORS=$s timeout 600 python bench.py --config c3 --steps 10 --sub "" --no-cpu-baseline --no-e2e --no-probe > $O/bench_c3_rkf$s.json 2> $O/bench_c3_rkf$s.err
done
for s in 0 1 2; do
  BFT_B200_RKF_SECTORS=$s timeout 600 python bench.py --config c1 --steps 10 --sub "" --no-cpu-baseline --no-e2e --no-probe > $O/bench_c1_rkf$s.json 2> $O/bench_c1_rkf$s.err
done
timeout 600 python bench.py --config c2 --steps 10 --sub "" --no-e2e > $O/bench_c2.json 2> $O/bench_c2.err
BFT_B200_RKF_SECTORS=0 timeout 600 python bench.py --config c2 --steps 10 --sub "" --no-cpu-baseline --no-e2e > $O/bench_c2_rkf0.json 2> $O/bench_c2_rkf0.err
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    r=d["roofline"]
    print(sys.argv[1].split("/")[-1], "value %.2f G  kernel_ms %.3f  bucket/k %.3f rejects/k %.3f rootkf_mb %s" % (d["value"]/1e9, r["kernel_ms"], r.get("bucket_accesses_per_kmer", r.get("bucket_accesses_per_window", 0)), r.get("filter_rejects_per_kmer", r.get("filter_rejects_per_window", 0)), d["config"].get("rootkf_mb")))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
ls -la $O
