#!/bin/bash
# Round-2 (second session) pass 8, one GPU: collapsed subtrees (one hashed block per root prefix that leads to child Nodes) —
# the whole parity suite, then A/B (BFT_B200_NO_DEEP=1 walks the Nodes) on the forced-deep trie and on c4 (1 Mbp fallback).
set -x
O=gpurun_out/r2o
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
for t in deep c4; do
  timeout 600 python bench.py --config $t --steps 10 --sub "" --no-cpu-baseline --no-e2e --no-probe > $O/bench_${t}_collapsed.json 2> $O/bench_${t}_collapsed.err
  BFT_B200_NO_DEEP=1 timeout 600 python bench.py --config $t --steps 10 --sub "" --no-cpu-baseline --no-e2e --no-probe > $O/bench_${t}_walk.json 2> $O/bench_${t}_walk.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2o/bench_*.json")):
    try:
        d=json.loads(open(f).read())
        print(f.split("/")[-1], "%.2f G  %.3f ms  collapsed %.0f MB degraded=%s" % (d["value"]/1e9, d["ms_per_step"], d["config"].get("collapsed_subtrees_mb", -1), bool(d["config"].get("degraded"))))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -2 $O/bench_deep_collapsed.err | cut -c1-400
