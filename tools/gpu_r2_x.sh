#!/bin/bash
# C5 at full size: class-row loads without the evict-last priority when the table (127 MB) cannot stay in L2
O=gpurun_out/r2x
mkdir -p $O
timeout 600 python bench.py --config c5 --steps 10 --sub "" --no-cpu-baseline --no-e2e --no-probe > $O/bench_c5_full.json 2> $O/bench_c5_full.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2x/bench_c5_full.json").read())
print("c5 full %.2f G k-mers/s %.3f ms degraded=%s" % (d["value"]/1e9, d["ms_per_step"], d["config"].get("degraded")))
PY
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "kmer_presence_and_colours" 2>&1 | tail -2
