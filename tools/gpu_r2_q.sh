#!/bin/bash
# quick A/B pass: the default bench line (all sub-records) without the CPU / e2e / probe legs
set -x
O=gpurun_out/r2q
mkdir -p $O
timeout 900 python bench.py --steps 10 --no-cpu-baseline --no-e2e --no-probe > $O/bench.json 2> $O/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2q/bench.json").read())
print("c3 %.2f G %.3f ms" % (d["value"]/1e9, d["ms_per_step"]), {t:(round(r["value"]/1e9,2), round(r["ms_per_step"],3)) for t,r in d["configs"].items()})
PY
