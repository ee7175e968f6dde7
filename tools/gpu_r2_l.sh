#!/bin/bash
# Round-2 (second session) pass 5, one GPU: do the sub-records run slower than the same configs alone because the persisting-L2
# set-aside of the first context (device-wide limit) outlives it? Same bench line with and without the access-policy window.
set -x
O=gpurun_out/r2l
mkdir -p $O
A="--steps 10 --sub c1,c5 --no-cpu-baseline --no-e2e --no-probe"
timeout 900 python bench.py $A > $O/bench_persist.json 2> $O/bench_persist.err
BFT_B200_NO_L2_PERSIST=1 timeout 900 python bench.py $A > $O/bench_nopersist.json 2> $O/bench_nopersist.err
BFT_B200_NO_L2_PERSIST=1 timeout 900 python bench.py --config c1 --steps 10 --sub "" --no-cpu-baseline --no-e2e --no-probe > $O/bench_c1_alone_nopersist.json 2> /dev/null
timeout 900 python bench.py --config c1 --steps 10 --sub "" --no-cpu-baseline --no-e2e --no-probe > $O/bench_c1_alone_persist.json 2> /dev/null
python - <<'PY'
import json
for n in ("persist","nopersist"):
    d=json.loads(open(f"gpurun_out/r2l/bench_{n}.json").read())
    print(n, "c3 %.2f G %.3f ms" % (d["value"]/1e9, d["ms_per_step"]), {t:(round(r["value"]/1e9,2), round(r["ms_per_step"],3)) for t,r in d["configs"].items()})
for n in ("c1_alone_nopersist","c1_alone_persist"):
    d=json.loads(open(f"gpurun_out/r2l/bench_{n}.json").read())
    print(n, "%.2f G %.3f ms" % (d["value"]/1e9, d["ms_per_step"]))
PY
