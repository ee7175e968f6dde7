#!/bin/bash
# Round-2 (second session) pass 6, one GPU: one BASELINE config at its FULL size (the file is pushed for this call only):
#   tools/gpu_r2_m.sh c5 k_query_kmers_wide 8      1000 colours x 0.5 Mbp, 100 M colour-row retrievals
#   tools/gpu_r2_m.sh c4 k_query_branching 4       -query_branching at k = 63 on the 100 x 5 Mbp pan-genome
set -x
T=$1; K=$2; S=$3
O=gpurun_out/r2m_$T
mkdir -p $O
timeout 1500 python bench.py --config $T --steps 10 --sub "" > $O/bench_${T}_full.json 2> $O/bench_${T}_full.err
tail -3 $O/bench_${T}_full.err | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o $O/prof_$T python bench.py --config $T --steps 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > /dev/null 2> $O/ncu_$T.err
python - "$O/bench_${T}_full.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print("value %.2f G ms %.3f e2e %.2f G cpu %.2f M degraded=%s frac %.3f" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["cpu_baseline"]["value"]/1e6, d["config"].get("degraded"), d["roofline"]["frac"]))
PY
ls -la $O
