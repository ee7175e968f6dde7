#!/bin/bash
# Round-2 (second session) pass 4, one GPU: the driver's default bench command (all configs as sub-records), the reference arm,
# ncu --set full captures of every config's dominant kernel on the current build, and the launch list of the bench command.
set -x
O=gpurun_out/r2k
mkdir -p $O
timeout 1500 python bench.py --steps 20 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err
tail -4 $O/bench_n1.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
NCU="ncu --set full --clock-control none --import-source on"
A="--steps 3 --no-e2e --no-cpu-baseline --no-probe --sub"
timeout 600 $NCU -k regex:k_query_kmers_rows -s 4 -c 1 -o $O/prof_c3 python bench.py --config c3 $A "" > /dev/null 2> $O/ncu_c3.err
timeout 600 $NCU -k regex:k_query_kmers_rows -s 4 -c 1 -o $O/prof_c1 python bench.py --config c1 $A "" > /dev/null 2> $O/ncu_c1.err
timeout 600 $NCU -k regex:k_query_sequences -s 4 -c 1 -o $O/prof_c2 python bench.py --config c2 $A "" > /dev/null 2> $O/ncu_c2.err
timeout 600 $NCU -k regex:k_query_branching -s 4 -c 1 -o $O/prof_c4_fb python bench.py --config c4 $A "" > /dev/null 2> $O/ncu_c4.err
timeout 600 $NCU -k regex:k_query_kmers_wide -s 8 -c 1 -o $O/prof_c5_fb python bench.py --config c5 $A "" > /dev/null 2> $O/ncu_c5.err
timeout 600 $NCU -k regex:k_query_kmers_rows -s 4 -c 1 -o $O/prof_deep python bench.py --config deep $A "" > /dev/null 2> $O/ncu_deep.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-probe > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2k/bench_n1.json").read())
print("c3 value %.2f G ms %.3f e2e %.2f G wall %s" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["wall_s"]))
for t,r in d["configs"].items():
    print(t, "unavailable" if "unavailable" in r else "value %.2f G ms %.3f e2e %.2f G cpu %.1f M wall %s degraded=%s" % (r["value"]/1e9, r["ms_per_step"], r["e2e"]["value"]/1e9, r["cpu_baseline"]["value"]/1e6, r["wall_s"], bool(r["config"].get("degraded"))))
PY
ls -la $O
