#!/bin/bash
# sequence kernel pass: parity of everything that touches sequences, then c2 (and an ncu capture of it)
set -x
O=gpurun_out/r2s
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -k "sequence or golden or fasta or file_drivers or dropin or accelerator" > $O/pytest_seq.log 2>&1; echo "pytest rc=$?" >> $O/pytest_seq.log
tail -3 $O/pytest_seq.log
timeout 600 python bench.py --config c2 --steps 10 --sub "" > $O/bench_c2.json 2> $O/bench_c2.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_query_sequences -s 4 -c 1 -o $O/prof_c2 python bench.py --config c2 --steps 3 --no-e2e --no-cpu-baseline --no-probe --sub "" > /dev/null 2> $O/ncu_c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2s/bench_c2.json").read())
print("c2 %.2f G windows/s %.3f ms e2e %.2f G cpu %.1f M" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9, d["cpu_baseline"]["value"]/1e6))
PY
