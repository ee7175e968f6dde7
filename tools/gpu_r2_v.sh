#!/bin/bash
# last pass: smoke(), the sequence parity tests, c2 with the sequence kernel held at 12 blocks per SM
O=gpurun_out/r2v
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q -k "sequence or fasta or file_drivers or dropin or device_resident" > $O/pytest_seq.log 2>&1; echo "pytest rc=$?" >> $O/pytest_seq.log
tail -3 $O/pytest_seq.log
timeout 600 python bench.py --config c2 --steps 10 --sub "" --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2v/bench_c2.json").read())
print("c2 %.2f G windows/s %.3f ms e2e %.2f G" % (d["value"]/1e9, d["ms_per_step"], d["e2e"]["value"]/1e9))
PY
