#!/bin/bash
# Round-2 (second session) multi-GPU pass: the driver's torchrun command at N ranks (sub-records limited to what was pushed),
# then the host-path probe (pinned copies, all ranks at once).
#   tools/gpu_r2_p.sh N "c1,c5"
set -x
N=${1:-2}
SUB=${2-}
O=gpurun_out/r2p_n$N
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --sub "$SUB" > $O/bench.json 2> $O/bench.err
tail -3 $O/bench.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/host_path_probe.py > $O/hostpath.json 2> $O/hostpath.err
python - "$O/bench.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read())
print("N=%d value %.2f G ms/step %.3f kernel_ms %.3f e2e %.2f G reduction=%s" % (d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"]/1e9, d["config"]["hit_count_reduction"]))
for t,r in d["configs"].items():
    print(t, "unavailable" if "unavailable" in r else "value %.2f G ms %.3f e2e %.2f G" % (r["value"]/1e9, r["ms_per_step"], r["e2e"]["value"]/1e9))
PY
cat $O/hostpath.json | cut -c1-600
