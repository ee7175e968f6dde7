#!/bin/bash
# Round-2 multi-GPU pass: the driver's exact bench command (all configs as sub-records) under torchrun at N ranks.
set -x
N=${1:-2}
O=gpurun_out/r2g_n$N
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
tail -5 $O/bench.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 tools/host_path_probe.py > $O/hostpath.json 2> $O/hostpath.err
ls -la $O
