#!/bin/bash
# Round-2 2-GPU pass: scaling of the headline bench at N=1 and N=2 (the driver's torchrun command), NCCL/peer tests, host path probe.
set -x
O=gpurun_out/r2d
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --sub "" --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --sub "" --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err
tail -3 $O/bench_n2.err
timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpus" > $O/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_2gpu.log
tail -3 $O/pytest_2gpu.log
timeout 300 python tools/host_path_probe.py > $O/hostpath_n1.json 2> $O/hostpath_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/host_path_probe.py > $O/hostpath_n2.json 2> $O/hostpath_n2.err
ls -la $O
