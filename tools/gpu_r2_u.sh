#!/bin/bash
# paired-bucket arena: whole -m gpu suite first; only if it is green, the full pass-4 measurement (bench line, reference arm, ncu
# captures, launch list) on the same box
O=gpurun_out/r2u
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
if [ $rc -ne 0 ]; then grep -E "Error|error|assert|FAILED" $O/pytest_gpu.log | head -20; exit 1; fi
bash tools/gpu_r2_k.sh 2>&1 | grep -v "^+"
