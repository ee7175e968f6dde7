#!/bin/bash
# whole -m gpu suite on the current build (no bench data needed)
O=gpurun_out/r2r
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -15 $O/pytest_gpu.log
