#!/bin/bash
# Round-2 GPU check (run under gpurun): parity tests, then the kernel variants of variants/ on cfg2.
# usage: scripts/r2_check.sh <tag> [packets]
TAG=${1:-r2a}
PK=${2:-5e7}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -5 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python scripts/tune.py $PK > gpurun_out/${TAG}_tune.log 2>&1
cat gpurun_out/${TAG}_tune.log | cut -c1-400
