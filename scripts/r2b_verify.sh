#!/bin/bash
# Verification run of the restored tree on one B200: GPU parity tests, the default bench line, the drop-in's log.
TAG=${1:-r2b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv,noheader > gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -4 gpurun_out/${TAG}_gpu_tests.log
SK_KEEP_LOG=gpurun_out/${TAG}_skirt_b200_cfg2_log.txt timeout 1200 python bench.py > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_cfg2.json
