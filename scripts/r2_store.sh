#!/bin/bash
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -3 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python scripts/tune.py 4e6 --config cfg4 > gpurun_out/${TAG}_tune_cfg4.log 2>&1
cut -c1-420 gpurun_out/${TAG}_tune_cfg4.log
timeout 600 python scripts/tune.py 1e7 --config cfg1 > gpurun_out/${TAG}_tune_cfg1.log 2>&1
cut -c1-420 gpurun_out/${TAG}_tune_cfg1.log
