#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -12 gpurun_out/${TAG}_gpu_tests.log
