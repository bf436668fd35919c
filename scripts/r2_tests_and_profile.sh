#!/bin/bash
# GPU tests, then an ncu capture of the forward (RF-storing) trace kernel on cfg4
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -3 gpurun_out/${TAG}_gpu_tests.log
CFG=cfg4 scripts/profile_gpu.sh ${TAG}_cfg4 2e6 "sk_wf_trace<2, 0" 4 1
