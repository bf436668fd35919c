#!/bin/bash
# GPU parity tests of the several-component build, kernel variants on cfg2, and the cost of a second material mix.
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -15 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python scripts/tune.py 5e7 --config cfg2 > gpurun_out/${TAG}_tune_cfg2.log 2>&1
cut -c1-420 gpurun_out/${TAG}_tune_cfg2.log
SK_BENCH_SECOND_MIX=1 timeout 600 python scripts/tune.py 5e7 --config cfg2 b_cur > gpurun_out/${TAG}_tune_cfg2_second_mix.log 2>&1
cut -c1-420 gpurun_out/${TAG}_tune_cfg2_second_mix.log
