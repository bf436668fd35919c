#!/bin/bash
TAG=${1:-r2j}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -3 gpurun_out/${TAG}_gpu_tests.log
SK_BENCH_SITES=100000 timeout 1200 python scripts/tune.py 1e7 --config cfg5 > gpurun_out/${TAG}_tune_cfg5.log 2>&1
cut -c1-400 gpurun_out/${TAG}_tune_cfg5.log
timeout 1200 python bench.py --config cfg5 --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err
tail -c 400 gpurun_out/${TAG}_bench_cfg5.err; cut -c1-600 gpurun_out/${TAG}_bench_cfg5.json
