#!/bin/bash
# After r2q: the GPU tests of the last commit, the cfg4 bench line with its parity block, the two-mix diagnostic line.
TAG=${1:-r2r}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 1500 python bench.py --config cfg4 > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err
echo "bench cfg4 rc=$?"; cut -c1-200 gpurun_out/${TAG}_bench_cfg4.json
SK_BENCH_SECOND_MIX=1 timeout 600 python bench.py --no-cpu-baseline --no-parity > gpurun_out/${TAG}_bench_cfg2_second_mix.json 2> gpurun_out/${TAG}_bench_cfg2_second_mix.err
cut -c1-200 gpurun_out/${TAG}_bench_cfg2_second_mix.json
