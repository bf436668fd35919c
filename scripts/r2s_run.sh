#!/bin/bash
# r2s: the GPU tests of the build with kinematics, the default bench line (cfg2) and the two-mix diagnostic line of the
# several-component kernels, which now also carry the per-cell look-ups.
TAG=${1:-r2s}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -25 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
echo "bench cfg2 rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_cfg2.json
SK_BENCH_SECOND_MIX=1 timeout 600 python bench.py --no-cpu-baseline --no-parity > gpurun_out/${TAG}_bench_cfg2_second_mix.json 2> gpurun_out/${TAG}_bench_cfg2_second_mix.err
cut -c1-200 gpurun_out/${TAG}_bench_cfg2_second_mix.json
