#!/bin/bash
# r2t: the tests of the dynamic-state iterations on the GPU (engine through the host mirror, drop-in binary), the whole GPU
# suite, and the kinematics diagnostic line (cfg2 with a rotating ring and source).
TAG=${1:-r2t}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -15 gpurun_out/${TAG}_gpu_tests.log
SK_BENCH_KINEMATICS=1 timeout 600 python bench.py --no-cpu-baseline --no-parity > gpurun_out/${TAG}_bench_cfg2_kinematics.json 2> gpurun_out/${TAG}_bench_cfg2_kinematics.err
cut -c1-200 gpurun_out/${TAG}_bench_cfg2_kinematics.json; tail -3 gpurun_out/${TAG}_bench_cfg2_kinematics.err
