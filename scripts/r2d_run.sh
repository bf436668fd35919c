#!/bin/bash
# GPU parity tests, advance-kernel variants on cfg2, trace-kernel variants on the two-mix diagnostic workload.
TAG=${1:-r2d}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -6 gpurun_out/${TAG}_gpu_tests.log
timeout 900 python scripts/tune.py 5e7 --config cfg2 b_cur j_adv1 k_adv4 l_adv5 > gpurun_out/${TAG}_tune_cfg2.log 2>&1
cut -c1-420 gpurun_out/${TAG}_tune_cfg2.log
SK_BENCH_SECOND_MIX=1 timeout 600 python scripts/tune.py 5e7 --config cfg2 b_cur m_multi1 n_multi2 > gpurun_out/${TAG}_tune_cfg2_second_mix.log 2>&1
cut -c1-420 gpurun_out/${TAG}_tune_cfg2_second_mix.log
