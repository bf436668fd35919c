#!/bin/bash
# r2z: GPU suite and the bench lines of the four workloads with the default bank of 2^24 packets (kernels unchanged since r2v).
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -3 gpurun_out/${TAG}_gpu_tests.log
for c in cfg2 cfg1 cfg4 cfg5; do
  timeout 1200 python bench.py --config $c > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
  echo "bench $c rc=$?"; cut -c1-160 gpurun_out/${TAG}_bench_$c.json
done
