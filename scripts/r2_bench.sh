#!/bin/bash
# Round-2 GPU run (under gpurun): parity tests, then bench.py on the workloads.  usage: scripts/r2_bench.sh <tag> [configs...]
TAG=${1:-r2}
shift
CFGS=${@:-cfg2 cfg1 cfg4}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -3 gpurun_out/${TAG}_gpu_tests.log
for c in $CFGS; do
  timeout 900 python bench.py --config $c > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
  echo "$c rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_$c.err; cut -c1-700 gpurun_out/${TAG}_bench_$c.json
done
