#!/bin/bash
# Device tessellation: GPU tests, then the cfg5 bench line (5e5 cells built on the device) with the set-up timing.
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -6 gpurun_out/${TAG}_gpu_tests.log
SK_DEBUG_TIMING=1 timeout 1500 python bench.py --config cfg5 > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err
echo "bench cfg5 rc=$?"; grep "build_voronoi" gpurun_out/${TAG}_bench_cfg5.err | head -8; cut -c1-300 gpurun_out/${TAG}_bench_cfg5.json
