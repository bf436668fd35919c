#!/bin/bash
# Voronoi plane records: GPU parity tests first, then the kernel variants.
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -6 gpurun_out/${TAG}_gpu_tests.log
timeout 1200 python scripts/tune_voronoi.py 2e5 1e7 > gpurun_out/${TAG}_tune_voronoi.log 2>&1
cut -c1-420 gpurun_out/${TAG}_tune_voronoi.log
