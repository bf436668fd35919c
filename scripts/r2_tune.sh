#!/bin/bash
TAG=${1:-r2n}; CFG=${2:-cfg2}; PK=${3:-5e7}
mkdir -p gpurun_out
timeout 1500 python scripts/tune.py $PK --config $CFG > gpurun_out/${TAG}_tune_${CFG}.log 2>&1
cut -c1-420 gpurun_out/${TAG}_tune_${CFG}.log
