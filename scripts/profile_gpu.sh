#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the life-cycle kernel.
# usage: scripts/profile_gpu.sh <tag> [packets]
TAG=${1:-r1}
PK=${2:-2e6}
mkdir -p gpurun_out
CMD="python bench.py --packets $PK --steps 2 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sk_life_cycle -s 1 -c 1 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_prof_bench.log 2>&1
ls -la gpurun_out
