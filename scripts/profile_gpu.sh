#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full ncu capture of the three trace kernels of one mid-run round.
# usage: scripts/profile_gpu.sh <tag> [packets]
TAG=${1:-r1}
PK=${2:-4e6}
mkdir -p gpurun_out
CMD="python bench.py --packets $PK --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sk_wf_trace -s 6 -c 3 -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_prof_bench.log 2>&1
ls -la gpurun_out
