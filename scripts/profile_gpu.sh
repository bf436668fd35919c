#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full ncu capture of the trace kernels of one mid-run round.
# usage: [CFG=cfg1|cfg2|cfg4|cfg5] scripts/profile_gpu.sh <tag> [packets] [kernel regex] [skip] [count]
TAG=${1:-r2}
PK=${2:-2e7}
KR=${3:-sk_wf_trace}
SKIP=${4:-6}
CNT=${5:-2}
mkdir -p gpurun_out
CFG=${CFG:-cfg2}
CMD="python bench.py --config $CFG --packets $PK --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:$KR" -s $SKIP -c $CNT -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_prof_bench.log 2>&1
ls -la gpurun_out | tail -5
