#!/bin/bash
# Run on the GPU box (under gpurun): full ncu capture of the event kernels (launch, advance, detect, sample) of an
# early round with a full bank.  usage: scripts/profile_events.sh <tag> [packets]
TAG=${1:-r1}
PK=${2:-2e7}
mkdir -p gpurun_out
CMD="python bench.py --packets $PK --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:'sk_wf_(launch|advance|detect|sample)' -s 4 -c 4 -o gpurun_out/${TAG}_events -f $CMD > gpurun_out/${TAG}_events_bench.log 2>&1
