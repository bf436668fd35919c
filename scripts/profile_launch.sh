#!/bin/bash
# full ncu capture of a mid-run launch kernel (free slots scattered over the bank) and the advance kernel before it
TAG=${1:-r1}
PK=${2:-4e7}
mkdir -p gpurun_out
CMD="python bench.py --packets $PK --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:'sk_wf_(launch|advance)' -s 24 -c 2 -o gpurun_out/${TAG}_launch -f $CMD > gpurun_out/${TAG}_launch_bench.log 2>&1
