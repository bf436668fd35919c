#!/bin/bash
# Multi-GPU run (gpurun --gpus N): multi-GPU tests, then bench.py under torchrun on cfg2 and cfg4.
# usage: r2_multi.sh <tag> <N> [notests]
TAG=${1:-r2m}
N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt
if [[ "$3" != notests ]]; then
timeout 1500 python -m pytest tests/test_multigpu_nccl.py tests/test_shim_ski.py -m gpu -q -k "two" > gpurun_out/${TAG}_multigpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_multigpu_tests.log
tail -4 gpurun_out/${TAG}_multigpu_tests.log
fi
for c in cfg2 cfg4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --config $c --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_${c}_${N}gpu.json 2> gpurun_out/${TAG}_bench_${c}_${N}gpu.err
  echo "$c rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_${c}_${N}gpu.err; cut -c1-400 gpurun_out/${TAG}_bench_${c}_${N}gpu.json
done
