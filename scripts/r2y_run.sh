#!/bin/bash
# r2y: bank capacity sweep on cfg2 (packets in flight per GPU; SK_BANK), kernels only.
mkdir -p gpurun_out
for b in 4194304 8388608 12582912 16777216 33554432; do
  SK_BANK=$b timeout 600 python bench.py --no-cpu-baseline --no-parity --no-e2e --steps 3 --warmup 3 > gpurun_out/r2y_bank_$b.json 2> gpurun_out/r2y_bank_$b.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2y_bank_$b.json").read().strip().splitlines()[-1])
print($b, d["value"], d["ms_per_step"], {k:round(v,1) for k,v in d["kernel"]["stage_ms_per_step"].items() if v}, d["kernel"]["launches_per_step"])
PY
done
