#!/bin/bash
# r2w: the bench lines with the parity block on the reference's own set-up and N = launched packets in the error estimate.
TAG=${1:-r2w}
mkdir -p gpurun_out
for c in cfg2 cfg4 cfg1; do
  timeout 1200 python bench.py --config $c > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
  echo "bench $c rc=$?"; tail -2 gpurun_out/${TAG}_bench_$c.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_$c.json").read().strip().splitlines()[-1])
p=d.get("parity") or {}
print("$c", d["value"], d["ms_per_step"], {k:p.get(k) for k in ("against","max_sigma","rms_sigma","bins","bins_over_4_sigma","pass","bound_sigma","max_sigma_per_component","bound_sigma_per_component","reference_vs_reference")})
PY
done
