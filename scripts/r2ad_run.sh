#!/bin/bash
# r2ad: the last commit -- GPU suite, smoke, the bench lines of cfg2 / cfg4 / cfg1 (cfg4 e2e with the radiation field read into
# a page-locked table of the caller).
TAG=${1:-r2ad}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -3 gpurun_out/${TAG}_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
for c in cfg2 cfg4 cfg1; do
  timeout 1200 python bench.py --config $c > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
  echo "bench $c rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_$c.json").read().strip().splitlines()[-1])
p=d.get("parity") or {}
print("$c", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["parts"], {k:p.get(k) for k in ("max_sigma","rms_sigma","bins","pass")})
PY
done
