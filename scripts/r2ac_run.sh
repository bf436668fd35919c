#!/bin/bash
# r2ac: sk_engine_prepare_secondary with page-locked staging of its per-cell arrays: dust-emission parity tests and the cfg4 line.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "dust_emission or cfg4s or cfg12me or cfg17c or cfg7v or cfg16d or dynamic" > gpurun_out/r2ac_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2ac_tests.log
timeout 1200 python bench.py --config cfg4 > gpurun_out/r2ac_bench_cfg4.json 2> gpurun_out/r2ac_bench_cfg4.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ac_bench_cfg4.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], {k:d["parity"][k] for k in ("max_sigma","rms_sigma","bins","pass")})
print({k:v for k,v in d.items() if "host" in k or "prepare" in k})
PY
