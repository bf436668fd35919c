#!/bin/bash
# r2u: kinematics kernels after the local table search / 32-byte velocity records / prefetch: parity tests of the kinematics paths,
# then the diagnostic line (cfg2 with a rotating ring and source) for three occupancy variants of the KIN trace kernels.
TAG=${1:-r2u}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "kinematics or velocit or cfg15k" > gpurun_out/${TAG}_gpu_tests_kinematics.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests_kinematics.log
tail -4 gpurun_out/${TAG}_gpu_tests_kinematics.log
for v in kin1 kin2 kin3; do
  SK_ENGINE_LIB=variants/lib_$v.so SK_BENCH_KINEMATICS=1 timeout 600 python bench.py --no-cpu-baseline --no-parity --packets 3e7 > gpurun_out/${TAG}_bench_cfg2_kinematics_$v.json 2> gpurun_out/${TAG}_bench_$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_cfg2_kinematics_$v.json").read().strip().splitlines()[-1])
print("$v", d["value"], d["ms_per_step"], d["kernel"]["stage_ms_per_step"])
PY
done
