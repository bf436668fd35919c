#!/bin/bash
# GPU parity tests (contiguous Voronoi neighbour records, several components), Voronoi kernel variants, two-mix diagnostic.
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -6 gpurun_out/${TAG}_gpu_tests.log
timeout 1200 python scripts/tune_voronoi.py 2e5 1e7 > gpurun_out/${TAG}_tune_voronoi.log 2>&1
cut -c1-420 gpurun_out/${TAG}_tune_voronoi.log
SK_BENCH_SECOND_MIX=1 SK_ENGINE_LIB= timeout 600 python - > gpurun_out/${TAG}_second_mix.log 2>&1 <<'P'
import sys, json
sys.path.insert(0, ".")
import bench
from skirt9_b200 import abi
for second in (False, True):
    import os
    if not second: os.environ.pop("SK_BENCH_SECOND_MIX", None)
    else: os.environ["SK_BENCH_SECOND_MIX"] = "1"
    sim = bench.make_sim("cfg2", 50000000)
    e = sim.configure(abi.Engine(sim.config_struct(device=0)))
    out = []
    for k in range(3):
        e.clear_instruments(); sim.run(e, stream_id=k); out.append(e.last_kernel_ms())
    c = e.counters()
    print(json.dumps({"second_mix": second, "ms": out, "stages": {k: round(v, 1) for k, v in e.last_stage_ms().items() if v},
                      "crossings": (c["forward_segments"] + c["peel_segments"] + c["replay_segments"]) / 3}), flush=True)
    e.close()
P
cut -c1-420 gpurun_out/${TAG}_second_mix.log
