#!/usr/bin/env python
"""Kernel-variant tuning on the GPU box: one host setup of the workload, then every variants/lib_*.so timed on the same input.
usage: python scripts/tune.py [packets] [--config cfg1|cfg2|cfg4|cfg5] [names...]"""
import glob, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skirt9_b200 import abi
import bench

argv = sys.argv[1:]
config = "cfg2"
if "--config" in argv:
    i = argv.index("--config")
    config = argv[i + 1]
    del argv[i:i + 2]
packets = int(float(argv[0])) if argv else 20000000
names = argv[1:]
libs = sorted(glob.glob(os.path.join(ROOT, "variants", "lib_*.so")))
if names:
    libs = [l for l in libs if os.path.basename(l)[4:-3] in names]
if os.environ.get("SK_TUNE_CHILD") != "1" and len(libs) > 1:
    # one process per variant: the libraries export the same symbols, and in-library calls bind to the first one loaded
    import subprocess
    for path in libs:
        subprocess.call([sys.executable, __file__, str(packets), "--config", config, os.path.basename(path)[4:-3]],
                        env=dict(os.environ, SK_TUNE_CHILD="1"))
    sys.exit(0)
sim = bench.make_sim(config, packets)
for path in libs:
    lib = abi.load_engine_library(path)
    e = sim.configure(abi.Engine(sim.config_struct(device=0), lib=lib))
    acc = {"ms": 0.0, "stages": {}}
    inner = e.run_segment

    def run_segment(*a, **kw):
        inner(*a, **kw)
        acc["ms"] += e.last_kernel_ms()
        for k, v in e.last_stage_ms().items():
            acc["stages"][k] = acc["stages"].get(k, 0.0) + v
    e.run_segment = run_segment
    ms = []
    for k in range(3):
        e.clear_instruments()
        acc["ms"], acc["stages"] = 0.0, {}
        sim.run(e, stream_id=k)
        ms.append(acc["ms"])
    stages = {k: round(v, 1) for k, v in acc["stages"].items() if v}
    c = e.counters()
    print(json.dumps({"variant": os.path.basename(path)[4:-3], "config": config, "packets": packets, "ms": ms,
                      "pkt_per_s": c["packets"] / 3 / (min(ms[1:]) * 1e-3), "stages_ms": stages, "rounds": c["rounds"] / 3}), flush=True)
    e.close()
