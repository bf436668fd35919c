#!/usr/bin/env python
"""Kernel-variant tuning on the GPU box: one host setup of cfg2, then every variants/lib_*.so timed on the same input.
usage: python scripts/tune.py [packets] [names...]"""
import glob, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skirt9_b200 import abi, configs

packets = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20000000
names = sys.argv[2:]
libs = sorted(glob.glob(os.path.join(ROOT, "variants", "lib_*.so")))
if names:
    libs = [l for l in libs if os.path.basename(l)[4:-3] in names]
if os.environ.get("SK_TUNE_CHILD") != "1" and len(libs) > 1:
    # one process per variant: the libraries export the same symbols, and in-library calls bind to the first one loaded
    import subprocess
    for path in libs:
        subprocess.call([sys.executable, __file__, str(packets), os.path.basename(path)[4:-3]],
                        env=dict(os.environ, SK_TUNE_CHILD="1"))
    sys.exit(0)
sim = configs.cfg2(num_packets=packets).setup()
for path in libs:
    lib = abi.load_engine_library(path)
    e = sim.configure(abi.Engine(sim.config_struct(device=0), lib=lib))
    e.prepare_primary(packets)
    ms = []
    for k in range(3):
        e.clear_instruments()
        e.run_segment(0, packets, True, True, False, k)
        ms.append(e.last_kernel_ms())
    stages = {k: round(v, 1) for k, v in e.last_stage_ms().items()}
    c = e.counters()
    if os.environ.get("SK_TUNE_GATHER", "1") == "1" and path == libs[0]:
        print(json.dumps({"gather_peak_records_per_s": e.measure_gather_peak(sim.grid.num_cells), "records": int(sim.grid.num_cells)}), flush=True)
    print(json.dumps({"variant": os.path.basename(path)[4:-3], "packets": packets, "ms": ms,
                      "pkt_per_s": packets / (min(ms[1:]) * 1e-3), "stages_ms": stages, "rounds": c["rounds"] / 3}), flush=True)
    e.close()
