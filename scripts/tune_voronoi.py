#!/usr/bin/env python
"""Times the Voronoi traversal (cfg5-like) for every variants/lib_*.so: usage tune_voronoi.py [sites] [packets]"""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
sites = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200000
packets = int(float(sys.argv[2])) if len(sys.argv) > 2 else 20000000
if os.environ.get("SK_TUNE_CHILD") != "1":
    for path in sorted(glob.glob(os.path.join(ROOT, "variants", "lib_*.so"))):
        subprocess.call([sys.executable, __file__, str(sites), str(packets)], env=dict(os.environ, SK_TUNE_CHILD="1", SK_ENGINE_LIB=path))
    sys.exit(0)
from skirt9_b200 import abi, configs, host as H
pc = H.PC
rng = np.random.default_rng(12345)
R = rng.gamma(2.0, 3000.0, size=2 * sites); R = R[R < 15000.0][:sites]
phi = rng.uniform(0, 2 * np.pi, size=len(R)); z = np.clip(rng.laplace(0.0, 250.0, size=len(R)), -1900.0, 1900.0)
sim = configs.cfg5(np.stack([R * np.cos(phi), R * np.sin(phi), z], axis=1) * pc, num_packets=packets, num_pixels=256, record_statistics=False).setup()
e = sim.configure(abi.Engine(sim.config_struct(device=0)))
e.prepare_primary(packets)
ms = []
for k in range(3):
    e.clear_instruments(); e.run_segment(0, packets, True, True, False, k); ms.append(e.last_kernel_ms())
c = e.counters()
print(json.dumps({"lib": os.path.basename(os.environ.get("SK_ENGINE_LIB", "default")), "cells": sim.grid.num_cells, "packets": packets, "ms": ms,
                  "pkt_per_s": packets / (min(ms[1:]) * 1e-3), "segments_per_packet": (c["forward_segments"] + c["peel_segments"]) / c["packets"],
                  "stages_ms": {k: round(v, 1) for k, v in e.last_stage_ms().items()}}), flush=True)
