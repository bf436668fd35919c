"""Times the setup kernels on cfg2: octree construction by the density policy + density sampling
on the device, next to the numpy host mirror; and the parts of Simulation.configure() that bench.py's e2e leg pays.
Usage: python scripts/time_setup.py [--host] [--max-level 9]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from skirt9_b200 import abi, configs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--host", action="store_true", help="also time the numpy host mirror (tens of seconds)")
ap.add_argument("--max-level", type=int, default=9)
ap.add_argument("--max-dust-fraction", type=float, default=3.5e-6)
ap.add_argument("--repeat", type=int, default=3)
args = ap.parse_args()

out = {}
sim = configs.cfg2(num_packets=1000, seed=0, max_level=args.max_level, max_dust_fraction=args.max_dust_fraction)
sim.deviceSetup = True
sim.setup()
e = abi.Engine(sim.config_struct())
geom = sim.medium.density_geometry()
pol = sim.grid.tree_policy(sim.numDensitySamples)
e.build_octree(sim.grid.extent, pol, [geom])  # warm-up (context, module load)
tb, ts = [], []
for _ in range(args.repeat):
    t0 = time.perf_counter()
    nn, nc = e.build_octree(sim.grid.extent, pol, [geom])
    t1 = time.perf_counter()
    e.sample_medium(geom, sim.numDensitySamples, nc)
    t2 = time.perf_counter()
    tb.append(t1 - t0)
    ts.append(t2 - t1)
out["device"] = {"nodes": nn, "cells": nc, "build_octree_s": min(tb), "sample_medium_s": min(ts),
                 "num_density_samples": sim.numDensitySamples}
t0 = time.perf_counter()
fc = e.read_octree()
dens, vol = e.read_medium()
out["device"]["read_back_s"] = time.perf_counter() - t0

# configure() parts with host arrays (what e2e pays when the grid comes from the host)
sim.grid.adopt(fc)
sim.density, sim.volume = dens, vol
sim.deviceSetup = False
parts = {}
for _ in range(args.repeat):
    t0 = time.perf_counter()
    e2 = abi.Engine(sim.config_struct())
    t1 = time.perf_counter()
    sim.grid.configure(e2)
    t2 = time.perf_counter()
    e2.set_medium(sim.density, sim.volume)
    t3 = time.perf_counter()
    sim.configure(e2)
    t4 = time.perf_counter()
    e2.prepare_primary(1000)
    e2.run_segment(0, 1000, True, True, False, 1)
    t5 = time.perf_counter()
    for k, v in (("create", t1 - t0), ("set_grid_octree", t2 - t1), ("set_medium", t3 - t2), ("configure_all", t4 - t3),
                 ("first_segment_1000", t5 - t4)):
        parts[k] = min(parts.get(k, 1e9), v)
    e2.close()
out["configure_parts_s"] = parts

if args.host:
    h = configs.cfg2(num_packets=1000, seed=0, max_level=args.max_level, max_dust_fraction=args.max_dust_fraction)
    t0 = time.perf_counter()
    h.setup()
    out["host_numpy"] = {"cells": h.grid.num_cells, "setup_s": time.perf_counter() - t0}
print(json.dumps(out))
