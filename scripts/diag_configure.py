"""Diagnostic: what makes Simulation.configure() slow inside bench.py's process?  Times set_grid_octree / set_medium under
growing amounts of process state."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skirt9_b200 import abi, configs  # noqa: E402

sim = configs.cfg2(num_packets=1000, seed=0)
sim.deviceSetup = True
sim.setup()
e0 = abi.Engine(sim.config_struct())
sim.configure(e0)
sim.fetch_device_setup(e0)
sim.deviceSetup = False
keep = []


def measure(tag, close=True):
    e = abi.Engine(sim.config_struct())
    t0 = time.perf_counter()
    sim.grid.configure(e)
    t1 = time.perf_counter()
    e.set_medium(sim.density, sim.volume)
    t2 = time.perf_counter()
    print(f"{tag:40s} grid {1e3 * (t1 - t0):7.1f} ms   medium {1e3 * (t2 - t1):6.1f} ms", flush=True)
    if close:
        e.close()
    else:
        keep.append(e)


measure("plain")
measure("plain again")
measure("plain, engine kept alive", close=False)
measure("after keeping one alive")
e0.prepare_primary(20_000_000)
e0.run_segment(0, 20_000_000, True, True, False, 1)
os.environ["SK_DEBUG_TIMING"] = "1"
measure("after a 2e7-packet segment (bank 2.3 GB)")
del os.environ["SK_DEBUG_TIMING"]
import torch  # noqa: E402
torch.cuda.set_device(0)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda:0")
flush.zero_()
torch.cuda.synchronize()
measure("after torch context + 256 MB tensor")
s = torch.cuda.ExternalStream(e0.cuda_stream(), device=0)
with torch.cuda.stream(s):
    flush.zero_()
torch.cuda.synchronize()
measure("after torch ExternalStream use")
pinned = torch.empty(105_000_000 // 8, dtype=torch.float64).pin_memory()
measure("after pinning 105 MB host memory")
import numpy as np  # noqa: E402
out = e0.read_ifu(0, 0)
measure("after read_ifu (pinned staging in engine)")
