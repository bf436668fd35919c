#!/usr/bin/env python
"""World-size-N NCCL check of the sharded life cycle (run under torchrun, one rank per GPU): a small dust-emission model with
secondary-emission iterations is run (a) sharded over the ranks -- contiguous history blocks, NCCL all-reduce of the
radiation field after every segment and of the detector arrays at the end (skirt9_b200/parallel.py) -- and (b) on rank 0
alone; the random streams are keyed by history index, so both must give the same tallies up to the rounding of the
floating-point sums.  Prints 'NCCL PARITY PASS' on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from skirt9_b200 import abi, parallel
    from tests import models
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = parallel.Comm(dist)
    sim = models.small_dust_emission(num_packets=200000)
    sim.setup()
    e = sim.configure(abi.Engine(sim.config_struct(device=local)))
    sim.run(e, comm=comm)
    sharded_conv = [dict(c) for c in sim.convergence]
    ok = True
    if rank == 0:
        one = sim.configure(abi.Engine(sim.config_struct(device=local)))
        sim.run(one)
        assert len(sim.convergence) == len(sharded_conv)
        for a, b in zip(sharded_conv, sim.convergence):
            for key in ("dust_luminosity", "absorbed_primary", "absorbed_secondary"):
                ok &= abs(a[key] - b[key]) <= 1e-9 * abs(b[key])
        for comp in (abi.SK_COMP_TRANSPARENT, abi.SK_COMP_PRIMARY_DIRECT, abi.SK_COMP_PRIMARY_SCATTERED,
                     abi.SK_COMP_SECONDARY_DIRECT, abi.SK_COMP_SECONDARY_SCATTERED, abi.SK_COMP_TOTAL):
            x, y = e.read_sed(0, comp), one.read_sed(0, comp)
            ok &= bool(np.allclose(x, y, rtol=1e-9, atol=1e-12 * max(y.max(), 1e-300)))
        for which in (0, 1):
            x, y = e.read_rf(which), one.read_rf(which)
            ok &= bool(np.allclose(x, y, rtol=1e-9, atol=1e-12 * y.max()))
        c1 = one.counters()
        print("single rank packets", c1["packets"], "iterations", len(sim.convergence))
    c = e.counters()
    cnt = torch.tensor([c["packets"]], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(cnt)
    if rank == 0:
        ok &= int(cnt.item()) == c1["packets"]     # every history ran exactly once, on one of the ranks
        print("NCCL PARITY PASS" if ok else "NCCL PARITY FAIL", "world", dist.get_world_size())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
