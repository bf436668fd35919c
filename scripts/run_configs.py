#!/usr/bin/env python
"""Runs the multi-GPU workloads of BASELINE.json `configs` at full size and prints one JSON line per workload with timings
and size-independent checks (they are parity cases, not bench lines; bench.py measures configs[1]).

  torchrun --nproc-per-node N scripts/run_configs.py cfg3|cfg4|cfg5 [--packets P]

cfg3: the cfg2 spiral ski with 1e9 packets sharded over the ranks, one all-reduce of the instrument arrays.
cfg4: dust emission with secondary-emission iterations on a ~1e6-cell octree; the radiation field is all-reduced after
      every segment (MediumSystem::communicateRadiationField).
cfg5: Voronoi grid on 5e5 SPH-like particle positions, 1e8 packets.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from skirt9_b200 import abi, configs, parallel
    from skirt9_b200 import host as H

    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["cfg3", "cfg4", "cfg5"])
    ap.add_argument("--packets", type=float, default=None)
    ap.add_argument("--sites", type=int, default=500000)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = parallel.Comm(dist if world > 1 else None)
    t0 = time.perf_counter()
    if args.config == "cfg3":
        n = args.packets or 1e9
        sim = configs.cfg2(num_packets=n)
    elif args.config == "cfg4":
        n = args.packets or 1e7
        sim = configs.cfg4(num_packets=n, max_level=8, max_dust_fraction=3.3e-6)
    else:
        n = args.packets or 1e8
        pc = H.PC
        rng = np.random.default_rng(12345)   # SURVEY.md 8d cfg5 recipe
        R = rng.gamma(2.0, 3000.0, size=2 * args.sites)
        R = R[R < 15000.0][:args.sites]
        phi = rng.uniform(0, 2 * np.pi, size=len(R))
        z = np.clip(rng.laplace(0.0, 250.0, size=len(R)), -1900.0, 1900.0)
        sim = configs.cfg5(np.stack([R * np.cos(phi), R * np.sin(phi), z], axis=1) * pc, num_packets=n, num_pixels=256,
                           record_statistics=False)
    sim.setup()
    t_setup = time.perf_counter() - t0
    e = sim.configure(abi.Engine(sim.config_struct(device=local)))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    sim.run(e, comm=comm)
    e.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_run = time.perf_counter() - t1
    c = e.counters()
    cnt = torch.tensor([c["packets"], c["forward_segments"] + c["peel_segments"], c["scatterings"], c["detections"],
                        c["kernel_launches"]], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(cnt)
    if rank == 0:
        out = {"config": args.config, "n_gpus": world, "cells": int(sim.grid.num_cells), "packets_per_segment": n,
               "host_setup_s": round(t_setup, 2), "run_s": round(t_run, 3), "packets": int(cnt[0].item()),
               "segments": int(cnt[1].item()), "scatterings": int(cnt[2].item()), "detections": int(cnt[3].item()),
               "kernel_launches": int(cnt[4].item()), "packets_per_s": cnt[0].item() / t_run}
        tr = e.read_sed(0, abi.SK_COMP_TRANSPARENT)
        di = e.read_sed(0, abi.SK_COMP_PRIMARY_DIRECT)
        sc = e.read_sed(0, abi.SK_COMP_PRIMARY_SCATTERED)
        out["sed_transparent_sum_W"] = float(tr.sum())
        out["direct_over_transparent"] = float(di.sum() / tr.sum())
        out["scattered_over_transparent"] = float(sc.sum() / tr.sum())
        out["source_luminosity_W"] = float(sum(s.luminosity for s in sim.sources))
        if sim.dustEmissionWLG is not None:
            out["iterations"] = [{k: (v / H.LSUN if k != "iteration" and k != "converged" else v) for k, v in it.items()}
                                 for it in sim.convergence]
            out["dust_luminosity_Lsun"] = sim.dust_luminosity / H.LSUN
            sec = e.read_sed(0, abi.SK_COMP_SECONDARY_DIRECT) + e.read_sed(0, abi.SK_COMP_SECONDARY_SCATTERED)
            out["secondary_sed_sum_W"] = float(sec.sum())
            # energy balance: what the dust absorbs (primary + secondary) is what it emits
            last = sim.convergence[-1]
            out["absorbed_over_emitted"] = (last["absorbed_primary"] + last["absorbed_secondary"]) / sim.dust_luminosity
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
