#!/usr/bin/env python
"""bench.py -- photon packets/s of the life-cycle hot path on BASELINE.json configs[1]
(dusty spiral galaxy, ~9.3e5-cell octree, 50 wavelength bins, 256^2 FullInstrument, 1e8 packets per GPU).

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA engine (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W  # the UNMODIFIED reference (oracle/_ref) on the host cores

A step = one primary-emission segment (MonteCarloSimulation::runPrimaryEmission): all histories of the rank's
shard through the life-cycle kernel, then (N>1) the reduction of the instrument arrays over NCCL where the reference
calls ProcessManager::sumToRoot (FluxRecorder.cpp:487-493).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "photon packets/sec"
UNIT = "packets/s"
SKI = os.path.join(ROOT, "tests", "golden", "ski", "cfg2.ski")
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "release", "SKIRT", "main", "skirt")
WORKLOAD = "cfg2: dusty spiral (spiral exp-disk source, ring dust tau_Z=1), OctTree ~9.3e5 cells (levels 3-9), " \
           "50 log wavelength bins 0.1-10 micron, FullInstrument 256x256 i=60deg, forced scattering, no RF"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        for key in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
            if key in d:
                return float(d[key]), "measured (MEASURED_PEAKS.json)"
        if isinstance(d.get("hbm"), dict) and "gbs" in d["hbm"]:
            return float(d["hbm"]["gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k].lower().startswith("active")
                                                         for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def run_reference_once(num_packets, threads, workdir):
    """Times the unmodified reference on the same ski: returns packets/s from its own TimeLogger line."""
    ski = os.path.join(workdir, "cfg2.ski")
    text = open(SKI).read().replace('numPackets="1e6"', f'numPackets="{num_packets:g}"')
    open(ski, "w").write(text)
    out = os.path.join(workdir, "out")
    os.makedirs(out, exist_ok=True)
    subprocess.check_call([REF_EXE, "-t", str(threads), "-b", "-o", out, ski], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    log = open(os.path.join(out, "cfg2_log.txt")).read()
    LAST_REFERENCE_SETUP.clear()
    LAST_REFERENCE_SETUP.update(reference_setup_times(log, threads))
    secs = emission_seconds(log)
    return num_packets / secs, secs


def log_stamp(log, pattern):
    """Seconds since midnight of the first log line matching `pattern` (the reference stamps every line to the ms)."""
    m = re.search(r"\d+/\d+/\d+ (\d+):(\d+):(\d+\.\d+)[ \-!*]+" + pattern, log)
    return None if m is None else 3600 * int(m.group(1)) + 60 * int(m.group(2)) + float(m.group(3))


def emission_seconds(log):
    """Duration of the primary emission segment of a reference (or skirt_b200) run: the TimeLogger pair 'Starting primary
    emission...' / 'Finished primary emission in X s' by their millisecond time stamps; X itself has 0.1 s resolution."""
    t0, t1 = log_stamp(log, "Starting primary emission"), log_stamp(log, "Finished primary emission")
    if t0 is not None and t1 is not None and (t1 - t0) % 86400 > 0:
        return (t1 - t0) % 86400
    return float(re.search(r"Finished primary emission in ([0-9.]+) s", log).group(1))


LAST_REFERENCE_SETUP = {}


def reference_setup_times(log, threads):
    """Grid construction and medium-state sampling times of a reference run, from the millisecond time stamps of its log
    ('Constructing the spatial tree grid...' -> 'Finished construction of the spatial tree grid';
    'Determining medium properties for N cells...' -> 'Done determining medium properties')."""
    t = [log_stamp(log, p) for p in ("Constructing the spatial tree grid", "Finished construction of the spatial tree grid",
                            r"Determining medium properties for \d+ cells", "Done determining medium properties")]
    if any(v is None for v in t):
        return {}
    cells = re.search(r"Determining medium properties for (\d+) cells", log)
    return {"threads": threads, "cells": int(cells.group(1)), "construct_tree_s": (t[1] - t[0]) % 86400,
            "medium_properties_s": (t[3] - t[2]) % 86400}


SHIM_EXE = os.path.join(ROOT, "shim", "_build", "skirt_b200")


def run_shim_once(num_packets, threads):
    """The same cfg2 ski through skirt_b200 (the unmodified reference driven by the C++ shim, life cycle on the GPU):
    packets/s from the reference's own TimeLogger line, exactly as for the CPU reference."""
    with tempfile.TemporaryDirectory() as d:
        ski = os.path.join(d, "cfg2.ski")
        open(ski, "w").write(open(SKI).read().replace('numPackets="1e6"', f'numPackets="{num_packets:g}"'))
        subprocess.check_call([SHIM_EXE, "-t", str(threads), "-b", "-o", d, ski], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
        log = open(os.path.join(d, "cfg2_log.txt")).read()
    secs = emission_seconds(log)
    cells = re.search(r"Determining medium properties for (\d+) cells", log)
    return {"packets": num_packets, "seconds": secs, "packets_per_s": num_packets / secs,
            "cells": int(cells.group(1)) if cells else None,
            "what": "skirt_b200 -t %d cfg2.ski: reference object model + C++ shim + GPU life cycle, time from the "
                    "time stamps of the reference's TimeLogger lines 'Starting / Finished primary emission' (%.3f s)" % (threads, secs)}


def run_port_once(num_packets):
    """Fallback when oracle/_ref is absent: times the single-threaded C port of the oracle."""
    from skirt9_b200 import configs
    from tests.oracle_lib import OracleEngine
    sim = configs.cfg2(num_packets=num_packets).setup()
    e = sim.configure(OracleEngine(sim.config_struct()))
    t = time.time()
    sim.run(e)
    dt = time.time() - t
    return num_packets / dt, dt


def cpu_baseline(sample_packets):
    cores = os.cpu_count() or 1
    if os.path.exists(REF_EXE):
        with tempfile.TemporaryDirectory() as d:
            rate, secs = run_reference_once(sample_packets, cores, d)
        return {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"unmodified SKIRT 9 (oracle/_ref) -t {cores}, same cfg2 ski, {sample_packets:g} packets, "
                          f"log time stamps 'Starting / Finished primary emission': {secs:.3f} s"}
    rate, secs = run_port_once(min(sample_packets, 2e5))
    return {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle/sk_oracle.c single thread, {min(sample_packets, 2e5):g} packets in {secs:.1f} s"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = args.ref_packets
    rates = []
    with tempfile.TemporaryDirectory() as d:
        have_ref = os.path.exists(REF_EXE)
        for it in range(args.warmup + args.steps):
            rate, secs = run_reference_once(n, cores, d) if have_ref else run_port_once(min(n, 2e5))
            if it >= args.warmup:
                rates.append((rate, secs))
    value = sum(r for r, _ in rates) / len(rates)
    ms = 1e3 * sum(s for _, s in rates) / len(rates)
    kind = "reference" if have_ref else "port"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "packets_per_step": n, "timing": "reference TimeLogger 'Finished primary emission'"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores if have_ref else 1, "kind": kind,
                             "sample": f"{n:g} packets per step, {'skirt -t %d' % cores if have_ref else 'C port, 1 thread'}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def device_setup_times(device):
    """SURVEY.md 8f row f2 on the bench workload: DensityTreePolicy::constructTree and the density sampling of
    MediumSystem::setupSelfAfter as CUDA kernels (sk_engine_build_octree / sk_engine_sample_medium), host wall time around
    the blocking C-ABI calls, best of 3 after a warm-up."""
    from skirt9_b200 import abi, configs
    sim = configs.cfg2(num_packets=1000)
    sim.deviceSetup = True
    sim.setup()
    e = abi.Engine(sim.config_struct(device=device))
    geom, pol = sim.medium.density_geometry(), sim.grid.tree_policy(sim.numDensitySamples)
    e.build_octree(sim.grid.extent, pol, [geom])
    best = [1e9, 1e9]
    for _ in range(3):
        t0 = time.perf_counter()
        nn, nc = e.build_octree(sim.grid.extent, pol, [geom])
        t1 = time.perf_counter()
        e.sample_medium(geom, sim.numDensitySamples, nc)
        t2 = time.perf_counter()
        best = [min(best[0], t1 - t0), min(best[1], t2 - t1)]
    e.close()
    return {"nodes": nn, "cells": nc, "num_density_samples": sim.numDensitySamples, "device_build_octree_s": best[0],
            "device_sample_medium_s": best[1]}


def native_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from skirt9_b200 import abi, configs, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    # stdout carries the one JSON line only: anything native libraries print (the NCCL version banner under NCCL_DEBUG)
    # goes to stderr while the bench runs
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    packets_per_gpu = int(args.packets)
    total = packets_per_gpu * world

    sim = configs.cfg2(num_packets=total).setup()
    engine = sim.configure(abi.Engine(sim.config_struct(device=local)))
    stream = torch.cuda.ExternalStream(engine.cuda_stream(), device=local)
    det = engine.device_tensor(3)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=f"cuda:{local}")  # > 126 MB L2
    first, count = parallel.history_block(total, rank, world)   # contiguous block of this rank
    assert count == packets_per_gpu

    def step(stream_id):
        with torch.cuda.stream(stream):
            flush.zero_()                      # L2 flush between iterations
            engine.clear_instruments()
            engine.prepare_primary(total)
            engine.launch_segment(first, packets_per_gpu, True, True, False, stream_id)
            if world > 1:
                dist.all_reduce(det)           # FluxRecorder::calibrateAndWrite -> ProcessManager::sumToRoot

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(args.warmup):
        step(w)
        engine.synchronize()
    engine.counters(reset=True)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, stage_ms = [], []
    ev0.record(stream)
    for k in range(args.steps):
        step(args.warmup + k)
        engine.synchronize()                   # also reads back the engine's own CUDA-event durations
        kernel_ms.append(engine.last_kernel_ms())
        stage_ms.append(engine.last_stage_ms())
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    cnt = engine.counters()

    # ---- end-to-end through the public API with host buffers: tables H2D, run, tallies D2H, every step
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 2))
    h2d = (sim.grid.first_child.nbytes + sim.density.nbytes + sim.volume.nbytes
           + sum(g.borderv.nbytes + g.ellv.nbytes + g.lambdav.nbytes + g.dlambdav.nbytes for g in sim.grids)
           + 4 * sim.medium.mix.lambda_border.nbytes + sum(3 * s.sed.lambdav.nbytes for s in sim.sources))
    d2h = 0
    # host buffers of the caller (allocated and touched once, like the reference's own FluxRecorder arrays)
    nl, npix = sim.defaultWavelengthGrid.num_bins, sim.instruments[0].numPixelsX * sim.instruments[0].numPixelsY
    # (page-locked, as the contract's "pinned host memory"; sk_engine_read_* then skips its own staging buffer)
    ifu_host = [torch.zeros((nl, npix), dtype=torch.float64).pin_memory().numpy() for _ in range(4)]
    e2e_parts = {"configure_s": 0.0, "run_s": 0.0, "read_s": 0.0, "destroy_s": 0.0}
    t0 = time.perf_counter()
    # pass -1 is an untimed warm-up of the whole end-to-end sequence: on first use the device's memory pool grows by a
    # second packet bank next to the one of the timed steps above
    for k in range(-1 if e2e_steps else 0, e2e_steps):
        if k == 0:
            barrier()
            t0 = time.perf_counter()
            e2e_parts = {"configure_s": 0.0, "run_s": 0.0, "read_s": 0.0, "destroy_s": 0.0}
        ta = time.perf_counter()
        e2 = abi.Engine(sim.config_struct(device=local))
        tcreate = time.perf_counter() - ta
        sim.configure(e2)
        tb = time.perf_counter()
        detail = dict(sim.last_configure_parts, create=tcreate)
        for kk, vv in detail.items():
            e2e_parts["configure_" + kk + "_s"] = e2e_parts.get("configure_" + kk + "_s", 0.0) + vv / e2e_steps
        e2.prepare_primary(total)
        e2.run_segment(first, packets_per_gpu, True, True, False, 1000 + k)
        tseg = time.perf_counter()
        e2e_parts["segment_device_s"] = e2e_parts.get("segment_device_s", 0.0) + e2.last_kernel_ms() * 1e-3 / e2e_steps
        e2e_parts["segment_host_s"] = e2e_parts.get("segment_host_s", 0.0) + (tseg - tb) / e2e_steps
        if world > 1:
            with torch.cuda.stream(torch.cuda.ExternalStream(e2.cuda_stream(), device=local)):
                dist.all_reduce(e2.device_tensor(3))
            e2.synchronize()
        tc = time.perf_counter()
        e2e_parts["allreduce_s"] = e2e_parts.get("allreduce_s", 0.0) + (tc - tseg) / e2e_steps
        outs = [e2.read_sed(0, c) for c in (0, 1, 2, 3)] + [e2.read_ifu(0, c, out=ifu_host[c]) for c in (0, 1, 2, 3)]
        d2h = sum(o.nbytes for o in outs)
        td = time.perf_counter()
        e2.close()            # sk_engine_destroy: stream-ordered frees back into the device's memory pool
        tf = time.perf_counter()
        e2e_parts["configure_s"] += (tb - ta) / e2e_steps
        e2e_parts["run_s"] += (tc - tb) / e2e_steps
        e2e_parts["read_s"] += (td - tc) / e2e_steps
        e2e_parts["destroy_s"] += (tf - td) / e2e_steps
    barrier()
    e2e_s = (time.perf_counter() - t0) / max(e2e_steps, 1)
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = total / (ms_per_step * 1e-3)
        pk = max(cnt["packets"], 1)
        S = (cnt["forward_segments"] + cnt["peel_segments"]) / pk
        S_fwd = cnt["forward_segments"] / pk
        P_peel = cnt["peel_paths"] / pk
        bytes_per_packet = 60.0 * S + 32.0 * P_peel + 64.0     # SURVEY.md 8d accounting (no RF store in cfg2)
        kms = sum(kernel_ms) / len(kernel_ms)
        stages = {k: sum(d[k] for d in stage_ms) / len(stage_ms) for k in stage_ms[0]}
        # the dominant kernel: the trace kernel with the largest share of the step; its algorithmic bytes are 60 B per
        # segment it crosses (cell bounds 48 B + density 8 B + link 4 B), its time the sum of its launches
        seg = {"trace_forward": cnt["forward_segments"], "trace_interaction": cnt["replay_segments"],
               "trace_peel": cnt["peel_segments"]}
        dom = max(seg, key=lambda k: stages[k])
        dom_name = {"trace_forward": "sk_wf_trace<2,0,false>", "trace_interaction": "sk_wf_trace<2,1,false>",
                    "trace_peel": "sk_wf_trace<2,2,false>"}[dom]
        dom_bytes = 60.0 * seg[dom] / args.steps
        dom_launches = cnt["rounds"] / args.steps
        achieved = dom_bytes / (stages[dom] * 1e-3) / 1e9
        whole_step = bytes_per_packet * packets_per_gpu / (kms * 1e-3) / 1e9
        peak, which = measured_peak()
        traffic = None
        try:
            # dram__bytes_read.sum + dram__bytes_write.sum of one captured launch of the dominant kernel (ncu --set full,
            # profiles/traffic.json), scaled from that launch's algorithmic bytes to this run's average launch
            cap = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = cap["dram_bytes_per_launch"] * (dom_bytes / max(dom_launches, 1)) / cap["algorithmic_bytes_of_captured_launch"]
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "packets_per_gpu": packets_per_gpu, "cells": int(sim.grid.num_cells),
                           "l2": "256 MB buffer written between iterations", "sharding": "contiguous history blocks, "
                           "replicated grid, NCCL all-reduce of the instrument arrays per step" if world > 1 else "single GPU"},
                "e2e": {"value": total / e2e_s if e2e_steps else None, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "includes": "engine create + octree link build + all table uploads + all stage kernels + read-back of 4 SED and 4 IFU arrays into pinned host buffers + engine destroy; one untimed pass first",
                        "parts": e2e_parts},
                "gpu_launches": int(cnt["kernel_launches"]),
                "kernel": {"name": dom_name, "launches_per_step": dom_launches, "ms_per_step": stages[dom],
                           "ms_per_launch": stages[dom] / max(dom_launches, 1), "share_of_step": stages[dom] / kms,
                           "stage_ms_per_step": stages, "step_ms": kms,
                           "segments_per_packet": S, "forward_segments_per_packet": S_fwd,
                           "replay_segments_per_packet": cnt["replay_segments"] / pk,
                           "peel_paths_per_packet": P_peel, "scatterings_per_packet": cnt["scatterings"] / pk,
                           "segments_per_s": (cnt["forward_segments"] + cnt["peel_segments"] + cnt["replay_segments"])
                           / args.steps / (kms * 1e-3), "tree_fallbacks_per_packet": cnt["fallbacks"] / pk,
                           "rounds_per_step": cnt["rounds"] / args.steps},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": which,
                             "bytes_per_launch": dom_bytes / max(dom_launches, 1),
                             "whole_step_achieved": whole_step, "whole_step_frac": whole_step / peak,
                             "bytes_per_packet": bytes_per_packet,
                             "note": "achieved = 60 B x segments crossed by the dominant trace kernel / its CUDA-event time; "
                                     "whole_step = (60 B/segment + 32 B/detection + 64 B/launch) x packets / step time. The "
                                     "30 MB of cell records are L2-resident, so DRAM traffic is far below the algorithmic "
                                     "bytes: the kernel is latency / fp64-issue bound, not HBM bound"},
                "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_packets)
            line["setup"] = device_setup_times(local)
            if LAST_REFERENCE_SETUP:
                line["setup"]["reference_cpu"] = dict(LAST_REFERENCE_SETUP)
            if os.path.exists(SHIM_EXE) and not args.no_e2e:
                try:
                    line["ski_e2e"] = run_shim_once(4e8, os.cpu_count() or 1)
                except Exception as ex:  # the drop-in binary is informational here; the C-ABI e2e above is the contract
                    line["ski_e2e"] = {"error": str(ex)[:200]}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--packets", type=float, default=1e8, help="packets per GPU per step (BASELINE configs[1]: 1e8)")
    ap.add_argument("--cpu-packets", type=float, default=2e6, help="bounded sample for the cpu_baseline leg")
    ap.add_argument("--ref-packets", type=float, default=1e6, help="packets per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs only)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        native_arm(args)


if __name__ == "__main__":
    main()
