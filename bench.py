#!/usr/bin/env python
"""bench.py -- photon packets/s of the life-cycle hot path on the workloads of BASELINE.json `configs`.

  python bench.py --gpus N --steps K --warmup W                   # this repo's CUDA engine (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W           # the UNMODIFIED reference (oracle/_ref) on the host cores
  python bench.py --config cfg1|cfg2|cfg4|cfg5 ...                 # another workload (default cfg2 = the headline, configs[1])

A step = one pass of the hot path over one batch of histories: for cfg1 / cfg2 / cfg5 one primary-emission segment
(MonteCarloSimulation::runPrimaryEmission), for cfg4 the whole sequence of segments of a dust-emission run (primary
emission, the secondary-emission iterations, the final secondary emission); all histories of the rank's shard go through
the stage kernels, then (N>1) the reductions over NCCL where the reference calls ProcessManager::sumToAll / sumToRoot
(MediumSystem.cpp:1304-1313, FluxRecorder.cpp:487-493).  Prints ONE JSON line on rank 0.
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
SKI_DIR = os.path.join(ROOT, "tests", "golden", "ski")
SKI = os.path.join(SKI_DIR, "cfg2.ski")
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "release", "SKIRT", "main", "skirt")
SHIM_EXE = os.path.join(ROOT, "shim", "_build", "skirt_b200")
PC = 3.0856775814913673e16

WORKLOADS = {
    "cfg2": {
        "text": "cfg2: dusty spiral (spiral exp-disk source, ring dust tau_Z=1), OctTree ~9.3e5 cells (levels 3-9), "
                "50 log wavelength bins 0.1-10 micron, FullInstrument 256x256 i=60deg, forced scattering, no RF",
        "packets": 1e8, "cpu_packets": 2e6, "ref_packets": 1e6, "parity_packets": 1e8, "ski": "cfg2.ski", "store": False,
        "grid": 2},
    "cfg1": {
        "text": "cfg1: point source in a uniform dust sphere (tau=1), Cartesian 32^3 grid, one wavelength 0.55 micron, "
                "FullInstrument 64x64 i=60deg, forced scattering, radiation field stored",
        "packets": 1e7, "cpu_packets": 1e6, "ref_packets": 1e6, "parity_packets": 1e7, "ski": "cfg1.ski", "store": True,
        "grid": 1},
    "cfg4": {
        "text": "cfg4: dust emission with secondary-emission iterations (10000 K point source in an r^-2 dust shell, "
                "tau_Z=20), OctTree ~1e6 cells (levels 3-8), RF grid 40 bins, emission grid 60 bins, SEDInstrument; a step "
                "= primary emission + iterations (convergence 1%/3%, at most 5) + final secondary emission",
        "packets": 1e7, "cpu_packets": 2e5, "ref_packets": 2e5, "parity_packets": 1e7, "ski": "cfg4s.ski", "store": True,
        "grid": 2},
    "cfg5": {
        "text": "cfg5: Voronoi grid on 5e5 SPH-like particle positions (ImportedSites), exp-disk source, one wavelength, "
                "FullInstrument 256x256, forced scattering, no RF",
        "packets": 1e8, "cpu_packets": 2e4, "ref_packets": 2e4, "parity_packets": 0, "ski": None, "store": False,
        "grid": 3},
}


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


# ---------------------------------------------------------------------------------------------------
# the reference (and the drop-in binary) on a ski file: log parsing
# ---------------------------------------------------------------------------------------------------
def log_stamp(log, pattern, last=False):
    """Seconds since midnight of the first (or last) log line matching `pattern` (the reference stamps every line to the ms)."""
    ms = list(re.finditer(r"\d+/\d+/\d+ (\d+):(\d+):(\d+\.\d+)[ \-!*]+" + pattern, log))
    if not ms:
        return None
    m = ms[-1] if last else ms[0]
    return 3600 * int(m.group(1)) + 60 * int(m.group(2)) + float(m.group(3))


def emission_seconds(log):
    """Duration of the primary emission segment of a reference (or skirt_b200) run: the TimeLogger pair 'Starting primary
    emission...' / 'Finished primary emission in X s' by their millisecond time stamps; X itself has 0.1 s resolution."""
    t0, t1 = log_stamp(log, "Starting primary emission"), log_stamp(log, "Finished primary emission")
    if t0 is not None and t1 is not None and (t1 - t0) % 86400 > 0:
        return (t1 - t0) % 86400
    return float(re.search(r"Finished primary emission in ([0-9.]+) s", log).group(1))


def all_segments_seconds(log):
    """From the start of the primary emission to the end of the last emission segment of the run (dust-emission runs: the
    iterations with their per-iteration set-up, and the final secondary emission), and the number of segments."""
    t0 = log_stamp(log, "Starting primary emission")
    t1 = log_stamp(log, r"Finished (?:primary|secondary) emission", last=True)
    nseg = len(re.findall(r"Finished (?:primary|secondary) emission", log))
    return (t1 - t0) % 86400, nseg


def run_phases(log):
    """Wall time of the phases of a reference / skirt_b200 run from its log stamps: setup, the run, final output, total."""
    t = {k: log_stamp(log, p, last=l) for k, p, l in (("start", "Starting simulation", False), ("setup", "Finished setup in", False),
                                                       ("setup_out", "Finished setup output", False),
                                                       ("run0", "Starting primary emission", False),
                                                       ("run1", r"Finished (?:primary|secondary) emission", True),
                                                       ("end", "Finished simulation", True))}
    out = {}
    if t["start"] is not None and t["setup"] is not None:
        out["setup_s"] = (t["setup"] - t["start"]) % 86400
    if t["run0"] is not None and t["run1"] is not None:
        out["emission_s"] = (t["run1"] - t["run0"]) % 86400
    if t["start"] is not None and t["end"] is not None:
        out["total_s"] = (t["end"] - t["start"]) % 86400
    return out


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


def ski_text(name, packets, statistics=False):
    """The workload's ski file with its number of packets (per segment) replaced; cfg4 scales the fixture's octree to the
    ~1e6 cells of BASELINE.json configs[3]."""
    w = WORKLOADS[name]
    text = open(os.path.join(SKI_DIR, w["ski"])).read()
    text = re.sub(r'numPackets="[^"]*"', 'numPackets="%g"' % packets, text, count=1)
    if name == "cfg4":
        text = re.sub(r'maxLevel="\d+"', 'maxLevel="8"', text, count=1)
        text = re.sub(r'maxDustFraction="[^"]*"', 'maxDustFraction="3.3e-6"', text, count=1)
    if statistics:
        text = text.replace('recordStatistics="false"', 'recordStatistics="true"')
    return text


def run_ski(exe, name, packets, threads, workdir, statistics=False, extra=()):
    """Runs `exe` (the reference or skirt_b200) on the workload's ski; returns (log, host wall seconds, output dir, prefix)."""
    prefix = name
    ski = os.path.join(workdir, prefix + ".ski")
    open(ski, "w").write(ski_text(name, packets, statistics))
    out = os.path.join(workdir, "out")
    os.makedirs(out, exist_ok=True)
    t0 = time.perf_counter()
    subprocess.check_call([exe, "-t", str(threads), "-b", "-o", out, *extra, ski], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    wall = time.perf_counter() - t0
    return open(os.path.join(out, prefix + "_log.txt")).read(), wall, out, prefix


def run_reference_once(num_packets, threads, workdir, name="cfg2"):
    """Times the unmodified reference on the workload's ski: returns packets/s from its own TimeLogger lines."""
    log, wall, _, _ = run_ski(REF_EXE, name, num_packets, threads, workdir)
    LAST_REFERENCE_SETUP.clear()
    LAST_REFERENCE_SETUP.update(reference_setup_times(log, threads))
    LAST_REFERENCE_SETUP.update({"phases": run_phases(log), "wall_s": wall})
    if name == "cfg4":
        secs, nseg = all_segments_seconds(log)
        return nseg * num_packets / secs, secs
    secs = emission_seconds(log)
    return num_packets / secs, secs


def run_shim_once(num_packets, threads, name="cfg2"):
    """The same ski through skirt_b200 (the unmodified reference driven by the C++ shim, life cycle on the GPU): packets/s from
    the reference's own TimeLogger lines, exactly as for the CPU reference, and the wall time of the WHOLE run (set-up by the
    reference's host code, engine configuration, emission, output files)."""
    with tempfile.TemporaryDirectory() as d:
        log, wall, _, _ = run_ski(SHIM_EXE, name, num_packets, threads, d)
    if os.environ.get("SK_KEEP_LOG"):  # diagnostic: keep the drop-in's log next to the bench line
        open(os.environ["SK_KEEP_LOG"], "w").write(log)
    gpu_path = "GPU life cycle:" in log and "CPU life cycle (reference)" not in log
    if not gpu_path:
        raise RuntimeError("skirt_b200 did not run the GPU life cycle: " + log[-400:])
    secs = emission_seconds(log)
    cells = re.search(r"Determining medium properties for (\d+) cells", log)
    tree = re.search(r"GPU tree construction: .*", log)
    return {"packets": num_packets, "gpu_path": gpu_path, "seconds": secs, "packets_per_s": num_packets / secs,
            "gpu_tree_construction": tree.group(0) if tree else None,
            "total_wall_s": wall, "phases": run_phases(log), "setup": reference_setup_times(log, threads),
            "cells": int(cells.group(1)) if cells else None,
            "what": "skirt_b200 -t %d %s.ski: reference object model + C++ shim + GPU life cycle; `seconds` from the time "
                    "stamps of the reference's TimeLogger lines 'Starting / Finished primary emission', `total_wall_s` = "
                    "the whole process (ski parsing, set-up, engine configuration, emission, FITS/text output)" % (threads, name)}


def make_sim(name, total_packets, statistics=False):
    """The workload as host mirror objects (setup() done: grid built, densities sampled)."""
    from skirt9_b200 import configs
    if name == "cfg2":
        sim = configs.cfg2(num_packets=total_packets, record_statistics=statistics)
        if os.environ.get("SK_BENCH_SECOND_MIX"):
            # diagnostic (not a BASELINE workload): cfg2 with a second dust component of another material mix, the cost of the
            # several-component trace kernels next to the single-medium ones
            from skirt9_b200 import host as H
            mix2 = H.MeanListDustMix([0.1e-6, 0.55e-6, 10e-6], [1500.0, 1200.0, 400.0], [0.8, 0.7, 0.5], [0.3, 0.2, 0.0])
            sim.medium.tau = 0.6
            sim.extraMedia = [H.GeometricMedium(H.ExpDiskGeometry(5000 * PC, 250 * PC, 0.0, 20000 * PC, 2000 * PC), mix2,
                                                opticalDepth=0.4, wavelength=0.55e-6)]
        if os.environ.get("SK_BENCH_KINEMATICS"):
            # diagnostic (not a BASELINE workload): cfg2 with a rotating dust ring (220 km/s, flat rotation curve) and a source
            # that rotates with it -- the cost of the trace kernels that look the sections up per cell at the perceived wavelength
            from skirt9_b200 import host as H
            sim.medium.velocityMagnitude = 220e3
            sim.medium.velocityDistribution = H.CylindricalVectorField()
            for s in sim.sources:
                s.velocityMagnitude = 220e3
                s.velocityDistribution = H.CylindricalVectorField()
        return sim.setup()
    if name == "cfg1":
        return configs.cfg1(num_packets=total_packets, record_statistics=statistics).setup()
    if name == "cfg4":
        return configs.cfg4(num_packets=total_packets, max_level=8, max_dust_fraction=3.3e-6, record_statistics=statistics).setup()
    import numpy as np
    rng = np.random.default_rng(12345)   # SURVEY.md 8d cfg5 recipe
    sites = int(os.environ.get("SK_BENCH_SITES", "500000"))
    R = rng.gamma(2.0, 3000.0, size=2 * sites)
    R = R[R < 15000.0][:sites]
    phi = rng.uniform(0, 2 * np.pi, size=len(R))
    z = np.clip(rng.laplace(0.0, 250.0, size=len(R)), -1900.0, 1900.0)
    sim = configs.cfg5(np.stack([R * np.cos(phi), R * np.sin(phi), z], axis=1) * PC, num_packets=total_packets,
                       num_pixels=256, record_statistics=False)
    # the tessellation (VoronoiMeshSnapshot::buildMesh) is built by the engine when it is configured: sk_engine_build_voronoi
    sim.deviceSetup = os.environ.get("SK_BENCH_HOST_TESSELLATION") is None
    return sim.setup()


def run_port_once(num_packets, name="cfg2", sim=None):
    """Fallback when oracle/_ref is absent (or the workload has no ski): times the single-threaded C port of the oracle."""
    from tests.oracle_lib import OracleEngine
    if sim is None:
        sim = make_sim(name, num_packets)
    else:
        sim.numPackets = num_packets     # the set-up (grid, densities) does not depend on the number of packets
    e = sim.configure(OracleEngine(sim.config_struct()))
    t = time.time()
    sim.run(e)
    dt = time.time() - t
    return e.counters()["packets"] / dt, dt


def cpu_baseline(sample_packets, name="cfg2", sim=None):
    cores = os.cpu_count() or 1
    w = WORKLOADS[name]
    if os.path.exists(REF_EXE) and w["ski"]:
        with tempfile.TemporaryDirectory() as d:
            rate, secs = run_reference_once(sample_packets, cores, d, name)
        return {"value": rate, "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"unmodified SKIRT 9 (oracle/_ref) -t {cores}, same {name} ski, {sample_packets:g} packets per segment, "
                          f"log time stamps of the emission segments: {secs:.3f} s"}
    n = min(sample_packets, 2e5)
    rate, secs = run_port_once(n, name, sim)
    return {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"oracle/sk_oracle.c single thread, {n:g} packets in {secs:.1f} s"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.config
    w = WORKLOADS[name]
    cores = os.cpu_count() or 1
    n = args.ref_packets or w["ref_packets"]
    rates = []
    have_ref = os.path.exists(REF_EXE) and w["ski"] is not None
    with tempfile.TemporaryDirectory() as d:
        for it in range(args.warmup + args.steps):
            rate, secs = run_reference_once(n, cores, d, name) if have_ref else run_port_once(min(n, 2e5), name)
            if it >= args.warmup:
                rates.append((rate, secs))
    value = sum(r for r, _ in rates) / len(rates)
    ms = 1e3 * sum(s for _, s in rates) / len(rates)
    kind = "reference" if have_ref else "port"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["text"], "packets_per_step": n,
                       "timing": "reference TimeLogger lines of the emission segments (millisecond log stamps)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores if have_ref else 1, "kind": kind,
                             "sample": f"{n:g} packets per segment per step, {'skirt -t %d' % cores if have_ref else 'C port, 1 thread'}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# parity next to the number: the engine's SED against the reference's on the same ski, with both sides' Sum w^k statistics
# ---------------------------------------------------------------------------------------------------
PARITY_PROBES = ('<probeSystem type="ProbeSystem"><ProbeSystem><probes type="Probe">'
                 '<TreeSpatialGridTopologyProbe probeName="topo"/>'
                 '<SpatialCellPropertiesProbe probeName="cells" wavelength="0.55 micron"/>'
                 '</probes></ProbeSystem></probeSystem>')


def adopt_reference_setup(sim, name, outdir):
    """Gives the engine's model the set-up of a reference run (its octree from the TreeSpatialGridTopologyProbe, its sampled
    densities from the SpatialCellPropertiesProbe), the way the golden fixtures do at small size (tests/golden/make_golden.py):
    both sides then see identical inputs and differ by the random streams of the life cycle only.  Returns a description."""
    import numpy as np
    import pandas as pd
    from skirt9_b200 import host as H
    cells = pd.read_csv(os.path.join(outdir, name + "_cells_cellprops.dat"), comment="#", sep=r"\s+", header=None, engine="c").to_numpy()
    topo_path = os.path.join(outdir, name + "_topo_treetop.dat")
    if isinstance(sim.grid, H.PolicyTreeSpatialGrid):
        topo = np.array([int(t) for t in open(topo_path).read().split("\n") if t and not t.startswith("#")], dtype=np.int8)
        ext = sim.grid.extent
        sim.grid = H.FileTreeSpatialGrid(ext[0], ext[3], ext[1], ext[4], ext[2], ext[5], topo, policyOrder=True)
    sim.density = cells[:, 6] * (H.MSUN / H.PC ** 3) / sim.medium.mix.MU
    sim.deviceSetup = False
    sim.setup()
    boxes = sim.grid.cell_boxes()
    centre = 0.5 * (boxes[:, :3] + boxes[:, 3:]) / H.PC
    if len(centre) != len(cells) or np.abs(centre - cells[:, 1:4]).max() > 1e-6 * np.abs(cells[:, 1:4]).max():
        raise RuntimeError("the cells of the imported set-up are not the reference's")
    what = "its octree and sampled densities" if isinstance(sim.grid, H.FileTreeSpatialGrid) else "its sampled densities"
    return "the reference's own set-up (%d cells: %s, imported from its probes)" % (len(cells), what)


def sed_parity(name, device, ref_packets, gpu_packets):
    """Runs the unmodified reference (recordStatistics on) and the engine on the workload and compares the calibrated SED
    columns bin by bin in units of sigma = sqrt(R^2 + R_ref^2) max(F_ref, F_ref_total), R from both sides' Sum w^k
    (4 sigma per bin).  The engine runs on the set-up of the reference run it is compared with (its octree and its
    sampled densities, imported from its probes), so the two differ by the random streams of the life cycle only.  The dust-emission
    columns also carry the noise of the radiation field behind the dust temperatures, which Sum w^k of the last segment does
    not know; so the same statistic is evaluated between two reference runs (seeds 0, 1), and its rms there (1 for a perfect
    error model) scales the 4 sigma bound."""
    import numpy as np
    from skirt9_b200 import abi
    from tests.skirt_files import read_columns
    from tests import mcstats
    cores = os.cpu_count() or 1
    instr = "sed" if name == "cfg4" else "i60"
    refs = []
    sim = make_sim(name, gpu_packets, statistics=True)
    setup_used = "the host mirror's set-up"
    for seed in (0, 1):
        with tempfile.TemporaryDirectory() as d:
            text = ski_text(name, ref_packets, statistics=True).replace('Random seed="0"', 'Random seed="%d"' % seed)
            if "SpatialCellPropertiesProbe" not in text:   # (cfg2: the probes that export the reference's set-up)
                text = text.replace('<probeSystem type="ProbeSystem"><ProbeSystem/></probeSystem>', PARITY_PROBES)
            ski = os.path.join(d, name + ".ski")
            open(ski, "w").write(text)
            subprocess.check_call([REF_EXE, "-t", str(cores), "-b", "-o", d, ski], stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL)
            refs.append((read_columns(os.path.join(d, f"{name}_{instr}_sed.dat")),
                         read_columns(os.path.join(d, f"{name}_{instr}_sedstats.dat"))))
            if seed == 0 and name in ("cfg1", "cfg2", "cfg4") and not os.environ.get("SK_BENCH_SECOND_MIX") \
                    and not os.environ.get("SK_BENCH_KINEMATICS"):
                try:
                    setup_used = adopt_reference_setup(sim, name, d)
                except Exception as err:   # (the comparison then includes the difference between two set-ups)
                    sim = make_sim(name, gpu_packets, statistics=True)
                    setup_used = "the host mirror's set-up (the reference's could not be imported: %s)" % err
    ref, ref_stats = refs[0]
    # N of FluxRecorder.hpp:50-63: the packets launched during the segments that peel off (primary emission, and the final
    # secondary emission of a dust-emission run); w_i = 0 for the histories that do not reach a bin
    peel_segments = 2.0 if sim.dustEmissionWLG is not None else 1.0
    n_ref, n_own = peel_segments * ref_packets, peel_segments * gpu_packets

    def rel_error(stats, launched):
        return mcstats.rel_error(stats, launched)
    e = sim.configure(abi.Engine(sim.config_struct(device=device)))
    sim.run(e, stream_id=7)
    own_stats = e.read_sed_stats(0)
    names = {1: "total", 2: "transparent", 3: "direct", 4: "scattered", 5: "secondary direct", 6: "secondary scattered",
             7: "secondary transparent"}
    cols = ((1, abi.SK_COMP_TOTAL), (2, abi.SK_COMP_TRANSPARENT), (3, abi.SK_COMP_PRIMARY_DIRECT), (4, abi.SK_COMP_PRIMARY_SCATTERED))
    if sim.dustEmissionWLG is not None:
        cols += ((5, abi.SK_COMP_SECONDARY_DIRECT), (6, abi.SK_COMP_SECONDARY_SCATTERED), (7, abi.SK_COMP_SECONDARY_TRANSPARENT))

    def zscores(f, stats, col, launched):
        # (dust emission: only the bins whose error estimate is reliable by the reference's own rule, R < 0.1 and VOV < 0.1 on both
        #  sides -- in the far-UV bins of that workload a handful of heavily weighted packets carry the flux)
        sigma = np.hypot(rel_error(ref_stats[:, 1:].T, n_ref), rel_error(stats, launched))
        scale = np.maximum(ref[:, col], ref[:, 1]) * sigma
        ok = scale > 0
        if sim.dustEmissionWLG is not None:
            ok &= mcstats.reliable(ref_stats[:, 1:].T, launched=n_ref) & mcstats.reliable(stats, launched=launched)
        return np.abs(f - ref[:, col])[ok] / scale[ok]

    own, rr = {}, {}
    for col, comp in cols:
        own[names[col]] = zscores(sim.sed_flux_density(e, 0, comp), own_stats, col, n_own)
        rr[names[col]] = zscores(refs[1][0][:, col], refs[1][1][:, 1:].T, col, n_ref)
    allz, allrr = np.concatenate(list(own.values())), np.concatenate(list(rr.values()))
    rms_rr = float(np.sqrt((allrr ** 2).mean()))
    bound = 4.0 * max(1.0, rms_rr)
    # per column: in a dust-emission run the columns differ by orders of magnitude in how well Sum w^k describes their scatter
    # (the secondary transparent flux has nearly equal weights, so a tiny R, but carries the noise of the radiation field behind
    # the dust temperatures in full), so each column is held to the reference's own run-to-run scatter IN THAT COLUMN
    rms_rr_col = {k: (float(np.sqrt((v ** 2).mean())) if len(v) else 0.0) for k, v in rr.items()}
    bound_col = {k: 4.0 * max(1.0, v) for k, v in rms_rr_col.items()}
    passed = bool(allz.max() <= bound)
    if sim.dustEmissionWLG is not None:
        passed = all((len(v) == 0 or float(v.max()) <= bound_col[k]) for k, v in own.items())
    tot_ref, tot_own = ref[:, 1].sum(), sim.sed_flux_density(e, 0, abi.SK_COMP_TOTAL).sum()
    overflow = e.counters()["pixel_overflows"]
    e.close()
    return {"against": f"unmodified reference, same {name} ski with recordStatistics, -t {cores}, {ref_packets:g} packets; engine "
                       f"{gpu_packets:g} packets on {setup_used}",
            "quantity": "calibrated SED (Jy): " + ", ".join(names[c] for c, _ in cols), "bins": int(len(allz)),
            "criterion": "|F - F_ref| <= 4 max(1, rms_ref_vs_ref) sqrt(R^2 + R_ref^2) max(F_ref, F_ref_total), R = sqrt(Sum w^2 / "
                         "(Sum w)^2 - 1/N) of both sides with N the packets launched in the peel-off segments (FluxRecorder.hpp:50-63); "
                         "rms_ref_vs_ref = rms of the same statistic between two reference runs (seeds 0, 1: these also differ in "
                         "their set-up, which the engine's run on the first one's set-up does not)",
            "max_sigma": float(allz.max()), "rms_sigma": float(np.sqrt((allz ** 2).mean())),
            "bins_over_4_sigma": int((allz > 4).sum()),
            "max_sigma_per_component": {k: float(v.max()) for k, v in own.items()},
            "reference_vs_reference": {"max_sigma": float(allrr.max()), "rms_sigma": rms_rr,
                                       "bins_over_4_sigma": int((allrr > 4).sum())},
            "bound_sigma": bound, "rms_sigma_per_component": {k: (float(np.sqrt((v ** 2).mean())) if len(v) else 0.0) for k, v in own.items()},
            "reference_vs_reference_rms_per_component": rms_rr_col, "bound_sigma_per_component": bound_col,
            "pass_rule": "per component (dust emission)" if sim.dustEmissionWLG is not None else "all bins against bound_sigma",
            "total_flux_ratio": float(tot_own / tot_ref), "pixel_overflows": int(overflow), "pass": passed}


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
    from skirt9_b200 import abi, parallel

    name = args.config
    w = WORKLOADS[name]
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
    comm = parallel.Comm(dist if world > 1 else None)
    packets_per_gpu = int(args.packets or w["packets"])
    total = packets_per_gpu * world     # histories per segment over all ranks (weak scaling)

    t_setup0 = time.perf_counter()
    sim = make_sim(name, total)
    host_setup_s = time.perf_counter() - t_setup0
    engine = sim.configure(abi.Engine(sim.config_struct(device=local)))
    stream = torch.cuda.ExternalStream(engine.cuda_stream(), device=local)
    det = engine.device_tensor(3)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=f"cuda:{local}")  # > 126 MB L2
    # every rank is given the whole segment and runs its interleaved share of it (blocks of 16384 histories)
    first, seg_count = 0, total
    engine.set_history_interleave(parallel.INTERLEAVE_BLOCK, world, rank)
    single_segment = name != "cfg4"
    # a step may consist of several segments: add up the engine's CUDA-event times of all of them, and the wall time of the
    # blocking C-ABI calls between them (dust emission: the per-iteration preparation and convergence test)
    acc = {"kernel_ms": 0.0, "stages": {}, "host_ms": {}}
    timed_calls = ("run_segment", "prepare_primary", "prepare_secondary", "absorbed_luminosity", "communicate_rf", "clear_rf",
                   "clear_instruments")

    def wrap(eng):
        for fn in timed_calls:
            inner = getattr(eng, fn)

            def call(*a, _inner=inner, _fn=fn, **kw):
                t0 = time.perf_counter()
                out = _inner(*a, **kw)
                acc["host_ms"][_fn] = acc["host_ms"].get(_fn, 0.0) + 1e3 * (time.perf_counter() - t0)
                if _fn == "run_segment":
                    acc["kernel_ms"] += eng.last_kernel_ms()
                    for k, v in eng.last_stage_ms().items():
                        acc["stages"][k] = acc["stages"].get(k, 0.0) + v
                return out
            setattr(eng, fn, call)
    if not (single_segment and not w["store"]):
        wrap(engine)
        if world > 1:
            ar = comm._all_reduce

            def timed_all_reduce(eng, which, _ar=ar):
                t0 = time.perf_counter()
                _ar(eng, which)
                acc["host_ms"]["all_reduce_%d" % which] = acc["host_ms"].get("all_reduce_%d" % which, 0.0) + 1e3 * (time.perf_counter() - t0)
            comm._all_reduce = timed_all_reduce

    def step(stream_id):
        with torch.cuda.stream(stream):
            flush.zero_()                      # L2 flush between iterations
        engine.clear_instruments()
        if single_segment and not w["store"]:
            with torch.cuda.stream(stream):
                engine.prepare_primary(total)
                engine.launch_segment(first, seg_count, True, True, False, stream_id)
                if world > 1:
                    dist.all_reduce(det)       # FluxRecorder::calibrateAndWrite -> ProcessManager::sumToRoot
        else:
            # the host mirror of MonteCarloSimulation::runSimulation: every segment, with the reference's reductions
            sim.run(engine, stream_id=stream_id, comm=comm if world > 1 else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        step(k)
        engine.synchronize()
    engine.counters(reset=True)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, stage_ms = [], []
    ev0.record(stream)
    acc["kernel_ms"], acc["stages"], acc["host_ms"] = 0.0, {}, {}
    for k in range(args.steps):
        step(args.warmup + k)
        engine.synchronize()                   # also reads back the engine's own CUDA-event durations
        if single_segment and not w["store"]:
            kernel_ms.append(engine.last_kernel_ms())
            stage_ms.append(engine.last_stage_ms())
    ev1.record(stream)
    if not kernel_ms:
        kernel_ms = [acc["kernel_ms"] / args.steps]
        stage_ms = [{k: acc["stages"].get(k, 0.0) / args.steps for k in engine.STAGES}]
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    cnt = engine.counters()
    cvec = torch.tensor([cnt["packets"]], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(cvec)
    packets_all_ranks = float(cvec[0].item())   # histories launched in the timed steps, all ranks (cfg4: all segments)

    # ---- end-to-end through the public API with host buffers: tables H2D, run, tallies D2H, every step
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 2))
    h2d = sim.density.nbytes + sim.volume.nbytes \
        + sum(g.borderv.nbytes + g.ellv.nbytes + g.lambdav.nbytes + g.dlambdav.nbytes for g in sim.grids) \
        + 4 * sim.medium.mix.lambda_border.nbytes + sum(3 * s.sed.lambdav.nbytes for s in sim.sources)
    if w["grid"] == 2:
        h2d += sim.grid.first_child.nbytes
    elif w["grid"] == 1:
        h2d += sim.grid.xv.nbytes + sim.grid.yv.nbytes + sim.grid.zv.nbytes
    else:
        h2d += sim.grid.sites.nbytes + sim.grid.nbr_offset.nbytes + sim.grid.nbr_index.nbytes
    d2h = 0
    comps = [c for c in (0, 1, 2, 3)]
    ins = sim.instruments[0]
    has_ifu = ins.kind != abi.SK_INSTR_SED
    # host buffers of the caller (allocated and touched once, like the reference's own FluxRecorder arrays); page-locked,
    # as the contract's "pinned host memory": sk_engine_read_* then skips its own staging buffer
    ifu_host = []
    if has_ifu:
        nl = (sim.defaultWavelengthGrid.num_bins if sim.oligoWavelengths is None else len(sim.oligoWavelengths))
        npix = ins.numPixelsX * ins.numPixelsY
        ifu_host = [torch.zeros((nl, npix), dtype=torch.float64).pin_memory().numpy() for _ in comps]
    rf_host = None
    e2e_parts = {"configure_s": 0.0, "run_s": 0.0, "read_s": 0.0, "destroy_s": 0.0}
    e2e_packets = 0
    t0 = time.perf_counter()
    # pass -1 is an untimed warm-up of the whole end-to-end sequence: on first use the device's memory pool grows by a
    # second packet bank next to the one of the timed steps above
    for k in range(-1 if e2e_steps else 0, e2e_steps):
        if k == 0:
            barrier()
            t0 = time.perf_counter()
            e2e_parts = {"configure_s": 0.0, "run_s": 0.0, "read_s": 0.0, "destroy_s": 0.0}
            e2e_packets = 0
        ta = time.perf_counter()
        e2 = abi.Engine(sim.config_struct(device=local))
        sim.configure(e2)
        e2.set_history_interleave(parallel.INTERLEAVE_BLOCK, world, rank)
        tb = time.perf_counter()
        if single_segment and not w["store"]:
            e2.prepare_primary(total)
            e2.run_segment(first, seg_count, True, True, False, 1000 + k)
            if world > 1:
                with torch.cuda.stream(torch.cuda.ExternalStream(e2.cuda_stream(), device=local)):
                    dist.all_reduce(e2.device_tensor(3))
                e2.synchronize()
        else:
            sim.run(e2, stream_id=1000 + k, comm=comm if world > 1 else None)
        tc = time.perf_counter()
        e2e_packets += e2.counters()["packets"]
        outs = [e2.read_sed(0, c) for c in comps]
        if has_ifu:
            outs += [e2.read_ifu(0, c, out=ifu_host[c]) for c in comps]
        if w["store"]:
            if rf_host is None:   # (allocated in the untimed pass: the caller's table, like MediumSystem::_rf1)
                rf_host = torch.zeros((e2.num_cells, e2.num_rf), dtype=torch.float64).pin_memory().numpy()
            outs.append(e2.read_rf(0, out=rf_host))
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
    te = torch.tensor([e2e_s, float(e2e_packets)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        tmax = te.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(te)
        te[0] = tmax[0]
    e2e_s, e2e_packets_all = float(te[0].item()), float(te[1].item())

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = packets_all_ranks / args.steps / (ms_per_step * 1e-3)
        pk = max(cnt["packets"], 1)
        S = (cnt["forward_segments"] + cnt["peel_segments"]) / pk
        S_fwd = cnt["forward_segments"] / pk
        P_peel = cnt["peel_paths"] / pk
        # SURVEY.md 8d accounting: per segment 60 B (tree, Cartesian) or 32 B + 28 B per neighbour (Voronoi), + 16 B per RF
        # deposit, + 32 B per detection, + 64 B per launch
        if w["grid"] == 3:
            nbrs = float(sim.grid.nbr_offset[-1]) / sim.grid.num_cells
            seg_bytes = 32.0 + 28.0 * nbrs
        else:
            seg_bytes = 60.0
        dep = cnt["rf_deposits"] / pk
        bytes_per_packet = seg_bytes * S + 16.0 * dep + 32.0 * P_peel + 64.0
        kms = sum(kernel_ms) / len(kernel_ms)
        stages = {k: sum(d[k] for d in stage_ms) / len(stage_ms) for k in stage_ms[0]}
        # the dominant kernel: the trace kernel with the largest share of the step.  With forced scattering the forward
        # kernel also walks to the interaction point (replay segments: the same cell records a second time)
        seg = {"trace_forward": cnt["forward_segments"] + (cnt["replay_segments"] if stages["trace_interaction"] == 0 else 0),
               "trace_interaction": cnt["replay_segments"] if stages["trace_interaction"] > 0 else 0,
               "trace_peel": cnt["peel_segments"]}
        dom = max(seg, key=lambda k: stages[k])
        g = w["grid"]
        store_flag = "true" if w["store"] else "false"
        dom_name = {"trace_forward": f"sk_wf_trace<{g},0,{store_flag}> (forward path + walk to the interaction point)",
                    "trace_interaction": f"sk_wf_trace<{g},1,false>", "trace_peel": f"sk_wf_trace<{g},2,false>"}[dom]
        dom_bytes = (seg_bytes * seg[dom] + (16.0 * cnt["rf_deposits"] if dom == "trace_forward" else 0.0)) / args.steps
        dom_launches = cnt["rounds"] / args.steps
        achieved = dom_bytes / (stages[dom] * 1e-3) / 1e9
        whole_step = bytes_per_packet * pk / args.steps / (kms * 1e-3) / 1e9
        peak, which = measured_peak()
        traffic = None
        try:
            # dram__bytes_read.sum + dram__bytes_write.sum of one captured launch of the dominant kernel of THIS build (ncu
            # --set full, profiles/traffic.json), scaled from that launch's algorithmic bytes to this run's average launch
            cap = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if cap.get("workload", "cfg2") == name and cap.get("kernel", dom) == dom:
                traffic = cap["dram_bytes_per_launch"] * (dom_bytes / max(dom_launches, 1)) / cap["algorithmic_bytes_of_captured_launch"]
        except Exception:
            pass
        # the second roofline, against the resource that binds: dependent, scattered 32-byte record fetches.  Measured
        # live on this device with the crossing loop's access pattern and nothing else (sk_engine_measure_gather_peak)
        binding = None
        if w["grid"] != 1:
            try:
                gp = engine.measure_gather_peak(max(int(sim.grid.num_cells), 1024))
                per_seg_fetches = 1.0 if w["grid"] == 2 else 1.0 + nbrs
                ach = seg[dom] * per_seg_fetches / args.steps / (stages[dom] * 1e-3)
                binding = {"bound": "dependent scattered 32-byte record fetches (L2-resident table, one 256-bit load per "
                                    "fetch, full occupancy, no arithmetic)", "achieved": ach, "peak": gp,
                           "unit": "record fetches/s", "frac": ach / gp,
                           "note": "achieved = cell records the dominant trace kernel fetched / its CUDA-event time; peak = "
                                   "sk_engine_measure_gather_peak on a table of the grid's size, measured in this run"}
            except Exception as ex:
                binding = {"error": str(ex)[:200]}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": w["text"], "name": name, "packets_per_gpu": packets_per_gpu, "cells": int(sim.grid.num_cells),
                           "l2": "256 MB buffer written between iterations", "sharding": "interleaved blocks of 16384 histories, "
                           "replicated grid, NCCL all-reduce of the instrument arrays per step" + (" and of the radiation "
                           "field per segment" if w["store"] else "") if world > 1 else "single GPU"},
                "e2e": {"value": e2e_packets_all / max(e2e_steps, 1) / e2e_s if e2e_steps else None, "unit": UNIT,
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "includes": "engine create + grid / link build + all table uploads + all stage kernels + read-back of the "
                                    "SED and IFU arrays (and the radiation field when stored) into pinned host buffers + engine "
                                    "destroy; one untimed pass first", "parts": e2e_parts},
                "gpu_launches": int(cnt["kernel_launches"]),
                "kernel": {"name": dom_name, "launches_per_step": dom_launches, "ms_per_step": stages[dom],
                           "ms_per_launch": stages[dom] / max(dom_launches, 1), "share_of_step": stages[dom] / kms,
                           "stage_ms_per_step": stages, "step_ms": kms,
                           "segments_per_packet": S, "forward_segments_per_packet": S_fwd,
                           "replay_segments_per_packet": cnt["replay_segments"] / pk,
                           "rf_deposits_per_packet": dep,
                           "peel_paths_per_packet": P_peel, "scatterings_per_packet": cnt["scatterings"] / pk,
                           "segments_per_s": (cnt["forward_segments"] + cnt["peel_segments"] + cnt["replay_segments"])
                           / args.steps / (kms * 1e-3), "tree_fallbacks_per_packet": cnt["fallbacks"] / pk,
                           "rounds_per_step": cnt["rounds"] / args.steps, "packets_per_step_this_rank": pk / args.steps},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": which,
                             "bytes_per_launch": dom_bytes / max(dom_launches, 1),
                             "whole_step_achieved": whole_step, "whole_step_frac": whole_step / peak,
                             "bytes_per_packet": bytes_per_packet,
                             "note": "achieved = algorithmic bytes (SURVEY.md 8d: %g B per segment, +16 B per RF deposit) of the "
                                     "segments the dominant trace kernel crossed / its CUDA-event time.  The cell records are "
                                     "L2-resident, so DRAM traffic is far below the algorithmic bytes and this fraction can "
                                     "exceed what HBM alone could deliver: the kernel is bound by dependent scattered record "
                                     "fetches, see roofline_binding" % seg_bytes},
                "roofline_binding": binding,
                "host_setup_s": host_setup_s,
                "clocks": clocks}
        if acc["host_ms"]:
            line["host_calls_ms_per_step"] = {k: v / args.steps for k, v in sorted(acc["host_ms"].items())}
        if name == "cfg4":
            line["iterations"] = len(sim.convergence)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_packets or w["cpu_packets"], name, sim)
            if LAST_REFERENCE_SETUP:
                line["cpu_baseline"]["whole_run"] = dict(LAST_REFERENCE_SETUP)
            if name == "cfg2":
                line["setup"] = device_setup_times(local)
                if LAST_REFERENCE_SETUP:
                    line["setup"]["reference_cpu"] = {k: v for k, v in LAST_REFERENCE_SETUP.items() if k not in ("phases", "wall_s")}
            if os.path.exists(REF_EXE) and w["ski"] and w["parity_packets"] and name in ("cfg1", "cfg2", "cfg4") and not args.no_parity:
                try:
                    # (cfg4: five times the packets of the timing sample, the dust-emission columns need the statistics)
                    line["parity"] = sed_parity(name, local, (args.cpu_packets or w["cpu_packets"]) * (5 if name == "cfg4" else 1),
                                                int(w["parity_packets"]))
                except Exception as ex:
                    line["parity"] = {"error": str(ex)[:300], "pass": False}
            if os.path.exists(SHIM_EXE) and w["ski"] and not args.no_e2e and name in ("cfg1", "cfg2"):
                try:
                    # the workload's own number of packets: what a user of the drop-in binary waits for, start to end
                    line["ski_e2e"] = run_shim_once(packets_per_gpu, os.cpu_count() or 1, name)
                except Exception as ex:  # the drop-in binary is informational here; the C-ABI e2e above is the contract
                    line["ski_e2e"] = {"error": str(ex)[:300], "gpu_path": False}
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
    ap.add_argument("--config", default="cfg2", choices=sorted(WORKLOADS), help="BASELINE.json workload (default: configs[1])")
    ap.add_argument("--packets", type=float, default=None, help="packets per GPU per segment (cfg2: 1e8 = BASELINE configs[1])")
    ap.add_argument("--cpu-packets", type=float, default=None, help="bounded sample for the cpu_baseline leg")
    ap.add_argument("--ref-packets", type=float, default=None, help="packets per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the SED comparison with the reference")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs only)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        native_arm(args)


if __name__ == "__main__":
    main()
