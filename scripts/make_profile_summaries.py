#!/usr/bin/env python
"""Turns the ncu artefacts of a gpurun (gpurun_out/<tag>_launches.csv, <tag>_prof.ncu-rep, optional <tag>_events.ncu-rep /
<tag>_launch.ncu-rep) into the small tracked summaries under profiles/: the launch list, per-kernel shares of the step,
and one JSON with the key metrics of every fully captured kernel.  usage: scripts/make_profile_summaries.py <tag> [reps...]"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# launch list -> shares
src = os.path.join(G, tag + "_launches.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, tag + "_launches.csv"))
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot = {}
    for r in rows:
        if r is hdr or len(r) <= mv or not r[mv].replace(".", "").replace(",", "").isdigit():
            continue
        name = r[kn].split("(")[0]
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += float(r[mv].replace(",", ""))
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(P, tag + "_launch_shares.csv"), "w") as f:
        f.write("kernel,launches,total_ns,share\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{v[0]},{v[1]:.0f},{v[1] / total:.4f}\n')

# full captures -> metrics
want = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
        "launch__registers_per_thread": "registers", "launch__grid_size": "grid",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_lsu_wavefronts_pct",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio": "avg_active_threads", "smsp__inst_executed.sum": "inst_executed"}
out = []
reps = sys.argv[2:] or [tag + "_prof", tag + "_events", tag + "_launch"]
for rep in reps:
    path = os.path.join(G, rep + ".ncu-rep")
    raw = os.path.join(G, rep + "_raw.csv")   # the raw page exported on the GPU box (scripts/final_profile.sh)
    if os.path.exists(raw):
        txt = open(raw).read()
    elif os.path.exists(path):
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        continue
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        d = {"report": rep, "kernel": r[hdr.index("Kernel Name")]}
        for k, name in want.items():
            if k in hdr:
                try:
                    d[name] = float(r[hdr.index(k)].replace(",", "")) * scale.get(units[hdr.index(k)], 1.0)
                except ValueError:
                    pass
        stalls = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(r[i]))
                  for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and "not_issued" not in h
                  and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        d["top_stalls_per_issue"] = dict(sorted(stalls, key=lambda x: -x[1])[:5])
        out.append(d)
json.dump(out, open(os.path.join(P, tag + "_ncu_full_kernels.json"), "w"), indent=1)
print(len(out), "kernels ->", os.path.join(P, tag + "_ncu_full_kernels.json"))
