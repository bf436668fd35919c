#!/usr/bin/env python
"""profiles/traffic.json from the full ncu capture of the dominant trace kernel (gpurun_out/<tag>_prof.ncu-rep) and the bench
line of the same build (gpurun_out/<tag>_bench_cfg2.json): DRAM bytes of the captured launch next to its algorithmic bytes.
usage: scripts/make_traffic.py <tag> <commit>"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, commit = sys.argv[1], sys.argv[2]
rep = os.path.join(ROOT, "gpurun_out", tag + "_prof.ncu-rep")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
bench = json.load(open(os.path.join(ROOT, "gpurun_out", tag + "_bench_cfg2.json")))
k = bench["kernel"]
paths = k["peel_paths_per_packet"]          # one forward path per peel-off path (emission + every scattering)
segs_per_ray = (k["forward_segments_per_packet"] + k["replay_segments_per_packet"]) / paths
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if "sk_wf_trace<2, 0" not in name:
        continue
    def val(m):
        i = hdr.index(m)
        return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
    rays = 1 << 23
    out = {"workload": "cfg2", "kernel": "trace_forward", "kernel_name": name, "commit": commit,
           "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
           "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
           "duration_ms": val("gpu__time_duration.sum"),
           "algorithmic_bytes_of_captured_launch": 60.0 * segs_per_ray * rays,
           "captured": "ncu --set full --clock-control none --import-source on, bench.py --packets 2e7 --steps 1 --warmup 1 "
                       "(scripts/final_profile.sh %s): a launch that walks the forward rays of a full bank (2^23 rays, %.1f "
                       "segments each including the walk to the interaction point)" % (tag, segs_per_ray),
           "note": "DRAM traffic is a few % of the algorithmic bytes: the 30 MB of cell records stay in L2; most of the DRAM bytes "
                   "are the packet bank streaming through.  bench.py scales dram_bytes_per_launch to its average launch (fewer "
                   "rays than a full bank) for roofline.traffic"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
    break
