#!/usr/bin/env python
"""Compares the SASS of the hot kernels of two builds of libskirt9_b200.so opcode by opcode (no GPU needed): shows that the
kernels of the last commit are the ones the ncu captures under profiles/ were taken from, although later commits recompiled them
(a changed parameter structure moves constant-bank loads around).
usage: scripts/sass_diff.py <old.so> <new.so> [label_old label_new]"""
import collections
import re
import subprocess
import sys

HOT = [("_Z11sk_wf_traceILi2ELi0ELb0ELb0ELb0ELb0EE", "sk_wf_trace<2,0> octree forward + interaction"),
       ("_Z11sk_wf_traceILi2ELi2ELb0ELb0ELb0ELb0EE", "sk_wf_trace<2,2> octree peel-off"),
       ("_Z13sk_wf_advanceILi2ELb0EE", "sk_wf_advance<2>"), ("_Z12sk_wf_launchILi2EE", "sk_wf_launch<2>"),
       ("_Z12sk_wf_detectILb0EE", "sk_wf_detect"),
       ("_Z11sk_wf_traceILi2ELi0ELb1ELb0ELb0ELb0EE", "sk_wf_trace<2,0,store> (cfg4)"),
       ("_Z11sk_wf_traceILi1ELi0ELb1ELb1ELb0ELb0EE", "sk_wf_trace<1,0,store,smem> (cfg1)"),
       ("_Z11sk_wf_traceILi3ELi0ELb0ELb0ELb0ELb0EE", "sk_wf_trace<3,0> Voronoi forward (cfg5)"),
       ("_Z11sk_wf_traceILi3ELi2ELb0ELb0ELb0ELb0EE", "sk_wf_trace<3,2> Voronoi peel-off (cfg5)")]


def sass(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    fun, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            fun[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m and cur:
            fun[cur].append(re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip()).split()[0])
    return fun


def main():
    old, new = sys.argv[1], sys.argv[2]
    lo, ln = (sys.argv[3], sys.argv[4]) if len(sys.argv) > 4 else ("old", "new")
    a, b = sass(old), sass(new)
    print("opcode comparison of the hot kernels: %s against %s" % (lo, ln))
    for key, name in HOT:
        x = a[[k for k in a if k.startswith(key)][0]]
        y = b[[k for k in b if k.startswith(key)][0]]
        hx, hy = collections.Counter(x), collections.Counter(y)
        d = {k: (hx[k], hy[k]) for k in sorted(set(hx) | set(hy)) if hx[k] != hy[k]}
        other = {k: v for k, v in d.items() if not (k.startswith("LDC") or k == "NOP")}
        print("%-45s %5d / %5d instructions; differing opcode counts (%s, %s): %s%s"
              % (name, len(x), len(y), lo, ln, d or "none",
                 "" if other else "   [constant-bank loads of kernel parameters and padding only]"))


if __name__ == "__main__":
    main()
