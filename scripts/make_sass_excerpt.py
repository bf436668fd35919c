#!/usr/bin/env python
"""SASS evidence for the trace kernels of the built library: for each kernel the crossing loop (the innermost loop around the
32-byte cell-record load, LDG.E.ENL2.256) is printed in full together with its instruction mix, and the whole-kernel counts
of the instructions that characterise the design (256-bit record loads, fp64 RED deposits, MATCH/REDUX warp aggregation,
UBLKCP bulk copies into shared memory).

usage: python scripts/make_sass_excerpt.py <tag> <commit>     ->  profiles/<tag>_sass_crossing_loop.txt
"""
import collections
import re
import subprocess
import sys

LIB = "skirt9_b200/csrc/libskirt9_b200.so"
KERNELS = [
    ("_Z11sk_wf_traceILi2ELi2ELb0ELb0ELb0ELb0EE", "sk_wf_trace<2,2,0,0,0>: octree, peel-off paths (observer direction shared by the launch)"),
    ("_Z11sk_wf_traceILi2ELi0ELb0ELb0ELb0ELb0EE", "sk_wf_trace<2,0,0,0,0>: octree, fused forward path + walk to the interaction point"),
    ("_Z11sk_wf_traceILi2ELi0ELb1ELb0ELb0ELb0EE", "sk_wf_trace<2,0,1,0,0>: octree, fused, radiation field stored (cfg4)"),
    ("_Z11sk_wf_traceILi1ELi0ELb1ELb1ELb0ELb0EE", "sk_wf_trace<1,0,1,1,0>: Cartesian, fused, radiation field stored, mesh tables in shared memory (cfg1)"),
    ("_Z11sk_wf_traceILi3ELi0ELb0ELb0ELb0ELb0EE", "sk_wf_trace<3,0,0,0,0>: Voronoi, fused (cfg5)"),
    ("_Z11sk_wf_traceILi2ELi0ELb0ELb0ELb1ELb0EE", "sk_wf_trace<2,0,0,0,1>: octree, fused, several medium components (MULTI)"),
]
INS = re.compile(r"^\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);")


def parse(lines):
    out = []
    for ln in lines:
        m = INS.match(ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(text):
    t = text.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0]


def main():
    tag, commit = sys.argv[1], sys.argv[2]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout.splitlines()
    starts = [i for i, l in enumerate(sass) if "Function :" in l] + [len(sass)]
    out = [f"# SASS of {LIB} (sm_100a), commit {commit}: crossing loops of the trace kernels",
           "# made by scripts/make_sass_excerpt.py; the loop shown is the innermost backward branch around the record load", ""]
    for key, title in KERNELS:
        idx = [k for k, i in enumerate(starts[:-1]) if key in sass[i]]
        if not idx:
            out.append(f"## {title}: not in the library\n")
            continue
        body = parse(sass[starts[idx[0]]:starts[idx[0] + 1]])
        mix = collections.Counter(opcode(t) for _, t in body)
        marks = {k: sum(v for o, v in mix.items() if o.startswith(k))
                 for k in ("LDG.E.ENL2.256", "LDG", "STG", "RED", "ATOM", "MATCH", "REDUX", "UBLKCP", "LDS", "DFMA", "DMUL",
                           "DADD", "DSETP", "MUFU", "LDL", "STL")}
        out.append(f"## {title}")
        out.append(f"whole kernel: {len(body)} instructions; " + ", ".join(f"{k} {v}" for k, v in marks.items() if v))
        loads = [a for a, t in body if "LDG.E.ENL2.256" in t] or [a for a, t in body if "LDG.E.128" in t]
        # innermost loop (smallest backward branch range) that contains the first record load
        best = None
        for a, t in body:
            m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?(0x[0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a and loads and any(tgt <= l <= a for l in loads):
                    if best is None or a - tgt < best[1] - best[0]:
                        best = (tgt, a)
        if best is None:
            out.append("(no loop around a record load found)\n")
            continue
        loop = [(a, t) for a, t in body if best[0] <= a <= best[1]]
        lm = collections.Counter(opcode(t).split(".")[0] for _, t in loop)
        out.append(f"crossing loop {best[0]:#06x}..{best[1]:#06x}: {len(loop)} instructions; mix: "
                   + ", ".join(f"{k} {v}" for k, v in lm.most_common()))
        out.append("```")
        out += [f"/*{a:04x}*/  {t}" for a, t in loop]
        out.append("```\n")
    path = f"profiles/{tag}_sass_crossing_loop.txt"
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", path, len(out), "lines")


if __name__ == "__main__":
    main()
