#!/usr/bin/env python3
"""Static view of libsvjg.so's kernels (no GPU needed): SASS instruction count and opcode histogram per
kernel, and the mnemonics that show the TMA bulk copy + mbarrier staging (UBLKCP / SYNCS) and the dp4a
bit gathers (IDP).      python profiles/sass_summary.py > profiles/rN/sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "svjedi-graph_b200", "lib", "libsvjg.so")],
                     capture_output=True, text=True, check=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    mangled = f.split("\n", 1)[0].strip()
    m = re.search(r"(\d+)([a-z_]+_kernel)", mangled)
    name = m.group(2) if m else mangled
    if "ILb1E" in mangled:
        name += "<exchange>"
    ops = collections.Counter()
    for line in f.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            ops[m.group(1)] += 1
    n = sum(ops.values())
    print(f"{name}: {n} SASS instructions")
    print("   " + ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
    marks = {k: ops[k] for k in ("UBLKCP", "SYNCS", "IDP", "MATCH", "REDUX", "ATOMG", "RED", "ATOM") if ops.get(k)}
    if marks:
        print("   marks: " + ", ".join(f"{k} {v}" for k, v in marks.items()))
