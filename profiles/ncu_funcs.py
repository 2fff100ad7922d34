#!/usr/bin/env python3
"""Roll an .ncu-rep's source page (needs -lineinfo and --import-source on) up by the source FUNCTION a
line belongs to; the scan kernel's own lines are split at its phase comments ("---- phase A", ...).
usage: ncu_funcs.py report.ncu-rep [source=svjedi-graph_b200/csrc/filter.cu]"""
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
SRC = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "svjedi-graph_b200", "csrc", "filter.cu")

starts = []
pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__global__|__device__)[^;{]*?\b([A-Za-z_]\w*)\s*\(")
in_scan = False
for no, line in enumerate(open(SRC), 1):
    m = pat.match(re.sub(r"__launch_bounds__\s*\([^)]*\)", "", line))
    if m and not line.rstrip().endswith(";"):
        starts.append((no, m.group(1)))
        in_scan = m.group(1) == "scan_kernel"
    elif re.match(r"^\s*(struct|class)\s+(\w+)", line):
        starts.append((no, "struct " + re.match(r"^\s*(struct|class)\s+(\w+)", line).group(2)))
    elif in_scan:
        m = re.match(r"\s*// ---- (phases? [A-D](?: and [A-D])?|lines with more nodes)", line)
        if m:
            starts.append((no, "scan_kernel:" + m.group(1)))
starts.sort()


def owner(n):
    name = "?"
    for s, f in starts:
        if s > n:
            break
        name = f
    return name


out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
hdr = None
acc = {}
tot = [0, 0, 0]
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        line, s, ins, th = int(r[0]), int(r[i_s]), int(r[i_i]), int(r[i_t])
    except ValueError:
        continue
    src = r[1]
    # lines of header files (shuffles, votes, atomics) carry their own numbering: keep them apart
    key = owner(line) if ("nvvm" not in src and "Atomic" not in src) else "(intrinsics headers)"
    a = acc.setdefault(key, [0, 0, 0])
    a[0] += s
    a[1] += ins
    a[2] += th
    tot[0] += s
    tot[1] += ins
    tot[2] += th
print(f"total: samples {tot[0]}  warp-inst {tot[1]}  thread-inst {tot[2]}")
for k, (s, ins, th) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    if ins == 0 and s == 0:
        continue
    print(f"{k:34s} {100 * s / max(1, tot[0]):5.1f}% samples {100 * ins / max(1, tot[1]):5.1f}% warp-inst ({ins / 1e6:7.1f} M)  active lanes {th / max(1, ins):4.1f}")
