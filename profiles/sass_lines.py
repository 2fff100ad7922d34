#!/usr/bin/env python3
"""Static code size of a kernel by the source function its SASS instructions come from (inlined code is
attributed to the function whose lines it carries; needs the library built with -lineinfo, no GPU).
    python profiles/sass_lines.py [kernel-substring=scan_parse_kernel] > profiles/rN/sass_<kernel>_by_function.txt"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "svjedi-graph_b200", "csrc", "filter.cu")
kernel = sys.argv[1] if len(sys.argv) > 1 else "scan_parse_kernel"

# function starts in the source: (line, name)
starts = []
pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:__global__|__device__)[^;{]*?\b([A-Za-z_]\w*)\s*\(")
pending_template = False
for no, line in enumerate(open(SRC), 1):
    m = pat.match(re.sub(r"__launch_bounds__\s*\([^)]*\)", "", line))
    if m and not line.rstrip().endswith(";"):
        starts.append((no, m.group(1)))
    elif re.match(r"^\s*(struct|class)\s+(\w+)", line):
        starts.append((no, "struct " + re.match(r"^\s*(struct|class)\s+(\w+)", line).group(2)))


def owner(line_no):
    name = "?"
    for s, n in starts:
        if s <= line_no:
            name = n
        else:
            break
    return name


with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "filter", os.path.join(ROOT, "svjedi-graph_b200", "lib", "libsvjg.so")],
                   cwd=tmp, capture_output=True, check=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True, check=True).stdout

per_fn, per_line = collections.Counter(), collections.Counter()
inside, cur = False, None
for line in text.splitlines():
    if line.startswith(".text."):
        inside = kernel in line
        cur = None
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith("filter.cu") else -1
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        per_fn[owner(cur) if cur and cur > 0 else "(other files / no line)"] += 1
        per_line[cur] += 1
total = sum(per_fn.values())
print(f"{kernel}: {total} SASS instructions = {total * 16 / 1024:.1f} KB of code (16 B each)")
for name, n in per_fn.most_common():
    print(f"  {name:28s} {n:6d}  {100 * n / total:5.1f} %")
print("largest source lines:", ", ".join(f"L{l}:{n}" for l, n in per_line.most_common(12)))
