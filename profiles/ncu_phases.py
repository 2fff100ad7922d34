#!/usr/bin/env python3
"""Roll an .ncu-rep source page up into kernel phases: every SASS instruction is
counted once, at the line of the kernel body it was inlined into.
usage: ncu_phases.py report.ncu-rep source.cu  (phases = '// ---- phase X' comments in the kernel)"""
import csv, re, subprocess, sys

def main():
    rep, cu = sys.argv[1], sys.argv[2]
    marks = []
    kstart = None
    for i, l in enumerate(open(cu), 1):
        if "__global__" in l and "filter_kernel" in l:
            kstart = i
            marks.append((i, "setup/load"))
        m = re.search(r"// ---- (phase \w|per-CTA statistics)", l)
        if m and kstart:
            marks.append((i, m.group(1)))
        if kstart and l.startswith("}") and i > kstart and len(marks) > 1 and "kend" not in dict((b, a) for a, b in marks):
            marks.append((i + 1, "kend"))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    infunc = False
    agg = {}
    for r in rows:
        if not r:
            continue
        if r[0] == "Function Name":
            infunc = "filter_kernel" in r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            i_b = hdr.index("stall_barrier")
            continue
        if not infunc or hdr is None or len(r) < len(hdr):
            continue
        try:
            line = int(r[0]); s, ins, th = int(r[i_s]), int(r[i_i]), int(r[i_t]); bar = int(r[i_b] or 0)
        except ValueError:
            continue
        if line < marks[0][0] or line >= marks[-1][0]:
            continue
        name = [n for l0, n in marks if l0 <= line][-1]
        a = agg.setdefault(name, [0, 0, 0, 0])
        a[0] += s; a[1] += ins; a[2] += th; a[3] += bar
    tot = [sum(v[i] for v in agg.values()) for i in range(4)]
    print(f"kernel-body totals: samples {tot[0]} warp-inst {tot[1]} thread-inst {tot[2]}")
    for _, n in marks[:-1]:
        if n in agg:
            s, ins, th, bar = agg[n]
            print(f"{n:22s} samples {100*s/tot[0]:5.1f}% (barrier {100*bar/tot[0]:4.1f}%)  warp-inst {100*ins/tot[1]:5.1f}%  thread-inst {100*th/tot[2]:5.1f}%  act {th/max(1,ins):4.1f}")

main()
