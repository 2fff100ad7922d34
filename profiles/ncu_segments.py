#!/usr/bin/env python3
"""Split the SASS of the main kernel of an .ncu-rep at its BAR.SYNC instructions and
report, per segment (= phase between two barriers, in address order) and per
other function: warp instructions, thread instructions, samples, top stall.
usage: ncu_segments.py report.ncu-rep [kernel-substring]"""
import csv, subprocess, sys

def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else "filter_kernel"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None; func = ""; segs = []; cur = None
    for r in rows:
        if not r: continue
        if r[0] in ("Function Name", "Kernel Name"):
            func = r[1]; cur = [func.split("(")[0][-48:], 0, 0, 0, {}, 0]; segs.append(cur); continue
        if r[0] == "Address":
            hdr = r
            i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
            st = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr) or cur is None: continue
        try:
            s, ins, th = int(r[i_s]), int(r[i_i]), int(r[i_t])
        except ValueError:
            continue
        cur[1] += ins; cur[2] += th; cur[3] += s; cur[5] += 1
        for i, h in st:
            if r[i].isdigit() and int(r[i]): cur[4][h] = cur[4].get(h, 0) + int(r[i])
        if pat in func and ("BAR.SYNC" in r[1] or " EXIT" in r[1] or " RET." in r[1]):
            cur = [f"  .. after barrier/exit/ret #{sum(1 for x in segs if x[0].startswith('  ..')) + 1}", 0, 0, 0, {}, 0]; segs.append(cur)
    ti = sum(x[1] for x in segs); tt = sum(x[2] for x in segs); ts = sum(x[3] for x in segs)
    print(f"totals: warp-inst {ti}  thread-inst {tt}  samples {ts}")
    for name, ins, th, s, stl, n in segs:
        if ins == 0 and s == 0: continue
        top = ",".join(f"{k[6:]}:{100*v/max(1,ts):.1f}%" for k, v in sorted(stl.items(), key=lambda kv: -kv[1])[:3])
        print(f"{name:50s} sass {n:5d} warp-inst {100*ins/ti:5.1f}% thread-inst {100*th/tt:5.1f}% act {th/max(1,ins):4.1f} samples {100*s/ts:5.1f}%  {top}")
main()
