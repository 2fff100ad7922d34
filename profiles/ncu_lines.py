#!/usr/bin/env python3
"""Per-source-line roll-up of an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep [top_n] [samples|inst|thread]
Prints, per CUDA source line of every function in the report: stall samples,
warp instructions, thread instructions, average active threads, and the two
dominant stall reasons.  Lines are ranked by samples."""
import csv
import subprocess
import sys

SORT = 0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    global SORT
    SORT = {"samples": 0, "inst": 1, "thread": 2}[sys.argv[3]] if len(sys.argv) > 3 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    func = ""
    acc = {}
    tot = [0, 0, 0]
    for r in rows:
        if not r:
            continue
        if r[0] == "Function Name":
            func = r[1].split("(")[0][-40:]
            continue
        if r[0] == "Line No":
            hdr = r
            i_s = hdr.index("# Samples")
            i_i = hdr.index("Instructions Executed")
            i_t = hdr.index("Thread Instructions Executed")
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None or len(r) < len(hdr) or r[0] in ("", "File Path", "File Name"):
            continue
        try:
            line = int(r[0])
            s, ins, th = int(r[i_s]), int(r[i_i]), int(r[i_t])
        except ValueError:
            continue
        stalls = {h: int(r[i]) for i, h in stall_cols if r[i].isdigit() and int(r[i])}
        key = (func, line, r[1].strip()[:90])
        a = acc.setdefault(key, [0, 0, 0, {}])
        a[0] += s
        a[1] += ins
        a[2] += th
        for h, v in stalls.items():
            a[3][h] = a[3].get(h, 0) + v
        tot[0] += s
        tot[1] += ins
        tot[2] += th
    print(f"total samples {tot[0]}  warp-inst {tot[1]}  thread-inst {tot[2]}")
    ranked = sorted(acc.items(), key=lambda kv: -kv[1][SORT])[:top]
    for (func, line, src), (s, ins, th, st) in ranked:
        top2 = ",".join(f"{k[6:]}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        act = th / ins if ins else 0
        print(f"{100 * s / max(1, tot[0]):5.1f}% smp {100 * ins / max(1, tot[1]):5.1f}% inst {100 * th / max(1, tot[2]):5.1f}% thr act {act:4.1f} "
              f"L{line:<4d} {top2:28s} | {src}")


if __name__ == "__main__":
    main()
