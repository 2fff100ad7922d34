#!/usr/bin/env python3
"""Sum the per-source-line counters of an .ncu-rep over named line ranges of one file.
usage: ncu_regions.py report.ncu-rep source.cu name:lo-hi [name:lo-hi ...]
Lines outside every range are reported as 'other' (with the top few listed)."""
import csv
import subprocess
import sys


def main():
    rep, src = sys.argv[1], sys.argv[2]
    regions = []
    for spec in sys.argv[3:]:
        name, r = spec.split(":")
        lo, hi = r.split("-")
        regions.append((name, int(lo), int(hi)))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    per = {}
    for r in rows:
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
        a = per.setdefault(line, [0, 0, 0])
        a[0] += s
        a[1] += ins
        a[2] += th
    tot = [sum(v[k] for v in per.values()) for k in range(3)]
    acc = {name: [0, 0, 0] for name, _, _ in regions}
    acc["other"] = [0, 0, 0]
    others = []
    for line, v in per.items():
        for name, lo, hi in regions:
            if lo <= line <= hi:
                for k in range(3):
                    acc[name][k] += v[k]
                break
        else:
            for k in range(3):
                acc["other"][k] += v[k]
            others.append((v[1], line))
    print(f"total: samples {tot[0]}  warp-inst {tot[1]}  thread-inst {tot[2]}")
    for name, v in acc.items():
        print(f"{name:14s} {100 * v[0] / max(1, tot[0]):5.1f}% samples {100 * v[1] / max(1, tot[1]):5.1f}% warp-inst "
              f"({v[1] / 1e6:7.1f} M)  active lanes {v[2] / max(1, v[1]):4.1f}")
    others.sort(reverse=True)
    print("other, top lines:", ", ".join(f"L{l}:{100 * i / max(1, tot[1]):.1f}%" for i, l in others[:12]))


if __name__ == "__main__":
    main()
