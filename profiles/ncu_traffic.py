#!/usr/bin/env python3
"""profiles/traffic.json from ncu launch lists (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum,smsp__inst_executed.sum --csv`): per workload the DRAM bytes, device time and warp
instructions of ONE filter chain (probe + scan + exact kernels), the median over the chains in the list.
bench.py reads the table for `roofline.traffic`.
    python profiles/ncu_traffic.py C2:scaled=profiles/r2/launches_c2.csv C5:full=profiles/r2/launches_c5full.csv ..."""
import csv
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHAIN = ("probe_kernel", "scan_kernel", "exact_kernel")


def chains(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    by_id = {}
    for r in rows[hdr + 1:]:
        if len(r) < len(h):
            continue
        rec = dict(zip(h, r))
        d = by_id.setdefault(int(rec["ID"]), {"name": rec["Kernel Name"]})
        d[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3,
                                                                              "ms": 1e6, "inst": 1}.get(rec["Metric Unit"], 1)
    out, cur = [], None
    for k in sorted(by_id):
        d = by_id[k]
        short = next((c for c in CHAIN if c in d["name"]), None)
        if short == "probe_kernel":
            cur = {}
        if cur is not None and short:
            cur[short] = d
            if short == "exact_kernel":
                out.append(cur)
                cur = None
    return out


def main():
    table = {}
    for arg in sys.argv[1:]:
        key, path = arg.split("=", 1)
        cs = chains(path)
        if not cs:
            continue
        med = lambda f: statistics.median(f(c) for c in cs)
        table[key] = {
            "dram_bytes": int(med(lambda c: sum(k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"] for k in c.values()))),
            "dram_read_bytes": int(med(lambda c: sum(k["dram__bytes_read.sum"] for k in c.values()))),
            "chain_ns_under_ncu": int(med(lambda c: sum(k["gpu__time_duration.sum"] for k in c.values()))),
            "scan_share_of_chain_time": round(med(lambda c: c["scan_kernel"]["gpu__time_duration.sum"] / sum(k["gpu__time_duration.sum"] for k in c.values())), 4),
            "warp_instructions": int(med(lambda c: sum(k["smsp__inst_executed.sum"] for k in c.values()))),
            "chains_in_list": len(cs), "source": os.path.relpath(path, ROOT),
        }
    out = os.path.join(ROOT, "profiles", "traffic.json")
    old = {}
    if os.path.exists(out):
        old = json.load(open(out))
    old.update(table)
    with open(out, "w") as fh:
        json.dump(old, fh, indent=1, sort_keys=True)
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main()
