#!/usr/bin/env python3
"""Wall-clock of the drop-in construct-graph.py beside the unmodified reference script on the same
catalogue (build container only: needs /root/reference).  Host-only stage, no GPU involved.

    python profiles/construct_timing.py <dir with *.vcf/*.fa pairs> [tag ...]
Outputs are compared byte for byte (GFA, svs_edges.json, ignored_svs.txt, stdout)."""
import filecmp
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(script, tag, who, cwd):
    t0 = time.perf_counter()
    with open(os.path.join(cwd, f"{who}_{tag}.out"), "w") as out:
        rc = subprocess.run([sys.executable, script, "-v", f"{tag}.vcf", "-r", f"{tag}.fa", "-o", f"{who}_{tag}.gfa"],
                            cwd=cwd, stdout=out).returncode
    return time.perf_counter() - t0, rc


def main():
    cwd = sys.argv[1]
    tags = sys.argv[2:] or sorted(f[:-4] for f in os.listdir(cwd) if f.endswith(".vcf"))
    print("catalogue\tSVs\tgenome bp\tdrop-in s\treference s\tratio\toutputs")
    for tag in tags:
        n_sv = sum(1 for l in open(os.path.join(cwd, tag + ".vcf")) if not l.startswith("#"))
        bp = sum(len(l) - 1 for l in open(os.path.join(cwd, tag + ".fa")) if not l.startswith(">"))
        ours, rc1 = run(os.path.join(ROOT, "svjedi-graph_b200", "construct-graph.py"), tag, "ours", cwd)
        ref, rc2 = run("/root/reference/construct-graph.py", tag, "ref", cwd)
        same = rc1 == rc2 and all(filecmp.cmp(os.path.join(cwd, f"ours_{tag}{s}"), os.path.join(cwd, f"ref_{tag}{s}"), shallow=False)
                                  for s in (".gfa", "_svs_edges.json", "_ignored_svs.txt", ".out"))
        print(f"{tag}\t{n_sv}\t{bp}\t{ours:.2f}\t{ref:.2f}\t{ref / ours:.1f}x\t{'identical' if same else 'DIFFERENT'}", flush=True)


if __name__ == "__main__":
    main()
