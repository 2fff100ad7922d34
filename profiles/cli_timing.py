#!/usr/bin/env python3
"""Wall-clock of the drop-in command lines on a BASELINE.json-sized input (files on local disk).
usage: cli_timing.py [workload=C2] [scale=1.0]   -> prints one line per stage"""
import io
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "svjedi-graph_b200"))
from svjg import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
d = f"/tmp/svjg_cli_{name}"
os.makedirs(d, exist_ok=True)
g, vcf, gaf = synth.make_workload(name, scale=scale, stream0=0)
p = os.path.join(d, "p")
with open(p + ".gfa", "w") as fh:
    g.write_gfa(fh)
open(p + "_svs_edges.json", "w").write(g.edges_json())
open(p + ".gaf", "w").write(gaf)
open(p + ".vcf", "w").write(vcf)
print(f"{name} x{scale}: {gaf.count(chr(10))} records, {len(gaf) / 1e6:.1f} MB GAF, {vcf.count(chr(10))} VCF lines")
for label, cmd in (
        ("filter-alignments.py", [sys.executable, os.path.join(ROOT, "svjedi-graph_b200", "filter-alignments.py"), "-a", p + ".gaf",
                                  "-g", p + ".gfa", "-p", p]),
        ("predict-genotype.py", [sys.executable, os.path.join(ROOT, "svjedi-graph_b200", "predict-genotype.py"), "-d",
                                 p + "_informative_aln.json", "-v", p + ".vcf", "-o", p + "_genotype.vcf"])):
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    dt = time.time() - t0
    print(f"{label:24s} rc={r.returncode} {dt:7.2f} s  {r.stdout.strip()[:60]} {r.stderr.strip()[-200:]}")
print("informative_aln.json", os.path.getsize(p + "_informative_aln.json") / 1e6, "MB;  genotype.vcf",
      os.path.getsize(p + "_genotype.vcf") / 1e6, "MB")
