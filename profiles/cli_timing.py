#!/usr/bin/env python3
"""Wall-clock of the drop-in command lines on a BASELINE.json-sized input (files on local disk).
usage: cli_timing.py [workload=C2] [scale=1.0] [ref]   -> prints one line per stage
With `ref`, the unmodified reference scripts (baseline/_ref/, copied there by __graft_entry__.build()) run on the
same files afterwards and the two sets of output files are compared byte for byte."""
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
    r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, SVJG_TIMING="1"))
    dt = time.time() - t0
    print(f"{label:24s} rc={r.returncode} {dt:7.2f} s  {r.stdout.strip()[:60]}")
    for ln in r.stderr.strip().splitlines()[-24:]:
        print("    " + ln)
print("informative_aln.json", os.path.getsize(p + "_informative_aln.json") / 1e6, "MB;  genotype.vcf",
      os.path.getsize(p + "_genotype.vcf") / 1e6, "MB")

if "ref" in sys.argv[3:]:
    import filecmp
    ref = os.path.join(ROOT, "baseline", "_ref")
    q = os.path.join(d, "q")
    for ext in (".gfa", "_svs_edges.json"):
        if not os.path.exists(q + ext):
            os.symlink(p + ext, q + ext)
    for label, cmd in (
            ("reference filter-alignments.py", [sys.executable, os.path.join(ref, "filter-alignments.py"), "-a", p + ".gaf",
                                                "-g", q + ".gfa", "-p", q]),
            ("reference predict-genotype.py", [sys.executable, os.path.join(ref, "predict-genotype.py"), "-d",
                                               q + "_informative_aln.json", "-v", p + ".vcf", "-o", q + "_genotype.vcf"])):
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True)
        print(f"{label:32s} rc={r.returncode} {time.time() - t0:7.2f} s  {r.stderr.strip()[-200:]}")
    for ext in ("_informative_aln.json", "_genotype.vcf"):
        print(ext, "byte-equal to the reference's:", filecmp.cmp(p + ext, q + ext, shallow=False))
