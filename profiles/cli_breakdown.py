#!/usr/bin/env python3
"""Where the wall-clock of the drop-in command lines goes (after profiles/cli_timing.py made the files)."""
import sys, time, os
t0 = time.time()
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "svjedi-graph_b200"))
def T(label):
    global t0
    t = time.time(); print(f"  {label:36s} {t - t0:6.2f} s", flush=True); t0 = t
import numpy as np; T("import numpy")
from svjg import alnfilter, capi, genotype; T("import svjg (ctypes, libsvjg.so)")
p = "/tmp/svjg_cli_C2/p"
tables = alnfilter.Tables.load(p + "_svs_edges.json", p + ".gfa"); T("tables: parse JSON + GFA, build")
tables.to_device(0); T("CUDA context + tables to device")
gaf = alnfilter.read_file_pinned(p + ".gaf"); T("pinned alloc + read GAF")
res = alnfilter.filter_host(tables, gaf); T("filter_host, first call")
res = alnfilter.filter_host(tables, gaf); T("filter_host, second call")
alnfilter.write_informative_json(tables, gaf, res, p + "_x.json"); T("write informative_aln.json")
counts = genotype.AlnCounts.load(p + "_x.json"); T("read informative_aln.json (C++)")
lines = open(p + ".vcf").readlines(); T("read vcf")
text, n = genotype.genotype_vcf_from_json(counts, lines, 3, 0.00005); T("genotype_vcf_from_json")
open(p + "_y.vcf", "w").write(text); T("write vcf")
