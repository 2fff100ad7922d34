"""Where an end-to-end step spends its time (page-locked GAF bytes in, JSON + VCF text out): python profiles/e2e_breakdown.py"""
import sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "svjedi-graph_b200")
import numpy as np, torch, io
from svjg import alnfilter, capi, genotype, synth
import pickle
g, vcf, gaf = pickle.load(open("/tmp/svjg_wl_C2_1_scaled_0.pkl","rb")) if os.path.exists("/tmp/svjg_wl_C2_1_scaled_0.pkl") else synth.make_workload("C2")
buf = io.StringIO(); g.write_gfa(buf)
tables = alnfilter.Tables.from_memory(g.edges_json(), buf.getvalue()).to_device(0)
h = torch.frombuffer(bytearray(gaf.encode()), dtype=torch.uint8).pin_memory()
hnp = h.numpy()
nvcf = genotype.NativeVcf.from_input(vcf.encode()); idx = nvcf.index_tables(tables); ty = nvcf.svtype
res = alnfilter.filter_host(tables, h)
out = alnfilter.HostBuffers(tables, res.n_hits + 1024)
for rep in range(3):
    t0 = time.perf_counter(); res = alnfilter.filter_host(tables, h, out=out)
    t1 = time.perf_counter(); gt, fl, ad, pl = genotype.genotype_host(res.counts, idx, ty)
    t2 = time.perf_counter(); text, n = nvcf.format(gt, fl, ad, pl)
    t3 = time.perf_counter(); js = alnfilter.JsonText(tables, hnp, res)
    t4 = time.perf_counter(); r2, text = alnfilter.filter_json_host(tables, h)
    t5 = time.perf_counter()
    sink = io.BytesIO(); nvcf.format(gt, fl, ad, pl, out=sink)
    t6 = time.perf_counter()
    print(f"filter_host {1e3*(t1-t0):.1f} ms  genotype_host {1e3*(t2-t1):.1f}  vcf format {1e3*(t3-t2):.1f}  host json {1e3*(t4-t3):.1f}  ({js.nbytes} B) | "
          f"filter + device json {1e3*(t5-t4):.1f} ms ({len(text)} B)  vcf format to sink {1e3*(t6-t5):.1f}")

# the halves of the pipelined end-to-end step (bench.py: two batches in flight), each alone
def avg(fn, n=10):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n
cnt = out.counts
print(f"begin alone (upload + filter + counters back)   {avg(lambda: alnfilter.filter_json_begin(tables, h, counts=cnt)):.2f} ms")
def both():
    alnfilter.filter_json_begin(tables, h, counts=cnt); alnfilter.filter_json_finish(tables)
print(f"begin + finish (text rendered, copied back)     {avg(both):.2f} ms")
def vcf_only():
    gt, fl, ad, pl = genotype.genotype_host(cnt, idx, ty); nvcf.format_buffer(gt, fl, ad, pl)
print(f"genotype_host + VCF text                        {avg(vcf_only):.2f} ms")
d = torch.empty(h.numel(), dtype=torch.uint8, device="cuda")
print(f"plain cudaMemcpyAsync of the {h.numel() / 1e6:.0f} MB               {avg(lambda: d.copy_(h, non_blocking=True)):.2f} ms")
