import sys, time, ctypes as C
sys.path.insert(0, "svjedi-graph_b200"); sys.path.insert(0, ".")
import numpy as np, torch
from svjg import alnfilter, capi, synth
import io
g, vcf, gaf = synth.make_workload("C2", scale=1.0, stream0=0)
buf = io.StringIO(); g.write_gfa(buf)
tables = alnfilter.Tables.from_memory(g.edges_json(), buf.getvalue()).to_device(0)
raw = gaf.encode(); n = len(raw)
h = torch.frombuffer(bytearray(raw), dtype=torch.uint8).pin_memory()
def T(f, k=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e3
print("filter_host want_hits", T(lambda: alnfilter.filter_host(tables, h, hit_cap=1109933)))
print("filter_host no hits  ", T(lambda: alnfilter.filter_host(tables, h, want_hits=False)))
d = torch.empty(n, dtype=torch.uint8, device="cuda")
print("plain H2D copy       ", T(lambda: d.copy_(h, non_blocking=True)))
# pinned outputs
cap = 1109933
sv2 = torch.empty(cap, dtype=torch.int32).pin_memory(); off = torch.empty(cap, dtype=torch.int64).pin_memory(); ln = torch.empty(cap, dtype=torch.int32).pin_memory()
counts = torch.zeros((tables.num_sv, 2), dtype=torch.int32).pin_memory(); st = capi.FilterStats()
a = h.numpy()
def pinned():
    capi.check(capi.lib.svjg_filter_host(tables._h, a.ctypes.data, n, 100, counts.data_ptr(), sv2.data_ptr(), off.data_ptr(), ln.data_ptr(), cap, C.byref(st)))
print("C ABI pinned outputs ", T(pinned))
