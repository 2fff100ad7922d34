import json, sys
sys.path.insert(0, "svjedi-graph_b200"); sys.path.insert(0, ".")
from svjg import alnfilter
from oracle import svjg_oracle as O
n_nodes, step = 120, 500
names = [f"chrL:{i * step + 1}-{(i + 1) * step}" for i in range(n_nodes)]
edges = {}
for i in range(n_nodes - 1):
    edges[f"{names[i]}@+@{names[i + 1]}@+"] = [[f"chrL:DEL-{(i + 1) * step}-{(i + 1) * step + 40}", 0]]
t = alnfilter.Tables.from_memory(json.dumps(edges), "").to_device(0)
for k in (2, 31, 32, 33, 34, 40, 63, 64, 65):
    idx = list(range(3, 3 + k))
    tlen = k * step
    path = "".join(">" + names[i] for i in idx)
    line = f"read\t{tlen}\t0\t{tlen}\t+\t{path}\t{tlen}\t120\t{tlen - 130}\t{tlen - 9}\t{tlen}\t60\ttp:A:P\n"
    res = alnfilter.filter_host(t, line.encode())
    want = O.hit_counts(O.filter_alignments([line], edges, {}))
    got = {t.sv_ids[i]: [int(res.counts[i, 0]), int(res.counts[i, 1])] for i in range(t.num_sv) if res.counts[i].any()}
    w = {a: list(b) for a, b in want.items()}
    miss = sorted(set(w) - set(got)); extra = sorted(set(got) - set(w))
    print(k, "ok" if got == w else f"MISMATCH missing {miss[:6]} extra {extra[:6]} n_got {len(got)} n_want {len(w)}", res.stats["n_generic"])
