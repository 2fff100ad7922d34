"""Deterministic synthetic workloads for the hot path.

minigraph and real read sets are not available offline, so every GAF is
synthesised here: an SV catalogue (VCF rows) -> variation-graph tables in
construct-graph's formats (:mod:`svjg.graphgen`) -> minigraph-style GAF lines
from reads laid on two haplotypes per chromosome.

Nothing here is on the timed path; bench.py and the tests call it to make
inputs.  All randomness comes from ``numpy.random.Generator(PCG64(seed))``.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import graphgen

CHROMS24 = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]
# GRCh38-like lengths (Mbp) for the 24 names above, total ~3.1 Gbp
_HUMAN_MBP = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107,
              102, 90, 83, 80, 58, 64, 46, 50, 156, 57]


def human_like_chroms(scale=1.0):
    return OrderedDict((c, max(20000, int(m * 1_000_000 * scale))) for c, m in zip(CHROMS24, _HUMAN_MBP))


def _rand_seq(rng, n):
    return rng.choice(np.frombuffer(b"ACGT", dtype="S1"), size=n).tobytes().decode()


# --------------------------------------------------------------------------
# catalogues
# --------------------------------------------------------------------------
def catalogue(kind, n_sv, chrom_len, seed, ins_max=None, del_max=None):
    """VCF body rows for one of the BASELINE.json config shapes.

    kind: 'delins'  (C2: isolated DEL / INS, 50 % each)
          'cluster' (C3: clustered, overlapping DEL/INS/INV, same-POS INS)
          'bnd'     (C4: the four BND ALT forms, intra + inter chromosomal, plus
                     duplications written as SVTYPE=INS)
          'mix'     (C5: 70 % delins, 20 % cluster, 10 % bnd)
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    chroms = list(chrom_len)
    weights = np.array([chrom_len[c] for c in chroms], dtype=float)
    weights /= weights.sum()
    ins_max = ins_max or 5000
    del_max = del_max or 10000
    rows = []          # (chrom_idx, pos, text)
    vid = [0]

    def emit(ci, pos, ref, alt, info):
        vid[0] += 1
        rows.append((ci, pos, vid[0], f"{chroms[ci]}\t{pos}\tsv{vid[0]}\t{ref}\t{alt}\t.\t.\t{info}"))

    def one_del(ci, pos, ln):
        emit(ci, pos, "N", "<DEL>", f"SVTYPE=DEL;END={pos + ln};SVLEN=-{ln}")

    def one_ins(ci, pos, ln, tag="INS"):
        emit(ci, pos, "N", _rand_seq(rng, ln), f"SVTYPE=INS;END={pos + 1};SVLEN={ln}" + (";DUP=1" if tag == "DUP" else ""))

    def one_inv(ci, pos, ln):
        emit(ci, pos, "N", "<INV>", f"SVTYPE=INV;END={pos + ln};SVLEN={ln}")

    def span(ci):
        return chrom_len[chroms[ci]]

    def gen_delins(n):
        per = rng.multinomial(n, weights)
        for ci, k in enumerate(per):
            if k == 0:
                continue
            L = span(ci)
            gap = (L - 4000) // (k + 1)
            if gap < 200:
                raise ValueError("chromosome too short for that many isolated SVs")
            for j in range(k):
                lo = 2000 + j * gap
                room = gap - 60
                if rng.random() < 0.5:
                    ln = int(min(rng.integers(50, del_max + 1), max(50, room // 2)))
                    pos = int(lo + rng.integers(0, max(1, room - ln)))
                    one_del(ci, pos, ln)
                else:
                    ln = int(rng.integers(50, ins_max + 1))
                    pos = int(lo + rng.integers(0, max(1, room)))
                    one_ins(ci, pos, ln)

    def gen_cluster(n):
        made = 0
        while made < n:
            ci = int(rng.choice(len(chroms), p=weights))
            L = span(ci)
            size = int(min(n - made, rng.integers(3, 13)))
            pos = int(rng.integers(2000, max(2001, L - 60000)))
            for _ in range(size):
                r = rng.random()
                if pos + 3000 >= L - 2:
                    break
                if r < 0.40:
                    ln = int(rng.integers(50, 1200))
                    one_del(ci, pos, ln)
                    step = int(rng.integers(-ln // 2, ln + 300))      # may overlap / be included
                elif r < 0.75:
                    ln = int(rng.integers(50, min(ins_max, 800) + 1))
                    one_ins(ci, pos, ln)
                    if rng.random() < 0.25:                            # second INS at the same POS
                        one_ins(ci, pos, int(rng.integers(50, min(ins_max, 800) + 1)))
                        made += 1
                    step = int(rng.integers(0, 300))
                else:
                    ln = int(rng.integers(60, 1500))
                    one_inv(ci, pos, ln)
                    step = int(rng.integers(ln // 3, ln + 300))
                made += 1
                pos += max(1, step)

    def gen_bnd(n):
        for _ in range(n):
            ci = int(rng.choice(len(chroms), p=weights))
            L = span(ci)
            pos = int(rng.integers(2000, L - 2000))
            r = rng.random()
            if r < 0.25:                                               # duplication as INS
                one_ins(ci, pos, int(rng.integers(50, min(ins_max, 3000) + 1)), tag="DUP")
                continue
            cj = ci if rng.random() < 0.4 else int(rng.choice(len(chroms), p=weights))
            p2 = int(rng.integers(2000, span(cj) - 2000))
            form = int(rng.integers(0, 4))
            mate = f"{chroms[cj]}:{p2}"
            alt = (f"N[{mate}[", f"N]{mate}]", f"]{mate}]N", f"[{mate}[N")[form]
            emit(ci, pos, "N", alt, f"SVTYPE=BND;END={pos + 1};SVLEN=0")

    if kind == "delins":
        gen_delins(n_sv)
    elif kind == "cluster":
        gen_cluster(n_sv)
    elif kind == "bnd":
        gen_bnd(n_sv)
    elif kind == "mix":
        n_c = n_sv // 5
        n_b = n_sv // 10
        gen_delins(n_sv - n_c - n_b)
        gen_cluster(n_c)
        gen_bnd(n_b)
    else:
        raise ValueError(kind)
    rows.sort(key=lambda r: (r[0], r[1], r[2]))
    return [r[3] for r in rows]


VCF_HEADER = (
    "##fileformat=VCFv4.2\n"
    "##source=svjg_b200.synth\n"
    '##INFO=<ID=SVTYPE,Number=1,Type=String,Description="Type of structural variant">\n'
    '##INFO=<ID=END,Number=1,Type=Integer,Description="End position of the variant described in this record">\n'
    '##INFO=<ID=SVLEN,Number=1,Type=Integer,Description="Difference in length between REF and ALT alleles">\n'
    '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
    "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
)


def vcf_text(rows):
    return VCF_HEADER + "".join(r + "\n" for r in rows)


# --------------------------------------------------------------------------
# haplotypes and reads
# --------------------------------------------------------------------------
class _Hap:
    __slots__ = ("names", "orient", "lens", "cum", "fwd", "fwd_off", "rev", "rev_off", "total")

    def __init__(self, names, orient, lens):
        self.names, self.orient = names, orient
        self.lens = np.asarray(lens, dtype=np.int64)
        self.cum = np.concatenate(([0], np.cumsum(self.lens)))
        self.total = int(self.cum[-1])
        pieces = [(">" if o else "<") + n for n, o in zip(names, orient)]
        self.fwd = "".join(pieces)
        self.fwd_off = np.concatenate(([0], np.cumsum([len(p) for p in pieces])))
        rpieces = [("<" if o else ">") + n for n, o in zip(reversed(names), reversed(orient))]
        self.rev = "".join(rpieces)
        self.rev_off = np.concatenate(([0], np.cumsum([len(p) for p in rpieces])))


def _build_haps(g, rng):
    """Two haplotypes per chromosome.  Each DEL/INS/INV is drawn 0/0 : 0/1 : 1/1
    = 1:2:1; an SV is dropped from a haplotype when it collides with one already
    laid there."""
    haps = {}
    for chrom in g.chrom_len:
        svs = []
        for sv_id in g.svs[chrom]:
            kind = sv_id.split("-")[0]
            if kind == "BND" or (chrom, sv_id) not in g.sv_alt_links:
                continue
            a, b = sv_id.split("-")[1:]
            pos = int(a)
            end = pos if kind == "INS" else int(b)
            svs.append((pos, end, kind, sv_id))
        svs.sort(key=lambda t: (t[0], t[1]))
        draws = rng.integers(0, 4, size=len(svs))       # 0: none, 1: hap0, 2: hap1, 3: both
        nodes = g.nodes[chrom]
        both = []
        for h in (0, 1):
            names, orient, lens = [], [], []
            cursor, busy = 0, 0

            def ref_run(upto):
                nonlocal cursor
                for k in range(cursor, upto + 1):
                    s, e = nodes[k]
                    names.append(f"{chrom}:{s}-{e}")
                    orient.append(True)
                    lens.append(e - s + 1)
                cursor = upto + 1

            for (pos, end, kind, sv_id), d in zip(svs, draws):
                if not (d == 3 or d == h + 1) or pos < busy:
                    continue
                il = g.by_end[chrom].get(pos)
                ir = g.by_start[chrom].get(end + 1)
                if il is None or ir is None or il < cursor - 1 or ir <= il:
                    continue
                ref_run(il)
                if kind == "DEL":
                    cursor = ir
                elif kind == "INS":
                    ins = g.sv_ins_node[(chrom, sv_id)]
                    names.append(ins)
                    orient.append(True)
                    lens.append(len(g.alt_nodes[ins]))
                else:                                   # INV: inner nodes reversed, '<'
                    for k in range(ir - 1, il, -1):
                        s, e = nodes[k]
                        names.append(f"{chrom}:{s}-{e}")
                        orient.append(False)
                        lens.append(e - s + 1)
                    cursor = ir
                busy = end + 1
            ref_run(len(nodes) - 1)
            both.append(_Hap(names, orient, lens))
        haps[chrom] = both
    return haps


def _hex_names(rng, n):
    h = rng.bytes(16 * n).hex()
    return [f"{h[i:i + 8]}-{h[i + 8:i + 12]}-{h[i + 12:i + 16]}-{h[i + 16:i + 20]}-{h[i + 20:i + 32]}"
            for i in range(0, 32 * n, 32)]


def simulate_gaf(g, n_records, seed, mean_len=31000, sigma=0.45, edge_frac=0.05,
                 bnd_frac=None, cg_frac=0.0, idf=False, bare_single=True, haps=None, stream=0):
    """Returns the GAF text (str).  Reads are laid on haplotypes; BND junction
    reads are made explicitly from the SV's alt link.  ``edge_frac`` of the
    multi-node records are clipped so that one side overlaps the breakpoint by
    98..101 bases (the filter's threshold is 100).  The haplotypes depend on
    ``seed`` only; ``stream`` selects an independent stream of reads on them
    (used by :func:`simulate_gaf_parallel` and for per-rank shards)."""
    if haps is None:
        haps = _build_haps(g, np.random.Generator(np.random.PCG64(seed)))
    rng = np.random.Generator(np.random.PCG64([seed, 7919 + stream]))
    chroms = list(g.chrom_len)
    bnd_links = [(c, sv, lk) for (c, sv), lks in g.sv_alt_links.items() if sv.startswith("BND-") for lk in lks]
    if bnd_frac is None:
        n_all = sum(len(v) for v in g.svs.values())
        bnd_frac = 0.0 if not bnd_links else min(0.5, 0.6 * len(bnd_links) / max(1, n_all))
    n_bnd = int(n_records * bnd_frac) if bnd_links else 0
    n_hap = n_records - n_bnd

    w = np.array([haps[c][0].total for c in chroms], dtype=float)
    w /= w.sum()
    ci = rng.choice(len(chroms), size=n_hap, p=w)
    hi = rng.integers(0, 2, size=n_hap)
    mu = np.log(mean_len) - 0.5 * sigma * sigma
    rlen = np.maximum(500, rng.lognormal(mu, sigma, size=n_hap).astype(np.int64))
    u = rng.random(n_hap)
    rev = rng.random(n_hap) < 0.5
    edge = rng.random(n_hap) < edge_frac
    edge_side = rng.integers(0, 2, size=n_hap)
    edge_r = rng.integers(98, 102, size=n_hap)
    ident = rng.random(n_records) * 0.08 + 0.88
    dv = rng.random(n_records) * 0.12
    names = _hex_names(rng, n_records)
    cgs = rng.random(n_records) < cg_frac if cg_frac > 0 else None

    lines = [None] * n_records
    order = rng.permutation(n_records)       # interleave BND reads among the others

    def fmt(idx, path, tlen, ts, te, n_nodes):
        alen = te - ts
        am = int(alen * ident[idx])
        qlen = alen + int(dv[idx] * 200)
        tags = f"tp:A:P\tcm:i:{am // 12}\ts1:i:{am - 37}\ts2:i:{am // 3}\tdv:f:{dv[idx]:.4f}"
        if idf:
            tags += f"\tid:f:{ident[idx]:.6f}"
        if cgs is not None and cgs[idx]:
            tags += f"\tcg:Z:{alen // 2}M3I{alen - alen // 2}M"
        if n_nodes == 1 and bare_single and (idx & 1):
            path = path[1:]                   # minigraph prints a lone '+' segment bare
        return f"{names[idx]}\t{qlen}\t{int(dv[idx] * 100)}\t{qlen}\t+\t{path}\t{tlen}\t{ts}\t{te}\t{am}\t{alen}\t60\t{tags}\n"

    for k in range(n_hap):
        hap = haps[chroms[ci[k]]][hi[k]]
        L = int(min(rlen[k], hap.total))
        start = int(u[k] * (hap.total - L + 1))
        stop = start + L                       # exclusive, haplotype coordinates
        i0 = int(np.searchsorted(hap.cum, start, side="right")) - 1
        i1 = int(np.searchsorted(hap.cum, stop, side="left")) - 1
        n_nodes = i1 - i0 + 1
        tlen = int(hap.cum[i1 + 1] - hap.cum[i0])
        if not rev[k]:
            path = hap.fwd[hap.fwd_off[i0]:hap.fwd_off[i1 + 1]]
            ts = start - int(hap.cum[i0])
            first_len, last_len = int(hap.lens[i0]), int(hap.lens[i1])
        else:
            m = len(hap.names)
            path = hap.rev[hap.rev_off[m - 1 - i1]:hap.rev_off[m - i0]]
            ts = int(hap.cum[i1 + 1]) - stop
            first_len, last_len = int(hap.lens[i1]), int(hap.lens[i0])
        te = ts + L
        if n_nodes >= 2 and edge[k]:
            r = int(edge_r[k])
            if edge_side[k] == 0 and first_len > r:
                ts = first_len - r
            elif last_len > r:
                te = tlen - 1 - last_len + r
            if te <= ts:
                te = ts + 1
        lines[order[k]] = fmt(order[k], path, tlen, ts, te, n_nodes)

    if n_bnd:
        pick = rng.integers(0, len(bnd_links), size=n_bnd)
        flip = rng.random(n_bnd) < 0.5
        offs = rng.integers(0, 4000, size=(n_bnd, 2))
        for k in range(n_bnd):
            _, _, (nl, sl, nr, sr) = bnd_links[pick[k]]
            ll = _ref_len(nl)
            lr = _ref_len(nr)
            if not flip[k]:
                path = (">" if sl == "+" else "<") + nl + (">" if sr == "+" else "<") + nr
                a_len, b_len = ll, lr
            else:
                path = ("<" if sr == "+" else ">") + nr + ("<" if sl == "+" else ">") + nl
                a_len, b_len = lr, ll
            tlen = a_len + b_len
            ts = max(0, a_len - 50 - int(offs[k, 0]))
            te = min(tlen, a_len + 50 + int(offs[k, 1]))
            idx = order[n_hap + k]
            lines[idx] = fmt(idx, path, tlen, ts, te, 2)
    return "".join(lines)


_PAR = {}


def _par_worker(args):
    k, n, kw = args
    return simulate_gaf(_PAR["g"], n, _PAR["seed"], haps=_PAR["haps"], stream=_PAR["stream0"] + k, **kw)


def simulate_gaf_parallel(g, n_records, seed, procs=None, chunk=200_000, stream0=0, **kw):
    """Same distribution as :func:`simulate_gaf`, generated in ``chunk``-record
    pieces on a fork pool; deterministic in (seed, chunk, stream0), independent
    of ``procs``."""
    import multiprocessing as mp
    import os
    haps = _build_haps(g, np.random.Generator(np.random.PCG64(seed)))
    sizes = [chunk] * (n_records // chunk) + ([n_records % chunk] if n_records % chunk else [])
    procs = procs or min(len(sizes), max(1, (os.cpu_count() or 2) - 1), 32)
    _PAR.update(g=g, seed=seed, haps=haps, stream0=stream0 * 100_000)
    jobs = [(k, n, kw) for k, n in enumerate(sizes)]
    if procs <= 1 or len(sizes) == 1:
        parts = [_par_worker(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(procs) as pool:
            parts = pool.map(_par_worker, jobs, chunksize=1)
    _PAR.clear()
    return "".join(parts)


def _ref_len(name):
    s, e = name.rsplit(":", 1)[1].split("-")
    return int(e) - int(s) + 1


# --------------------------------------------------------------------------
# named workloads (BASELINE.json configs)
# --------------------------------------------------------------------------
WORKLOADS = {
    # name: (kind, n_sv, n_records, genome scale, mean read length)
    "C2": ("delins", 25_000, 3_000_000, 1.0, 31_000),
    "C3": ("cluster", 100_000, 6_000_000, 1.0, 31_000),
    "C4": ("bnd", 20_000, 3_000_000, 1.0, 31_000),
    "C5": ("mix", 1_000_000, 200_000_000, 1.0, 15_500),
}


def make_workload(name, scale=1.0, seed=None, stream0=0, catalogue_scale=None, **gaf_kw):
    """(Graph, vcf_text, gaf_text) for a named config shrunk by ``scale`` in both
    SV count and record count (genome scaled alike so densities stay put).
    ``catalogue_scale`` sizes the SV catalogue and the genome on their own:
    ``make_workload("C5", 0.02, catalogue_scale=1.0)`` is a 4 M-record shard of the
    population config against its full 1 M-SV tables."""
    kind, n_sv, n_rec, gscale, mean_len = WORKLOADS[name]
    seed = seed if seed is not None else 1000 + int(name[1:])
    cscale = scale if catalogue_scale is None else catalogue_scale
    n_sv = max(24, int(n_sv * cscale))
    n_rec = max(100, int(n_rec * scale))
    chrom_len = human_like_chroms(gscale * max(cscale, 0.002))
    ins_max = 5000 if name != "C5" else 600
    rows = catalogue(kind, n_sv, chrom_len, seed, ins_max=ins_max)
    g = graphgen.build_graph(chrom_len, rows)
    if n_rec > 400_000:
        gaf = simulate_gaf_parallel(g, n_rec, seed + 7, mean_len=mean_len, stream0=stream0, **gaf_kw)
    else:
        gaf = simulate_gaf(g, n_rec, seed + 7, mean_len=mean_len, stream=stream0, **gaf_kw)
    return g, vcf_text(rows), gaf
