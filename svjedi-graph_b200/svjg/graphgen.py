"""Variation-graph tables in construct-graph's exact output formats, built in
O(SVs log SVs) instead of the reference's per-SV scans over every node.

Used by the synthetic workload generator (bench / tests).  It produces the two
inputs of the hot path that the reference's graph stage produces:
``<prefix>.gfa`` and ``<prefix>_svs_edges.json`` (reference:
construct-graph.py:67-554; formats :556-582).  On catalogues small enough for
the reference to finish, ``tests/golden/make_golden.py`` checks the output of
this module byte-for-byte against the unmodified construct-graph.py.

This is SURVEY.md §8(f) row N2 only as far as the generator needs it: VCF
records come in as already-split columns and chromosome *lengths* are enough
(reference nodes then carry a ``*`` placeholder sequence, which the filter
never reads — filter-alignments.py:109-113 only measures alt nodes).
"""
from __future__ import annotations

import json
from collections import OrderedDict


def info_get(info, label):
    """Value of ``label=`` in a VCF INFO string (first / last / middle)."""
    parts = info.split(";")
    if parts[0].startswith(label + "="):
        return info.split(label + "=")[1].split(";")[0]
    if parts[-1].startswith(label + "="):
        return info.split(";" + label + "=")[1]
    return info.split(";" + label + "=")[1].split(";")[0]


def bnd_id(pos, alt):
    """``BND-<ALT with the REF base replaced by POS>`` (construct-graph.py:615-660)."""
    for br in "[]":
        if br in alt:
            pieces = [p for p in alt.split(br) if p]
            t = pieces[0] if ":" in pieces[1] else pieces[1]
            return "BND-" + alt.replace(t, pos)
    return "BND-format"


def bnd_coords(chrom, sv_id):
    """((chrom, pos, strand), (chrom, pos, strand)) of the two joined ends, or
    None for an unsupported ALT (construct-graph.py:662-734)."""
    alt = sv_id.split("BND-")[1]
    for br, far_strand in (("[", "+"), ("]", "-")):
        if br not in alt:
            continue
        pieces = [p for p in alt.split(br) if p]
        if ":" in pieces[1]:                       # t[p[  /  t]p]
            c2, p2 = pieces[1].split(":")[0], pieces[1].split(":")[1]
            return [chrom, int(pieces[0]), "+"], [c2, int(p2), far_strand]
        if ":" in pieces[0]:                       # [p[t  /  ]p]t
            c2, p2 = pieces[0].split(":")[0], pieces[0].split(":")[1]
            near = "-" if br == "[" else "+"
            return [c2, int(p2), near], [chrom, int(pieces[1]), "+"]
        return None
    return None


class Graph:
    """Result of :func:`build_graph`."""

    def __init__(self):
        self.chrom_len = OrderedDict()
        self.chrom_seq = None           # optional chrom -> str
        self.svs = OrderedDict()        # chrom -> [sv_id]      (d_svs)
        self.discarded = []             # raw VCF lines ignored
        self.bkpts = {}                 # chrom -> sorted [int]
        self.nodes = {}                 # chrom -> [(start1, end1)] 1-based inclusive
        self.by_end = {}                # chrom -> {end1: idx}
        self.by_start = {}              # chrom -> {start1: idx}
        self.link_sv = {}               # key -> [(chrom:sv_id, allele)]
        self.ins_seq = {}               # sv_id -> sequence
        self.alt_nodes = OrderedDict()  # alt node name -> sequence
        self.alt_lines = []             # GFA lines after the reference part
        self.sv_alt_links = {}          # (chrom, sv_id) -> [(nL, sL, nR, sR)]
        self.sv_ins_node = {}           # (chrom, sv_id) -> alt node name
        self.warnings = []

    def node_name(self, chrom, idx):
        s, e = self.nodes[chrom][idx]
        return f"{chrom}:{s}-{e}"

    def edges_json(self):
        return json.dumps(self.link_sv, sort_keys=True, indent=4)

    def write_gfa(self, fh):
        for chrom, lst in self.svs.items():
            if lst:
                fh.write("#{}\t{}\n".format(chrom, ";".join(lst)))
        for chrom in self.chrom_len:
            names, lens = [], []
            seq = self.chrom_seq[chrom] if self.chrom_seq else None
            prev = None
            for s, e in self.nodes[chrom]:
                name = f"{chrom}:{s}-{e}"
                fh.write("S\t{}\t{}\n".format(name, seq[s - 1:e] if seq is not None else "*"))
                if prev is not None:
                    fh.write("L\t{}\t+\t{}\t+\t0M\n".format(prev, name))
                prev = name
                names.append(name)
                lens.append(str(e - s + 1))
            fh.write("P\t{}\t{}\t{}\n".format(chrom, "+,".join(names) + "+", "M,".join(lens) + "M"))
        for line in self.alt_lines:
            fh.write(line)

    def ignored_text(self):
        return "##The following SVs were ignored during graph construction due to wrong format" + "".join(
            "\n" + d for d in self.discarded)


def build_graph(chrom_len, vcf_rows, chrom_seq=None):
    """``chrom_len``: ordered chrom -> length (FASTA order); ``vcf_rows``:
    iterable of VCF body lines (str, tab separated, no newline needed)."""
    g = Graph()
    g.chrom_len = OrderedDict(chrom_len)
    g.chrom_seq = chrom_seq
    bk_set = {c: set() for c in g.chrom_len}
    bk_sv = {}
    for c in g.chrom_len:
        g.svs[c] = []
    ins_mult = {}

    def add_bkpt(c, p, sv_id):
        if 1 < p < g.chrom_len[c]:
            bk_set[c].add(p)
            bk_sv.setdefault(c, {}).setdefault(p, []).append(sv_id)

    for raw in vcf_rows:
        raw = raw.rstrip()
        if not raw or raw.startswith("#"):
            continue
        cols = raw.split("\t")
        chrom, pos, _vid, ref, alt, info = cols[0], cols[1], cols[2], cols[3], cols[4], cols[7]
        svtype = info_get(info, "SVTYPE")
        start = int(pos)
        if chrom not in g.chrom_len:
            raise SystemExit(f"Error: sequence '{chrom}' from input VCF is missing in reference genome")
        if svtype == "DEL" or svtype == "INV":
            end = int(info_get(info, "END"))
            sv_id = f"{svtype}-{pos}-{end}"
        elif svtype == "INS":
            end = start
            ins_mult[pos] = ins_mult.get(pos, 0) + 1
            sv_id = f"INS-{pos}-{ins_mult[pos]}"
            if len(ref) > 1:
                g.discarded.append(raw)
                continue
            if alt.startswith("<"):
                if "LEFT_SVINSSEQ=" in info or "RIGHT_SVINSSEQ=" in info:
                    g.discarded.append(raw)
                    continue
                if "SEQ=" in info:
                    g.ins_seq[sv_id] = info_get(info, "SEQ")
                else:
                    g.discarded.append(raw)
                    continue
            elif sv_id not in g.ins_seq:
                g.ins_seq[sv_id] = alt.upper()
        elif svtype == "BND":
            sv_id = bnd_id(str(start), alt)
        else:
            continue

        if svtype in ("DEL", "INS", "INV"):
            clen = g.chrom_len[chrom]
            if end >= clen - 1 or start >= clen - 1:
                g.discarded.append(raw)
                continue
            for p in {start, end}:
                add_bkpt(chrom, p, sv_id)
            g.svs[chrom].append(sv_id)
        else:
            coords = bnd_coords(chrom, sv_id)
            if coords is None:
                g.discarded.append(raw)
                continue
            left, right = coords
            if left[2] == "+" and right[2] == "+":
                right[1] -= 1
            elif left[2] == "-":
                left[1] -= 1
                right[1] -= 1
            for c, p in ((left[0], left[1]), (right[0], right[1])):
                add_bkpt(c, p, sv_id)
            g.svs[chrom].append(sv_id)

    # reference part: nodes, reference links and their allele-0 entries
    for chrom, clen in g.chrom_len.items():
        bks = sorted(b for b in bk_set[chrom] if b < clen - 1)
        g.bkpts[chrom] = bks
        edges = [0] + bks + [clen]
        nodes = [(edges[i] + 1, edges[i + 1]) for i in range(len(edges) - 1)]
        g.nodes[chrom] = nodes
        g.by_end[chrom] = {e: i for i, (_, e) in enumerate(nodes)}
        g.by_start[chrom] = {s: i for i, (s, _) in enumerate(nodes)}
        for i, b in enumerate(bks):
            key = f"{chrom}:{nodes[i][0]}-{nodes[i][1]}@+@{chrom}:{nodes[i + 1][0]}-{nodes[i + 1][1]}@+"
            g.link_sv[key] = [(f"{chrom}:{sv}", 0) for sv in bk_sv[chrom][b]]

    def add_alt(chrom, sv_id, link):
        key = "@".join(link)
        g.link_sv.setdefault(key, []).append((f"{chrom}:{sv_id}", 1))
        g.sv_alt_links.setdefault((chrom, sv_id), []).append(link)

    def gfa_link(link):
        return "L\t{}\t{}\t{}\t{}\t0M\n".format(*link)

    def name_at(chrom, table, p):
        idx = table[chrom].get(p)
        return None if idx is None else g.node_name(chrom, idx)

    for chrom, lst in g.svs.items():
        for sv_id in lst:
            kind = sv_id.split("-")[0]
            if kind == "DEL":
                pos, end = (int(x) for x in sv_id.split("-")[1:])
                ln, rn = name_at(chrom, g.by_end, pos), name_at(chrom, g.by_start, end + 1)
                if ln is None or rn is None:
                    raise TypeError(f"{sv_id}: flanking node missing (the reference crashes here too)")
                link = (ln, "+", rn, "+")
                g.alt_lines.append(gfa_link(link))
                add_alt(chrom, sv_id, link)
            elif kind == "INS":
                pos, cnt = sv_id.split("-")[1:]
                pos = int(pos)
                ins = f"{chrom}:{pos + 1}.{cnt}"
                g.alt_nodes[ins] = g.ins_seq[sv_id]
                g.sv_ins_node[(chrom, sv_id)] = ins
                g.alt_lines.append("S\t{}\t{}\n".format(ins, g.ins_seq[sv_id]))
                ln, rn = name_at(chrom, g.by_end, pos), name_at(chrom, g.by_start, pos + 1)
                if ln is None or rn is None:
                    raise TypeError(f"{sv_id}: flanking node missing (the reference crashes here too)")
                for link in ((ln, "+", ins, "+"), (ins, "+", rn, "+")):
                    g.alt_lines.append(gfa_link(link))
                    add_alt(chrom, sv_id, link)
            elif kind == "INV":
                pos, end = (int(x) for x in sv_id.split("-")[1:])
                ln, rn = name_at(chrom, g.by_end, pos), name_at(chrom, g.by_start, end + 1)
                # the reference's elif chain: a node already taken as left/right
                # flank is not considered as an inner node
                li = g.by_start[chrom].get(pos + 1)
                ri = g.by_end[chrom].get(end)
                lin = rin = None
                if li is not None:
                    s, e = g.nodes[chrom][li]
                    if e != pos and s != end + 1:
                        lin = g.node_name(chrom, li)
                if ri is not None:
                    s, e = g.nodes[chrom][ri]
                    if e != pos and s != end + 1:
                        rin = g.node_name(chrom, ri)
                if None in (ln, rn, lin, rin):
                    continue
                for link in ((ln, "+", rin, "-"), (lin, "-", rn, "+")):
                    g.alt_lines.append(gfa_link(link))
                    add_alt(chrom, sv_id, link)
            elif kind == "BND":
                left, right = bnd_coords(chrom, sv_id)
                ln = name_at(left[0], g.by_start if left[2] == "-" else g.by_end, left[1])
                rn = name_at(right[0], g.by_start if right[2] == "+" else g.by_end, right[1])
                if ln is None or rn is None:
                    g.warnings.append(f"Warning: no alternative link defined for {sv_id}")
                    continue
                if left[2] == "-":
                    link = (ln, "-", rn, "+")
                elif right[2] == "-":
                    link = (ln, "+", rn, "-")
                else:
                    link = (ln, "+", rn, "+")
                g.alt_lines.append(gfa_link(link))
                add_alt(chrom, sv_id, link)
    return g
