"""Variation-graph tables in construct-graph's exact output formats, built in
O(SVs log SVs) instead of the reference's per-SV scans over every node.

Used by the synthetic workload generator (bench / tests).  It produces the two
inputs of the hot path that the reference's graph stage produces:
``<prefix>.gfa`` and ``<prefix>_svs_edges.json`` (reference:
construct-graph.py:67-554; formats :556-582).  On catalogues small enough for
the reference to finish, ``tests/golden/make_golden.py`` checks the output of
this module byte-for-byte against the unmodified construct-graph.py.

This is SURVEY.md §8(f) row N2.  The workload generator passes chromosome
*lengths* only (reference nodes then carry a ``*`` placeholder sequence, which
the filter never reads — filter-alignments.py:109-113 only measures alt
nodes); the drop-in ``construct-graph.py`` front-end (:func:`construct_main`)
passes the sequences read by :func:`load_fasta`.  Inputs the reference stops on
(short lines, a DEL at POS 1, ...) stop this module with the same exception
class at the same stage, so the files left behind are the same too;
``tests/golden/make_graph_fuzz.py`` pins that on damaged catalogues.
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
    """The reference's ``construct_gfa`` state, built in three stages: :meth:`parse` (VCF body,
    construct-graph.py:102-273), :meth:`reference_part` (:293-378) and :meth:`alt_part` (:383-547)."""

    def __init__(self, chrom_len, chrom_seq=None, warn=None):
        self.chrom_len = OrderedDict(chrom_len)
        self.chrom_seq = chrom_seq      # optional chrom -> str
        self.svs = OrderedDict((c, []) for c in self.chrom_len)   # chrom -> [sv_id]      (d_svs)
        self.vcf_id = {}                # sv_id -> VCF ID column (d_sv_ID, last one wins)
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
        self._warn = warn or self.warnings.append
        self._bk_set = {c: set() for c in self.chrom_len}
        self._bk_sv = {}

    def node_name(self, chrom, idx):
        s, e = self.nodes[chrom][idx]
        return f"{chrom}:{s}-{e}"

    def edges_json(self):
        return json.dumps(self.link_sv, sort_keys=True, indent=4)

    def link_sv_as_json(self):
        """The table as ``json.load`` of the file gives it back (tuples become lists)."""
        return {k: [[sv, a] for sv, a in v] for k, v in self.link_sv.items()}

    def write_edges_json(self, fh):
        """The bytes of ``json.dumps(d_link_sv, sort_keys=True, indent=4)`` (construct-graph.py:553-554)
        without the pure-Python indenting encoder: one formatted block per key."""
        if not self.link_sv:
            fh.write("{}")
            return
        q = json.encoder.encode_basestring_ascii       # what json.dumps applies to every str (ensure_ascii)
        blocks, lead = [], "{\n"
        for key in sorted(self.link_sv):
            ents = self.link_sv[key]
            if ents:
                inner = ",\n".join(f"        [\n            {q(sv)},\n            {a}\n        ]" for sv, a in ents)
                blocks.append(f"    {q(key)}: [\n{inner}\n    ]")
            else:
                blocks.append(f"    {q(key)}: []")
            if len(blocks) >= 65536:
                fh.write(lead + ",\n".join(blocks))
                blocks, lead = [], ",\n"
        if blocks:
            fh.write(lead + ",\n".join(blocks))
        fh.write("\n}")

    def write_gfa(self, fh, with_alt=True):
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
        if with_alt:
            for line in self.alt_lines:
                fh.write(line)

    def ignored_text(self):
        return "##The following SVs were ignored during graph construction due to wrong format" + "".join(
            "\n" + d for d in self.discarded)

    # ------------------------------------------------------------------ stage 1
    def _add_bkpt(self, c, p, sv_id):
        if 1 < p < self.chrom_len[c]:
            self._bk_set[c].add(p)
            self._bk_sv.setdefault(c, {}).setdefault(p, []).append(sv_id)

    def parse(self, vcf_rows):
        """VCF body lines -> SV ids, breakpoints, ignored records (construct-graph.py:102-273).
        Raises what the reference raises: ValueError on a line with fewer than 8 columns (a blank
        line included) or a non-numeric POS/END, IndexError on a missing INFO label, KeyError on a
        BND mate on an unknown sequence, SystemExit on an unknown CHROM."""
        ins_mult = {}
        for line in vcf_rows:
            if line.startswith("#"):
                continue
            raw = line.rstrip()
            chrom, pos, vid, ref, alt, _, _, info, *_ = raw.split("\t")
            svtype = info_get(info, "SVTYPE")
            start = int(pos)
            if chrom not in self.chrom_len:
                raise SystemExit(f"Error: sequence '{chrom}' from input VCF is missing in reference genome, chromosomes "
                                 "must have the same names in input VCF and reference genome files")
            if svtype == "DEL" or svtype == "INV":
                end = int(info_get(info, "END"))
                sv_id = f"{svtype}-{pos}-{end}"
            elif svtype == "INS":
                end = start
                ins_mult[pos] = ins_mult.get(pos, 0) + 1
                sv_id = f"INS-{pos}-{ins_mult[pos]}"
                if len(ref) > 1:
                    self.discarded.append(raw)
                    continue
                if alt.startswith("<"):
                    if "LEFT_SVINSSEQ=" in info or "RIGHT_SVINSSEQ=" in info:
                        self.discarded.append(raw)
                        continue
                    if "SEQ=" in info:
                        self.ins_seq[sv_id] = info_get(info, "SEQ")
                    else:
                        self.discarded.append(raw)
                        continue
                elif sv_id not in self.ins_seq:
                    self.ins_seq[sv_id] = alt.upper()
            elif svtype == "BND":
                sv_id = bnd_id(str(start), alt)
            else:
                continue
            self.vcf_id[sv_id] = vid

            if svtype != "BND":
                clen = self.chrom_len[chrom]
                if end >= clen - 1 or start >= clen - 1:
                    self.discarded.append(raw)
                    continue
                for p in {start, end}:
                    self._add_bkpt(chrom, p, sv_id)
                self.svs[chrom].append(sv_id)
            else:
                coords = bnd_coords(chrom, sv_id)
                if coords is None:
                    self.discarded.append(raw)
                    continue
                left, right = coords
                if left[2] == "+" and right[2] == "+":
                    right[1] -= 1
                elif left[2] == "-":
                    left[1] -= 1
                    right[1] -= 1
                for c, p in ((left[0], left[1]), (right[0], right[1])):
                    self._add_bkpt(c, p, sv_id)
                self.svs[chrom].append(sv_id)
        return self

    # ------------------------------------------------------------------ stage 2
    def reference_part(self):
        """Sorted breakpoints -> reference nodes, reference links and their allele-0 entries
        (construct-graph.py:293-378)."""
        for chrom, clen in self.chrom_len.items():
            bks = sorted(b for b in self._bk_set[chrom] if b < clen - 1)
            self.bkpts[chrom] = bks
            edges = [0] + bks + [clen]
            nodes = [(edges[i] + 1, edges[i + 1]) for i in range(len(edges) - 1)]
            self.nodes[chrom] = nodes
            self.by_end[chrom] = {e: i for i, (_, e) in enumerate(nodes)}
            self.by_start[chrom] = {s: i for i, (s, _) in enumerate(nodes)}
            for i, b in enumerate(bks):
                key = f"{chrom}:{nodes[i][0]}-{nodes[i][1]}@+@{chrom}:{nodes[i + 1][0]}-{nodes[i + 1][1]}@+"
                self.link_sv[key] = [(f"{chrom}:{sv}", 0) for sv in self._bk_sv[chrom][b]]
        return self

    # ------------------------------------------------------------------ stage 3
    def _add_alt(self, chrom, sv_id, link):
        key = "@".join(link)
        self.link_sv.setdefault(key, []).append((f"{chrom}:{sv_id}", 1))
        self.sv_alt_links.setdefault((chrom, sv_id), []).append(link)
        self.alt_lines.append("L\t{}\t{}\t{}\t{}\t0M\n".format(*link))

    def _flanks(self, chrom, left_end, right_start):
        """Names of the node ending at ``left_end`` and of the node starting at ``right_start``: the two
        dictionary probes that replace the reference's scan over every node of the chromosome
        (construct-graph.py:406-413, :438-445).  That scan tests ``stop == pos`` first and the start only
        in the ``elif``, so one node can never be both; it splits the node name at ':' and stops with a
        ValueError when the chromosome name has one of its own."""
        if ":" in chrom:
            raise ValueError(f"node names of '{chrom}' cannot be split at ':' (the reference stops here too)")
        li = self.by_end[chrom].get(left_end)
        ri = self.by_start[chrom].get(right_start)
        if ri is not None and ri == li:
            ri = None
        if li is None or ri is None:
            raise TypeError("sequence item 1: expected str instance, NoneType found"
                            f" (no node next to the breakpoint of an SV on '{chrom}'; the reference stops here too)")
        return self.node_name(chrom, li), self.node_name(chrom, ri)

    def _bnd_node(self, chrom, pos, by_start):
        """find_node_by_start / find_node_by_end (construct-graph.py:584-604) as one dictionary probe."""
        if ":" in chrom and self.nodes[chrom]:
            raise ValueError(f"node names of '{chrom}' cannot be split at ':' (the reference stops here too)")
        idx = (self.by_start if by_start else self.by_end)[chrom].get(pos)
        if idx is None:
            self._warn(f"Warning: looked for nonexistant node {'starting' if by_start else 'ending'} at {pos} on {chrom}")
            return None
        return self.node_name(chrom, idx)

    def alt_part(self):
        """Alt nodes and links of every kept SV, in FASTA order of chromosomes and VCF order within
        (construct-graph.py:383-547)."""
        for chrom, lst in self.svs.items():
            for sv_id in lst:
                kind = sv_id.split("-")[0]
                if kind == "INS":
                    pos, cnt = sv_id.split("-")[1:]
                    pos = int(pos)
                elif kind != "BND":
                    pos, end = sv_id.split("-")[1:]
                    pos, end = int(pos), int(end)
                if kind == "DEL":
                    ln, rn = self._flanks(chrom, pos, end + 1)
                    self._add_alt(chrom, sv_id, (ln, "+", rn, "+"))
                elif kind == "INS":
                    ins = f"{chrom}:{pos + 1}.{cnt}"
                    self.alt_nodes[ins] = self.ins_seq[sv_id]
                    self.sv_ins_node[(chrom, sv_id)] = ins
                    self.alt_lines.append("S\t{}\t{}\n".format(ins, self.ins_seq[sv_id]))
                    ln, rn = self._flanks(chrom, pos, pos + 1)
                    self._add_alt(chrom, sv_id, (ln, "+", ins, "+"))
                    self._add_alt(chrom, sv_id, (ins, "+", rn, "+"))
                elif kind == "INV":
                    if ":" in chrom:
                        raise ValueError(f"node names of '{chrom}' cannot be split at ':' (the reference stops here too)")
                    li = self.by_end[chrom].get(pos)
                    ri = self.by_start[chrom].get(end + 1)
                    if ri is not None and ri == li:
                        ri = None
                    # the reference's elif chain: a node already taken as left/right
                    # flank is not considered as an inner node
                    lin = self.by_start[chrom].get(pos + 1)
                    rin = self.by_end[chrom].get(end)
                    if lin is not None and lin in (li, ri):
                        lin = None
                    if rin is not None and rin in (li, ri):
                        rin = None
                    if None in (li, ri, lin, rin):
                        continue
                    name = self.node_name
                    self._add_alt(chrom, sv_id, (name(chrom, li), "+", name(chrom, rin), "-"))
                    self._add_alt(chrom, sv_id, (name(chrom, lin), "-", name(chrom, ri), "+"))
                elif kind == "BND":
                    left, right = bnd_coords(chrom, sv_id)
                    ln = self._bnd_node(left[0], left[1], left[2] == "-")
                    rn = self._bnd_node(right[0], right[1], right[2] == "+")
                    if ln is None or rn is None:
                        self._warn(f"Warning: no alternative link defined for {sv_id} (ID: {self.vcf_id[sv_id]})")
                        continue
                    if left[2] == "-":
                        link = (ln, "-", rn, "+")
                    elif right[2] == "-":
                        link = (ln, "+", rn, "-")
                    else:
                        link = (ln, "+", rn, "+")
                    self._add_alt(chrom, sv_id, link)
        return self


def build_graph(chrom_len, vcf_rows, chrom_seq=None):
    """``chrom_len``: ordered chrom -> length (FASTA order); ``vcf_rows``:
    iterable of VCF lines (str, tab separated, no newline needed; '#' lines are skipped)."""
    return Graph(chrom_len, chrom_seq).parse(vcf_rows).reference_part().alt_part()


def load_fasta(path):
    """chrom -> upper-cased sequence, as construct-graph.py:79-92 reads it (text mode, so CR LF and a
    lone CR end lines; the name is the first word of the header; a header followed by no sequence is
    dropped unless it is the last one), without a Python-level loop over the lines."""
    import re
    from . import gzio
    with open(path, "rb") as fh:
        head = fh.read(2)
    if gzio.is_gzip(head):                          # extension (row N4): gzip / bgzip FASTA
        import io
        text = io.TextIOWrapper(io.BytesIO(gzio.read_bytes(path).tobytes())).read()
    else:
        with open(path, "r") as fh:
            text = fh.read()
    heads = [m.start() for m in re.finditer(r"^>", text, re.M)]
    if not heads or text[:heads[0]].replace("\n", "") != "":
        # sequence lines in front of the first header, or no header at all
        raise UnboundLocalError("cannot access local variable 'header' where it is not associated with a value")
    seqs = OrderedDict()
    pending = None
    for i, h in enumerate(heads):
        eol = text.find("\n", h)
        if eol < 0:
            eol = len(text)
        name = text[h + 1:eol].split()[0]          # IndexError on an empty header, as in the reference
        nxt = heads[i + 1] if i + 1 < len(heads) else len(text)
        body = text[eol + 1:nxt].replace("\n", "").upper()
        if body != "" or i + 1 == len(heads):
            seqs[name] = body
    return seqs


def construct_main(argv=None):
    """Drop-in ``construct-graph.py``: same flags (construct-graph.py:26-65), same three output files
    with the same bytes, same warnings on stdout, exit status 1 where the reference stops."""
    import argparse
    import sys
    ap = argparse.ArgumentParser()
    ap.add_argument("-v", "--vcf", metavar="<inputVCF>", type=str, nargs=1, required=True)
    ap.add_argument("-r", "--ref", metavar="<referenceGenome>", type=str, nargs=1, required=True)
    ap.add_argument("-o", "--output", metavar="<outputFile", type=str)
    args = ap.parse_args(argv)
    if args.output:
        out_gfa, prefix = args.output, args.output.replace(".gfa", "_")
    else:
        out_gfa, prefix = "variation_graph.gfa", ""

    def stop(exc):
        sys.stdout.flush()
        sys.exit(f"construct-graph: {type(exc).__name__}: {exc}")

    stops = (ValueError, IndexError, KeyError, TypeError, UnboundLocalError, OSError)
    try:
        seqs = load_fasta(args.ref[0])
        g = Graph(OrderedDict((c, len(s)) for c, s in seqs.items()), seqs, warn=print)
        from . import gzio
        g.parse(gzio.read_text_lines(args.vcf[0]))          # open(vcf).readlines(); a gzip / bgzip VCF is inflated (row N4)
    except stops as exc:
        stop(exc)
    with open(f"{prefix}ignored_svs.txt", "w") as fh:
        fh.write(g.ignored_text())
    g.reference_part()
    failed = None
    try:
        g.alt_part()
    except stops as exc:
        failed = exc
    with open(out_gfa, "w") as fh:              # up to the record the reference stops on, like its own file
        g.write_gfa(fh)
    if failed is not None:
        stop(failed)
    with open(f"{prefix}svs_edges.json", "w") as fh:
        g.write_edges_json(fh)
    return 0
