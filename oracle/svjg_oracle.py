"""CPU oracle for the SVJedi-graph post-mapping hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement of the reference algorithm; it is the
checker the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference``
legs may import it.  Nothing under ``svjedi-graph_b200/`` imports it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified
reference scripts (``/root/reference/filter-alignments.py``,
``predict-genotype.py``, ``construct-graph.py``) in the build container and
commits their outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every function below against those outputs and against the 40
known-answer rows of the reference's ``test-dir/expected_genotype.vcf``.

Each function cites the reference lines it restates (paths relative to
``/root/reference``).
"""
from __future__ import annotations

import json
import math
from fractions import Fraction

D_OVER_DEFAULT = 100          # filter-alignments.py:56
ERR_DEFAULT = 0.00005         # predict-genotype.py:61
MIN_SUPPORT_DEFAULT = 3       # predict-genotype.py:46


class OracleInputError(Exception):
    """Raised where the reference would die with a traceback (exit code 1)."""


# --------------------------------------------------------------------------
# table loading
# --------------------------------------------------------------------------
def load_link_table(svs_edges_path):
    """link key -> [(sv_id, allele), ...]   (filter-alignments.py:95-98)"""
    with open(svs_edges_path, "r") as fh:
        return json.load(fh)


def load_alt_node_len(gfa_path):
    """alt-node name -> len(sequence)   (filter-alignments.py:103-113).

    Only ``S`` lines whose name's last ':'-piece contains '.' are kept; a
    repeated name keeps the last length (dict assignment)."""
    out = {}
    with open(gfa_path, "r") as fh:
        for line in fh:
            if not line.startswith("S"):
                continue
            cols = line.split("\t")
            if "." in cols[1].split(":")[-1]:
                out[cols[1]] = len(line.rstrip().split("\t")[2])
    return out


# --------------------------------------------------------------------------
# filter  (filter-alignments.py:123-166, 184-219, 258-273, 343-373)
# --------------------------------------------------------------------------
def _node_len(name, alt_len):
    """filter-alignments.py:343-349 (+ :328-342)."""
    last = name.rsplit(":", 1)[-1]
    if "." in last:
        try:
            return alt_len[name]
        except KeyError as exc:
            raise OracleInputError(f"alt node {name!r} missing from GFA") from exc
    parts = last.split("-")
    try:
        return int(parts[1]) - int(parts[0]) + 1
    except (IndexError, ValueError) as exc:
        raise OracleInputError(f"bad node name {name!r}") from exc


def _tokens(path):
    """filter-alignments.py:351-373."""
    if path == "":
        raise OracleInputError("empty path column")
    if path[0] in "<>":
        toks, cur = [], []
        for ch in path:
            if ch == "<" or ch == ">":
                if cur:
                    toks.append("".join(cur))
                    cur = []
            else:
                cur.append(ch)
        if cur:
            toks.append("".join(cur))
        return toks
    return [piece[:-1] for piece in path.split(",") if piece]


def _strand(path, tok):
    """filter-alignments.py:206 — the character in front of the FIRST substring
    occurrence of the token decides ('>' -> '+', anything else -> '-')."""
    if tok == "":
        raise OracleInputError("empty token")
    j = path.find(tok)
    if j <= 0:
        raise OracleInputError("token at path start")
    return "+" if path[j - 1] == ">" else "-"


_FLIP = {"+": "-", "-": "+"}


def parse_record(line):
    """filter-alignments.py:126, 184-198 on ``line.rstrip()``.
    Returns (path, Tlen, Ts, Te).  Raises where the reference raises."""
    cols = line.rstrip().split("\t")
    if len(cols) < 12:
        raise OracleInputError("fewer than 12 columns")
    try:
        for i in (1, 2, 3, 6, 7, 8, 9, 10, 11):
            int(cols[i])
        tlen, ts, te = int(cols[6]), int(cols[7]), int(cols[8])
        if "id:f:" in line.rstrip():
            float(line.rstrip().split("id:f:")[-1].split("\t")[0])
        elif int(cols[10]) == 0:
            raise OracleInputError("Alen == 0 without id:f:")
    except ValueError as exc:
        raise OracleInputError(str(exc)) from exc
    return cols[5], tlen, ts, te


def text_mode_lines(text):
    """The lines `for line in open(path)` yields for a file with this content (filter-alignments.py
    :123-124 reads the GAF in text mode): universal newlines -- "\r\n" and a lone "\r" end a line
    just like "\n", and all of them come out as "\n"."""
    import io
    return list(io.StringIO(text, newline=None))


def record_hits(line, d_link_sv, alt_len, d_over=D_OVER_DEFAULT):
    """All (sv_id, allele) appends one GAF line causes, in reference order
    (link, then fwd/rev key, then entry) — filter-alignments.py:126-166."""
    path, tlen, ts, te = parse_record(line)
    toks = _tokens(path)
    n = len(toks)
    if n < 2:
        return []
    strands = [_strand(path, t) for t in toks]
    first = {}
    for i, t in enumerate(toks):
        first.setdefault(t, i)
    out = []
    lens = None
    tail_clip = tlen - te - 1
    for i in range(1, n):
        a, b = toks[i - 1], toks[i]
        fwd = "@".join((a, strands[i - 1], b, strands[i]))
        rev = "@".join((b, _FLIP[strands[i]], a, _FLIP[strands[i - 1]]))
        keys = [k for k in (fwd, rev) if k in d_link_sv]
        for k in keys:
            for sv_id, allele in d_link_sv[k]:
                # the reference evaluates the overlap test here, once per entry
                # (:156); it only depends on (link, record)
                il, ir = first[a], first[b]
                left = 0
                for j in range(il + 1):
                    left += _node_len(toks[j], alt_len)
                right = 0
                for j in range(ir, n):
                    right += _node_len(toks[j], alt_len)
                if left - ts >= d_over and right - tail_clip >= d_over:
                    if ":" not in sv_id:
                        raise OracleInputError("sv id without ':'")
                    if allele not in (0, 1, -1, -2) or isinstance(allele, float):
                        raise OracleInputError("allele index out of range")
                    out.append((sv_id, allele))
    return out


def kept_text(line):
    """filter-alignments.py:166 — text stored per hit: everything before the
    first 'cg:Z:' or, without one, the whole raw line including its newline."""
    j = line.find("cg:Z:")
    return line if j < 0 else line[:j]


def record_identity(line):
    """``aln["Aid"]`` as filter-alignments.py:193-196 computes it (and never uses it): float() of the text behind
    the last "id:f:" up to the next tab where the stripped line holds that tag, else Am / Alen."""
    s = line.rstrip()
    if "id:f:" in s:
        return float(s.split("id:f:")[-1].split("\t")[0])
    cols = s.split("\t")
    return int(cols[9]) / int(cols[10])


def filter_alignments(lines, d_link_sv, alt_len, d_over=D_OVER_DEFAULT, min_identity=None):
    """sv_id -> [[ref texts], [alt texts]]  (filter-alignments.py:119-166).  ``d_over``: the threshold of :56
    (what -O was meant to set).  ``min_identity``: the extension of the new front-end, off by default -- an
    alignment whose Aid (:193-196) is below it is appended nowhere (the gate predict-genotype.py:222 has
    commented out, applied where the identity is parsed)."""
    out = {}
    for line in lines:
        hits = record_hits(line, d_link_sv, alt_len, d_over)
        if not hits:
            continue
        if min_identity is not None and record_identity(line) < min_identity:
            continue
        text = kept_text(line)
        for sv_id, allele in hits:
            slot = out.get(sv_id)
            if slot is None:
                slot = out[sv_id] = [[], []]
            slot[allele].append(text)
    return out


def dumps_informative(d):
    """filter-alignments.py:174-175."""
    return json.dumps(d, sort_keys=True, indent=4)


def hit_counts(d):
    """sv_id -> (n_ref, n_alt): what predict-genotype.py:219-226 derives."""
    return {k: (len(v[0]), len(v[1])) for k, v in d.items()}


# --------------------------------------------------------------------------
# genotype  (predict-genotype.py:281-346), exact rational restatement
# --------------------------------------------------------------------------
def _py_num_str(twice, was_halved):
    """str() of a count the reference holds either as int or as round(n/2, 1)."""
    if not was_halved:
        return str(twice // 2)
    return f"{twice // 2}.{'5' if twice & 1 else '0'}"


def _round_half_even_half_units(twice):
    """int(round(c, 0)) for c = twice/2  (predict-genotype.py:291-292)."""
    q, r = divmod(twice, 2)
    if r == 0:
        return q
    return q + (q & 1)


def genotype_counts(n0, n1, svtype, min_support=MIN_SUPPORT_DEFAULT, e=ERR_DEFAULT):
    """Returns (GT, DP_str, AD_str, [PL0, PL1, PL2] as str) for raw hit counts.

    predict-genotype.py:281-338 with all Decimal arithmetic replaced by exact
    rationals of the same binary64 products (the 28-digit context never rounds
    a value that matters; tests/test_oracle_golden.py checks this against the
    unmodified reference on the committed random vectors)."""
    t1, t2 = 2 * n0, 2 * n1            # counts in half units
    h1 = h2 = False
    if svtype == "DEL" and n0 > 0:     # :327-338 (halves the 2-breakpoint allele)
        t1, h1 = n0, True
    elif svtype == "INS" and n1 > 0:
        t2, h2 = n1, True
    c1 = t1 / 2 if h1 else t1 // 2
    c2 = t2 / 2 if h2 else t2 // 2
    a, b, h = math.log10(1 - e), math.log10(e), math.log10(1 / 2)
    l0 = Fraction(c1 * a) + Fraction(c2 * b)          # :295
    l1 = Fraction((c1 + c2) * h)                      # :296
    l2 = Fraction(c2 * a) + Fraction(c1 * b)          # :297
    liks = (l0, l1, l2)
    best = max(liks)
    winners = [i for i, v in enumerate(liks) if v == best]
    gt = ("0/0", "0/1", "1/1")[winners[0]] if len(winners) == 1 else "./."
    if t1 + t2 < 2 * min_support:                     # :310 (on normalised counts)
        gt = "./."
    r1 = _round_half_even_half_units(t1)
    r2 = _round_half_even_half_units(t2)
    comb = Fraction(math.log10(math.comb(r1 + r2, r1)))   # :313
    pl = []
    for v in liks:
        q = -10 * (v + comb)
        pl.append(str(-((-q.numerator) // q.denominator) if q < 0 else q.numerator // q.denominator))
    halved_any = h1 or h2
    dp = _py_num_str(t1 + t2, halved_any)             # str(round(sum, 3)) :265
    ad = _py_num_str(t1, h1) + "," + _py_num_str(t2, h2)  # :248
    return gt, dp, ad, pl


# --------------------------------------------------------------------------
# VCF side  (predict-genotype.py:89-279)
# --------------------------------------------------------------------------
FORMAT_HEADER = (
    '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
    '##FORMAT=<ID=DP,Number=1,Type=Float,Description="Total number of informative read alignments across all alleles (after normalization for unbalanced SVs)">\n'
    '##FORMAT=<ID=AD,Number=2,Type=Float,Description="Number of informative read alignments supporting each allele (after normalization by breakpoint number for unbalanced SVs)">\n'
    '##FORMAT=<ID=PL,Number=3,Type=Integer,Description="Phred-scaled likelihood for each genotype">\n'
    "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n"
)


def _info_value(info, label):
    """predict-genotype.py:77-87."""
    pieces = info.split(";")
    if pieces[0].startswith(label + "="):
        return info.split(label + "=")[1].split(";")[0]
    try:
        if pieces[-1].startswith(label + "="):
            return info.split(";" + label + "=")[1]
        return info.split(";" + label + "=")[1].split(";")[0]
    except IndexError as exc:
        raise OracleInputError(f"INFO lacks {label}=") from exc


def vcf_sv_key(chrom, pos, alt, info, ins_seen):
    """(svtype, key, length) for one VCF body line — predict-genotype.py:123-211.
    ``ins_seen`` is the running POS -> count dict (:151-155)."""
    if "SVTYPE" in info:
        tail = info.split("SVTYPE=")
        if len(tail) < 2:
            raise OracleInputError("SVTYPE without '='")
        svtype = tail[1] if info.split(";")[-1].startswith("SVTYPE=") else tail[1].split(";")[0]
    else:
        svtype = ""
    end = None
    if svtype != "BND" and svtype != "INS":
        end = _info_value(info, "END")
    if svtype == "DEL" or svtype == "INV":
        try:
            length = int(end) - int(pos)
        except ValueError as exc:
            raise OracleInputError(str(exc)) from exc
        return svtype, f"{chrom}:{svtype}-{pos}-{end}", length
    if svtype == "INS":
        ins_seen[pos] = ins_seen.get(pos, 0) + 1
        return svtype, f"{chrom}:INS-{pos}-{ins_seen[pos]}", len(alt)
    if svtype == "BND":
        for br in "[]":
            if br in alt:
                pieces = [p for p in alt.split(br) if p]
                try:
                    if ":" in pieces[1]:
                        return svtype, f"{chrom}:BND-{pos}{br}{pieces[1]}{br}", 50
                    return svtype, f"{chrom}:BND-{br}{pieces[0]}{br}{pos}", 50
                except IndexError as exc:
                    raise OracleInputError("bad BND ALT") from exc
        return svtype, "wrong_format", 50
    return svtype, "unsupported_type", None


def genotype_vcf(counts, vcf_lines, min_support=MIN_SUPPORT_DEFAULT, e=ERR_DEFAULT):
    """counts: sv_id -> (n_ref, n_alt).  Returns (output text, genotyped count).
    predict-genotype.py:100-275."""
    out = []
    ins_seen = {}
    genotyped = 0
    for line in vcf_lines:
        if line.startswith("##FORMAT"):
            continue
        if line.startswith("##"):
            out.append(line)
            continue
        if line.startswith("#C"):
            out.append(FORMAT_HEADER)
            continue
        cols = line.rstrip("\n").split("\t")
        if len(cols) < 8:
            raise OracleInputError("VCF line with fewer than 8 columns")
        svtype, key, length = vcf_sv_key(cols[0], cols[1], cols[4], cols[7], ins_seen)
        if svtype in ("DEL", "INS", "INV", "BND") and key in counts and abs(length) >= 50:
            n0, n1 = counts[key]
            gt, dp, ad, pl = genotype_counts(n0, n1, svtype, min_support, e)
            genotyped += 1
        else:
            gt, dp, ad, pl = "./.", "0", "0,0", [".", ".", "."]
        head = line.rstrip("\n") if len(line.split("\t")) <= 8 else "\t".join(line.split("\t")[:8])
        out.append(f"{head}\tGT:DP:AD:PL\t{gt}:{dp}:{ad}:{','.join(pl)}\n")
    return "".join(out), genotyped
