#!/usr/bin/env python3
"""Regenerates tests/golden/fuzz_graph.json.gz: small reference genomes and SV catalogues, sound and
damaged, each run through the UNMODIFIED reference graph constructor
(`python3 /root/reference/construct-graph.py -v x.vcf -r x.fa -o x.gfa`, one process per case).
Stored per case: the inputs, the exit status, stdout, the last line of stderr and the SHA-256 of every
file the run left behind (x.gfa — complete or cut short where the reference stopped —,
x_svs_edges.json, x_ignored_svs.txt).

    python tests/golden/make_graph_fuzz.py      # build container only (needs /root/reference)
"""
import gzip
import hashlib
import json
import os
import random
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT_FILES = ("x.gfa", "x_svs_edges.json", "x_ignored_svs.txt")


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def fastas(seed=7):
    """(text, {chrom: length} or None when the reference cannot load it)."""
    rng = random.Random(seed)
    out = []

    def wrap(seq, w):
        return "".join(seq[i:i + w] + "\n" for i in range(0, len(seq), w)) if w else seq + "\n"

    for k in range(10):
        names = ["chr1", "chr2", "chrX", "scaf_7", "HLA-A*01", "1"][:rng.choice((1, 2, 3, 3, 4))]
        rng.shuffle(names)
        text, lens = "", {}
        for nm in names:
            n = rng.choice((60, 150, 400, 700))
            seq = rand_seq(rng, n, "ACGTacgtNn" if k % 3 == 0 else "ACGT")
            text += ">" + nm + rng.choice(("", " description here", "\tx=1")) + "\n" + wrap(seq, rng.choice((0, 60, 70, 13)))
            lens[nm] = n
        if k == 4:
            text = text.replace("\n", "\r\n")
        if k == 5:
            text = text.rstrip("\n")
        if k == 6:
            text = "\n\n" + text.replace("\n>", "\n\n>")
        out.append((text, lens))
    s = rand_seq(rng, 120)
    out.append((">a\n>b\n" + s + "\n>c\n", {"b": 120, "c": 0}))             # header without sequence
    out.append((">a\n" + s + "\n>a\n" + s[:50] + "\n>b\n" + s + "\n", {"a": 50, "b": 120}))   # duplicate name
    out.append((">a:1\n" + s + "\n>b\n" + s + "\n", {"a:1": 120, "b": 120}))  # ':' in a name
    out.append((s + "\n>a\n" + s + "\n", None))                             # sequence before any header
    out.append((">\n" + s + "\n", None))                                    # empty header
    out.append(("", None))
    out.append((">a " + "\n" + s + "\n" + "AC>GT\n", {"a": 125}))           # '>' inside a line
    return out


def record(rng, lens, vid):
    """One VCF body line on the given chromosomes; mostly sound, positions biased to the edges."""
    chroms = list(lens)
    c = rng.choice(chroms)
    n = max(lens[c], 4)

    def where():
        r = rng.random()
        if r < 0.08:
            return rng.choice((0, 1, 2, 3))
        if r < 0.16:
            return n - rng.choice((0, 1, 2, 3, 4))
        return rng.randrange(2, n)
    pos = where()
    kind = rng.choice(("DEL", "DEL", "INS", "INS", "INV", "BND", "BND", "DUP"))
    ref, alt = "N", "<" + kind + ">"
    if kind in ("DEL", "INV", "DUP"):
        end = min(n + 2, pos + rng.choice((1, 5, 30, 80, 200))) if rng.random() < 0.93 else pos - rng.choice((0, 1, 7))
        info = f"SVTYPE={kind};END={end};SVLEN={end - pos}"
        if rng.random() < 0.3:
            info = f"END={end};SVTYPE={kind}" if rng.random() < 0.5 else f"SVLEN=5;SVTYPE={kind};AF=1;END={end}"
    elif kind == "INS":
        seq = rand_seq(rng, rng.choice((1, 8, 40)), "ACGTacgt")
        r = rng.random()
        info = f"SVTYPE=INS;END={pos};SVLEN={len(seq)}"
        if r < 0.6:
            alt = seq
        elif r < 0.75:
            alt, info = "<INS>", f"SVTYPE=INS;SEQ={seq};SVLEN={len(seq)}"
        elif r < 0.82:
            alt, info = "<INS>", f"SVTYPE=INS;LEFT_SVINSSEQ={seq};SEQ={seq}"
        elif r < 0.9:
            alt = "<INS>"
        else:
            ref, alt = "NAC", seq
    else:
        c2 = rng.choice(chroms)
        p2 = rng.choice((0, 1, 2, max(lens[c2], 4) - 1, max(lens[c2], 4), rng.randrange(2, max(lens[c2], 4))))
        t = rng.choice(("N", "A", "NN"))
        mate = f"{c2}:{p2}"
        alt = rng.choice((f"{t}[{mate}[", f"{t}]{mate}]", f"]{mate}]{t}", f"[{mate}[{t}"))
        info = "SVTYPE=BND"
        if rng.random() < 0.06:
            alt = rng.choice(("N[", f"N[{c2}[", "<BND>", f"N[chr9:5[", f"[{mate}[", f"N{mate}", f"N[{c2}:x["))
    tail = rng.choice(("", "", "\tGT\t0/1", "\tGT"))
    return f"{c}\t{pos}\tsv{vid}\t{ref}\t{alt}\t.\tPASS\t{info}{tail}"


def damage(rng, line):
    cols = line.split("\t")
    if len(cols) < 8:                # already damaged
        return line
    k = rng.randrange(9)
    if k == 0:
        return "\t".join(cols[:rng.randrange(1, 8)])
    if k == 1:
        return ""
    if k == 2:
        cols[1] = rng.choice(("x", "", "-5", "+7", " 9", "1e3", "007"))
    elif k == 3:
        cols[7] = rng.choice(("", ".", "SVTYPE=DEL", "END=5", "SVTYPE=INV;SVLEN=3", "XSVTYPE=DEL;END=9", "SVTYPE=DEL;END=", "SVTYPE=DEL;END=z"))
    elif k == 4:
        cols[0] = rng.choice(("chrQ", "", "chr1 "))
    elif k == 5:
        return line + rng.choice((" ", "\t", "\t\t", "\r"))
    elif k == 6:
        cols[7] = cols[7] + "\t"
    elif k == 7:
        return "#" + line
    else:
        pos = rng.randrange(len(line))
        return line[:pos] + rng.choice("\t;=[]:<>-0 ") + line[pos + 1:]
    return "\t".join(cols)


def catalogues(fas, n_cases, seed=20261017):
    rng = random.Random(seed)
    header = "##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
    cases = []
    loadable = [i for i, (_, lens) in enumerate(fas) if lens]
    for ci in range(n_cases):
        fa = rng.choice(loadable) if rng.random() < 0.97 else rng.randrange(len(fas))
        lens = fas[fa][1] or {"a": 120}
        n = rng.choice((0, 1, 2, 3, 5, 8, 14, 25))
        recs = [record(rng, lens, i + 1) for i in range(n)]
        if recs and rng.random() < 0.35:                    # repeated positions: same-POS INS, shared breakpoints, duplicates
            for _ in range(rng.choice((1, 2, 4))):
                src = rng.choice(recs).split("\t")
                new = record(rng, lens, len(recs) + 1).split("\t")
                if rng.random() < 0.5:
                    new[0], new[1] = src[0], src[1]
                    recs.append("\t".join(new))
                else:
                    recs.append("\t".join(src))
        if recs and rng.random() < 0.3:
            for _ in range(rng.choice((1, 1, 2))):
                j = rng.randrange(len(recs))
                recs[j] = damage(rng, recs[j])
        eol = "\r\n" if rng.random() < 0.05 else "\n"
        body = "".join(r + eol for r in recs)
        if recs and rng.random() < 0.1:
            body = body[:-len(eol)]
        cases.append({"fa": fa, "vcf": (header if rng.random() < 0.9 else "") + body})
    return cases


def run_reference(tmp, fa_text, vcf_text):
    for f in OUT_FILES:
        if os.path.exists(os.path.join(tmp, f)):
            os.remove(os.path.join(tmp, f))
    with open(os.path.join(tmp, "x.fa"), "w", newline="") as fh:
        fh.write(fa_text)
    with open(os.path.join(tmp, "x.vcf"), "w", newline="") as fh:
        fh.write(vcf_text)
    p = subprocess.run([sys.executable, os.path.join(REF, "construct-graph.py"), "-v", "x.vcf", "-r", "x.fa", "-o", "x.gfa"],
                       cwd=tmp, capture_output=True, text=True)
    files = {}
    for f in OUT_FILES:
        path = os.path.join(tmp, f)
        files[f] = hashlib.sha256(open(path, "rb").read()).hexdigest() if os.path.exists(path) else None
    err = p.stderr.strip().splitlines()
    return {"rc": p.returncode, "stdout": p.stdout, "stderr_last": err[-1] if err else "", "files": files}


def main():
    fas = fastas()
    cases = catalogues(fas, 1500)
    with tempfile.TemporaryDirectory() as tmp:
        for c in cases:
            c.update(run_reference(tmp, fas[c["fa"]][0], c["vcf"]))
    n_ok = sum(c["rc"] == 0 for c in cases)
    print(f"{len(cases)} cases: {n_ok} complete, {len(cases) - n_ok} stopped by the reference", file=sys.stderr)
    kinds = {}
    for c in cases:
        if c["rc"]:
            k = c["stderr_last"].split(":")[0]
            kinds[k] = kinds.get(k, 0) + 1
    print(kinds, file=sys.stderr)
    blob = json.dumps({"fastas": [f for f, _ in fas], "cases": cases}, separators=(",", ":")).encode()
    with open(os.path.join(HERE, "fuzz_graph.json.gz"), "wb") as fh:
        fh.write(gzip.compress(blob, 9, mtime=0))
    print(f"fuzz_graph.json.gz: {os.path.getsize(os.path.join(HERE, 'fuzz_graph.json.gz'))} bytes", file=sys.stderr)


if __name__ == "__main__":
    main()
