#!/usr/bin/env python3
"""Regenerates the damaged-input fixtures, each case run through the UNMODIFIED reference scripts
(imported from /root/reference and their main() called in-process):

    python tests/golden/make_fuzz.py lines   # fuzz_lines.json.gz  2 500 damaged GAF lines on the c1 graph (+ a CR LF file)
    python tests/golden/make_fuzz.py vcf     # fuzz_vcf.json.gz    800 VCFs with a damaged body line (predict-genotype.py)
    python tests/golden/make_fuzz.py edges   # fuzz_edges.json     300 svs_edges.json with damaged entries
    python tests/golden/make_fuzz.py gfa     # fuzz_gfa.json.gz    1 200 damaged GFAs: the reference's alt_node_len dictionary
    python tests/golden/make_fuzz.py json    # fuzz_json.json      600 damaged informative_aln.json (predict-genotype.py)

Stored per case: the input (or, where it is regenerated from the seed, a hash of it), whether the
reference raised (exit status 1), and otherwise its output (list lengths, text, or a SHA-256 of it).
Build container only (needs /root/reference); the tests read the committed fixtures.
"""
import gzip
import hashlib
import importlib.util
import io
import json
import os
import random
import sys
import tempfile
from contextlib import redirect_stderr, redirect_stdout

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def read_golden(name):
    p = os.path.join(HERE, name)
    if name.endswith(".gz"):
        return gzip.open(p, "rt").read()
    return open(p).read()


def mutated_lines(n, seed=20261017):
    """Deterministic damage: byte edits, column edits, truncation, revisited nodes.  Avoids the two
    documented deviations of the CUDA path ('_' in integers, more than 18 digits in Tlen/Ts/Te)."""
    base = [l for l in read_golden("c1.gaf.gz").splitlines(True) if l.count(">") + l.count("<") >= 2][:300]
    rng = random.Random(seed)
    alphabet = "\t\t\t 09a><:-.,+\r\x0b;|"
    out = []
    while len(out) < n:
        line = rng.choice(base).rstrip("\n")
        for _ in range(rng.choice((1, 1, 2, 3))):
            k = rng.randrange(6)
            pos = rng.randrange(len(line) + 1)
            if k == 0 and pos < len(line):
                line = line[:pos] + rng.choice(alphabet) + line[pos + 1:]
            elif k == 1 and pos < len(line):
                line = line[:pos] + line[pos + 1:]
            elif k == 2:
                line = line[:pos] + rng.choice(alphabet) + line[pos:]
            elif k == 3:
                cols = line.split("\t")
                j = rng.randrange(len(cols))
                cols[j] = rng.choice(("", "0", "00", "+5", " 7 ", "12x", cols[j] + cols[j], cols[j][::-1]))
                line = "\t".join(cols)
            elif k == 4:
                line = line[:pos]
            else:
                cols = line.split("\t")
                if len(cols) > 5:
                    toks = [x for x in cols[5].replace("<", ">").split(">") if x]
                    if toks:
                        cols[5] = cols[5] + rng.choice(">< ") + rng.choice(toks)      # revisit / odd tail
                        line = "\t".join(cols)
        if "_" in line or "\n" in line or not line:
            continue
        cols = line.split("\t")
        if any(len(c.strip().lstrip("+-").lstrip("0")) > 18 for c in cols[6:9]):
            continue
        out.append(line + rng.choice(("\n", "\n", "")))
    return out


def main():
    spec = importlib.util.spec_from_file_location("ref_filter", os.path.join(REF, "filter-alignments.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    lines = mutated_lines(2500)
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "p.gfa"), "w").write(read_golden("c1.gfa.gz"))
        open(os.path.join(tmp, "p_svs_edges.json"), "w").write(read_golden("c1_svs_edges.json"))
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            for line in lines:
                with open("p.gaf", "w", newline="") as fh:
                    fh.write(line)
                if os.path.exists("p_informative_aln.json"):
                    os.remove("p_informative_aln.json")
                try:
                    sys.argv = ["filter-alignments.py", "-a", "p.gaf", "-g", "p.gfa", "-p", "p"]   # main() parses sys.argv
                    with redirect_stdout(io.StringIO()), redirect_stderr(io.StringIO()):
                        ref.main(sys.argv[1:])
                    d = json.load(open("p_informative_aln.json"))
                    cases.append({"line": line, "rc": 0, "counts": {k: [len(v[0]), len(v[1])] for k, v in d.items()}})
                except BaseException:                                   # any traceback or sys.exit: status 1
                    cases.append({"line": line, "rc": 1})
        finally:
            os.chdir(cwd)
        # a whole file with CR LF line ends, one line with a lone CR in its read name: Python's text mode
        # makes every one of them a line break and stores "\n" in the JSON
        os.chdir(tmp)
        try:
            src = read_golden("c1.gaf.gz").splitlines()[:400]
            src = [l for l in src if "cg:Z:" not in l]
            crlf = "\r\n".join(src) + "\r\n"
            with open("p.gaf", "w", newline="") as fh:
                fh.write(crlf)
            sys.argv = ["filter-alignments.py", "-a", "p.gaf", "-g", "p.gfa", "-p", "p"]
            with redirect_stdout(io.StringIO()), redirect_stderr(io.StringIO()):
                ref.main(sys.argv[1:])
            crlf_json = open("p_informative_aln.json").read()
        finally:
            os.chdir(cwd)
    with gzip.GzipFile(os.path.join(HERE, "fuzz_lines.json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps({"cases": cases, "crlf": {"gaf": crlf, "json": crlf_json}}, ensure_ascii=True).encode())
    print(len(cases), "cases;", sum(c["rc"] for c in cases), "raise;", sum(1 for c in cases if c.get("counts")), "with hits")


def damaged_vcfs(n, seed=20261018):
    """Small VCFs: the header of c1.vcf, a few of its body lines, one of them damaged."""
    text = read_golden("c1.vcf")
    header = [l for l in text.splitlines(True) if l.startswith("#")]
    body = [l for l in text.splitlines(True) if not l.startswith("#")]
    rng = random.Random(seed)
    alphabet = "\t;=:[]09ANt-,. <>"
    out = []
    while len(out) < n:
        pick = [rng.choice(body) for _ in range(rng.choice((2, 3, 4)))]
        j = rng.randrange(len(pick))
        line = pick[j].rstrip("\n")
        for _ in range(rng.choice((1, 1, 2))):
            k = rng.randrange(7)
            pos = rng.randrange(len(line) + 1)
            cols = line.split("\t")
            if k == 0 and pos < len(line):
                line = line[:pos] + rng.choice(alphabet) + line[pos + 1:]
            elif k == 1 and pos < len(line):
                line = line[:pos] + line[pos + 1:]
            elif k == 2:
                line = line[:pos] + rng.choice(alphabet) + line[pos:]
            elif k == 3 and len(cols) > 7:
                info = cols[7].split(";")
                rng.shuffle(info)
                if rng.random() < 0.5 and len(info) > 1:
                    info.pop(rng.randrange(len(info)))
                cols[7] = ";".join(info)
                line = "\t".join(cols)
            elif k == 4:
                line = "\t".join(cols[:rng.randrange(len(cols) + 1)])
            elif k == 5 and len(cols) > 7:
                cols[7] = cols[7].replace("SVTYPE=" + rng.choice(("DEL", "INS", "INV", "BND")), "SVTYPE=" + rng.choice(("DEL", "INS", "INV", "BND", "DUP", "")))
                line = "\t".join(cols)
            elif len(cols) > 4:
                cols[rng.choice((1, 4))] = rng.choice(("", "0", "x", "12", cols[4] * 3, "N[1:500[", "]2:7000]N", "A" * 70))
                line = "\t".join(cols)
        pick[j] = line + "\n"
        hdr = list(header)
        if rng.random() < 0.15:
            hdr.insert(rng.randrange(len(hdr)), rng.choice(("#odd header line\n", "##FORMAT=<ID=XX>\n", "##extra=1\n")))
        out.append("".join(hdr + pick))
    return out


def main_vcf():
    spec = importlib.util.spec_from_file_location("ref_pg", os.path.join(REF, "predict-genotype.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "aln.json"), "w").write(read_golden("c1_informative_aln.json.gz"))
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            for text in damaged_vcfs(800):
                with open("in.vcf", "w", newline="") as fh:
                    fh.write(text)
                if os.path.exists("out.vcf"):
                    os.remove("out.vcf")
                ms = random.Random(len(cases)).choice((3, 3, 1, 10))
                buf = io.StringIO()
                try:
                    sys.argv = ["predict-genotype.py", "-d", "aln.json", "-v", "in.vcf", "-o", "out.vcf", "-ms", str(ms)]
                    with redirect_stdout(buf), redirect_stderr(io.StringIO()):
                        ref.main(sys.argv[1:])
                    cases.append({"vcf": text, "ms": ms, "rc": 0, "out": open("out.vcf").read(), "stdout": buf.getvalue()})
                except BaseException:
                    cases.append({"vcf": text, "ms": ms, "rc": 1})
        finally:
            os.chdir(cwd)
    with gzip.GzipFile(os.path.join(HERE, "fuzz_vcf.json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps(cases, ensure_ascii=True).encode())
    print(len(cases), "vcf cases;", sum(c["rc"] for c in cases), "raise")


def damaged_edges(n, seed=7):
    """svs_edges.json of c1 with a few entries damaged: odd allele values, sv ids without ':', wrong
    shapes.  Deterministic, shared with tests/test_oracle_golden.py."""
    edges0 = json.loads(read_golden("c1_svs_edges.json"))
    rng = random.Random(seed)
    odd_alleles = [0, 1, 2, -1, -2, -3, True, False, None, "1", 1.0, [0]]
    odd_ids = ["noColon", "", "a:b", 5, None, ["x:y"], "1:DEL-1-2"]
    out = []
    for _ in range(n):
        e = json.loads(json.dumps(edges0))
        keys = list(e)
        for _ in range(rng.choice((1, 2, 4))):
            k = rng.choice(keys)
            m = rng.randrange(6)
            if m == 0:
                e[k][0][1] = rng.choice(odd_alleles)
            elif m == 1:
                e[k][0][0] = rng.choice(odd_ids)
            elif m == 2:
                e[k][0] = e[k][0] + [7]
            elif m == 3:
                e[k] = rng.choice((5, "ab", "abc", None, {}, {"x:y": 1}, [], [[]], "x"))
            elif m == 4:
                e[k].append(rng.choice((["9:ZZ-1", 0], ["bad", 1], "ab", 3)))
            else:
                e[k][0] = e[k][0][:1]
        out.append(json.dumps(e))
    return out


def edges_gaf_lines():
    return [l for l in read_golden("c1.gaf.gz").splitlines(True) if "cg:Z:" not in l][:150]


def main_edges():
    spec = importlib.util.spec_from_file_location("ref_filter", os.path.join(REF, "filter-alignments.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            open("p.gfa", "w").write(read_golden("c1.gfa.gz"))
            open("p.gaf", "w").write("".join(edges_gaf_lines()))
            for text in damaged_edges(300):
                open("p_svs_edges.json", "w").write(text)
                if os.path.exists("p_informative_aln.json"):
                    os.remove("p_informative_aln.json")
                try:
                    sys.argv = ["filter-alignments.py", "-a", "p.gaf", "-g", "p.gfa", "-p", "p"]
                    with redirect_stdout(io.StringIO()), redirect_stderr(io.StringIO()):
                        ref.main(sys.argv[1:])
                    js = open("p_informative_aln.json").read()
                    cases.append({"rc": 0, "sha256": hashlib.sha256(js.encode()).hexdigest()})
                except BaseException:
                    cases.append({"rc": 1})
        finally:
            os.chdir(cwd)
    with open(os.path.join(HERE, "fuzz_edges.json"), "w") as fh:
        json.dump(cases, fh)
    print(len(cases), "edges cases;", sum(c["rc"] for c in cases), "raise")

    main_vcf()
    main_edges()


def damaged_gfas(n, seed=11):
    """Small GFAs cut from c1.gfa (sequences of reference nodes shortened), a few lines damaged:
    byte edits, columns dropped or doubled, odd white space (ASCII and Unicode), CR and CR LF line ends."""
    src = read_golden("c1.gfa.gz").splitlines()
    alt = [l for l in src if l.startswith("S") and "." in l.split("\t")[1]]
    other = []
    for l in src:
        if l.startswith("S") and "." not in l.split("\t")[1]:
            c = l.split("\t")
            other.append("\t".join(c[:2] + [c[2][:24]]))
        elif l.startswith(("L", "#")):
            other.append(l[:200])
        elif l.startswith("P"):
            other.append(l[:120])
    rng = random.Random(seed)
    alphabet = "\t\t .:S-+ACGTacgt\r\x0b\x1c\xa0\u2003\u3000\x85é"
    out = []
    while len(out) < n:
        lines = [rng.choice(alt) for _ in range(rng.choice((1, 2, 4)))] + [rng.choice(other) for _ in range(rng.choice((0, 2, 5)))]
        rng.shuffle(lines)
        for _ in range(rng.choice((0, 1, 1, 2, 3))):
            j = rng.randrange(len(lines))
            line = lines[j]
            k = rng.randrange(8)
            pos = rng.randrange(len(line) + 1)
            cols = line.split("\t")
            if k == 0 and pos < len(line):
                line = line[:pos] + rng.choice(alphabet) + line[pos + 1:]
            elif k == 1:
                line = line[:pos] + rng.choice(alphabet) + line[pos:]
            elif k == 2:
                line = "\t".join(cols[:rng.randrange(len(cols) + 1)])
            elif k == 3:
                line = line + rng.choice((" ", "\t", " \t ", "\xa0", "\u2003\t", "\tLN:i:5", "\x1c", "é"))
            elif k == 4 and len(cols) > 2:
                cols[2] = rng.choice(("", "*", " ", "ACGT ", "ÀÉ", cols[2] + "\t" + cols[2], "A" * 300))
                line = "\t".join(cols)
            elif k == 5 and len(cols) > 1:
                cols[1] = rng.choice(("", "x", "a.b", "chr1:5.1", "1:2:3.4", "chr1:5.1:7", ".", cols[1] + ".2", "é:1.1", "chr1:0123456789.1"))
                line = "\t".join(cols)
            elif k == 6:
                line = rng.choice(("S", "Sx", "", "S\t", "S\t\t", "SEQ\tchr1:9.9\tAC")) + (line[pos:] if rng.random() < 0.5 else "")
            else:
                lines.insert(j, line)                       # the same node twice: the last length wins
            lines[j] = line
        eol = rng.choice(("\n", "\n", "\n", "\r\n", "\r"))
        text = eol.join(lines) + rng.choice((eol, eol, ""))
        out.append(text)
    return out


def main_gfa():
    """alt_node_len as the unmodified reference builds it (filter-alignments.py:103-113): main() is run
    on an empty GAF and the local dictionary is read from its frame when it returns."""
    spec = importlib.util.spec_from_file_location("ref_filter", os.path.join(REF, "filter-alignments.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    seen = {}

    def prof(frame, event, arg):
        if event == "return" and frame.f_code.co_name == "main" and frame.f_code.co_filename.endswith("filter-alignments.py"):
            seen["alt"] = dict(frame.f_locals.get("alt_node_len") or {})

    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            open("p.gaf", "w").write("")
            open("p_svs_edges.json", "w").write("{}")
            for text in damaged_gfas(1200):
                with open("p.gfa", "w", newline="") as fh:
                    fh.write(text)
                seen.clear()
                try:
                    sys.argv = ["filter-alignments.py", "-a", "p.gaf", "-g", "p.gfa", "-p", "p"]
                    sys.setprofile(prof)
                    try:
                        with redirect_stdout(io.StringIO()), redirect_stderr(io.StringIO()):
                            ref.main(sys.argv[1:])
                    finally:
                        sys.setprofile(None)
                    cases.append({"gfa": text, "rc": 0, "alt": seen["alt"]})
                except BaseException:
                    cases.append({"gfa": text, "rc": 1})
        finally:
            os.chdir(cwd)
    with gzip.GzipFile(os.path.join(HERE, "fuzz_gfa.json.gz"), "wb", mtime=0) as fh:
        fh.write(json.dumps(cases, ensure_ascii=True).encode())
    print(len(cases), "gfa cases;", sum(c["rc"] for c in cases), "raise;", sum(len(c.get("alt", ())) for c in cases), "alt nodes")


def damaged_jsons(n, seed=23):
    """informative_aln.json files as a user (or another tool) could hand them to predict-genotype.py: the
    c1 dictionary with short stand-in strings for the GAF lines (only the list lengths matter, :219-226),
    then odd values, escapes, repeated keys and byte-level damage."""
    base = json.loads(read_golden("c1_informative_aln.json.gz"))
    base = {k: [[f"r{i}\tx\n" for i in range(len(v[0]))], [f"a{i}\n" for i in range(len(v[1]))]] for k, v in base.items()}
    keys = sorted(base)
    rng = random.Random(seed)
    odd_values = ['"xyz"', '"x"', '""', '5', 'null', 'true', '[]', '[[]]', '[[],[]]', '[[], [], 5]', '{"0": [], "1": []}', '[5, []]',
                  '["ab", "c"]', '["é€", ""]', '[{"x":1,"x":2,"y":3}, [1,2,3]]', '[[1e5, -0.5, true, null, {"k": [1]}], [[[]]]]',
                  '[[NaN, Infinity], [-Infinity]]', '[[01],[]]', '[[1.],[]]', '[[1.5e],[]]', '[[-0, 0.0e0, 1E+2],[]]', '[["\t"],[]]',
                  '[["\\t\\u00e9\\ud83d\\ude00"],["\\ud800"]]', '[["\\x"],[]]', '[[1,],[]]', '[[1 2],[]]', '[["a" "b"],[]]',
                  '[null, null]', '[[], null]', '[3.5, "abc"]', '{"a": 1}', '[[], {"p": 1, "q": 2}]', '[["x"],["y"]] ', '[["x"]\r\n,\t["y"]]']
    alphabet = '"\\,:[]{}019eE.+-tfn \t\n\x01\xe9'
    out = []
    while len(out) < n:
        d = dict(base)
        text_keys = {}
        for _ in range(rng.choice((0, 1, 2, 3))):
            k = rng.choice(keys)
            m = rng.randrange(5)
            if m == 0:
                text_keys[k] = rng.choice(odd_values)
            elif m == 1:                                   # the key spelled with escapes
                text_keys[k] = None
            elif m == 2:
                d.pop(k, None)
            elif m == 3:
                d[k] = [d[k][0][: rng.randrange(4)], d[k][1][: rng.randrange(4)]] if k in d else [[], []]
            else:
                d[k + rng.choice(("", " ", "x"))] = [[], []]
        parts = []
        for k in rng.sample(sorted(d), len(d)) if rng.random() < 0.3 else sorted(d):
            ks = json.dumps(k)
            if k in text_keys and text_keys[k] is None:
                j = rng.randrange(len(k))
                ks = '"' + k[:j] + "\\u%04x" % ord(k[j]) + k[j + 1:] + '"'
            v = text_keys[k] if text_keys.get(k) is not None else json.dumps(d[k])
            parts.append(ks + ": " + v)
            if rng.random() < 0.03:                        # the same key again: the last one wins
                parts.append(ks + ": " + rng.choice(('[[],[]]', '[["q"],["q","q"]]', '7')))
        text = "{" + rng.choice((", ", ",\n    ")).join(parts) + "}" + rng.choice(("", "\n", " "))
        if rng.random() < 0.35:
            for _ in range(rng.choice((1, 1, 2))):
                pos = rng.randrange(len(text) + 1)
                k = rng.randrange(3)
                if k == 0 and pos < len(text):
                    text = text[:pos] + rng.choice(alphabet) + text[pos + 1:]
                elif k == 1 and pos < len(text):
                    text = text[:pos] + text[pos + 1:]
                else:
                    text = text[:pos] + rng.choice(alphabet) + text[pos:]
        out.append(text)
    return out


def main_json():
    spec = importlib.util.spec_from_file_location("ref_pg", os.path.join(REF, "predict-genotype.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        cwd = os.getcwd()
        os.chdir(tmp)
        try:
            open("in.vcf", "w").write(read_golden("c1.vcf"))
            for text in damaged_jsons(600):
                with open("aln.json", "w", newline="") as fh:
                    fh.write(text)
                if os.path.exists("out.vcf"):
                    os.remove("out.vcf")
                buf = io.StringIO()
                try:
                    sys.argv = ["predict-genotype.py", "-d", "aln.json", "-v", "in.vcf", "-o", "out.vcf"]
                    with redirect_stdout(buf), redirect_stderr(io.StringIO()):
                        ref.main(sys.argv[1:])
                    cases.append({"in": hashlib.sha256(text.encode()).hexdigest()[:12], "rc": 0,
                                  "sha256": hashlib.sha256(open("out.vcf", "rb").read()).hexdigest(), "stdout": buf.getvalue()})
                except BaseException:
                    cases.append({"in": hashlib.sha256(text.encode()).hexdigest()[:12], "rc": 1})
        finally:
            os.chdir(cwd)
    with open(os.path.join(HERE, "fuzz_json.json"), "w") as fh:      # the inputs are regenerated from the seed (damaged_jsons)
        json.dump(cases, fh)
    print(len(cases), "json cases;", sum(c["rc"] for c in cases), "raise;", len({c.get("sha256") for c in cases}), "distinct outputs")


if __name__ == "__main__":
    {"lines": main, "vcf": main_vcf, "edges": main_edges, "gfa": main_gfa, "json": main_json}[sys.argv[1] if len(sys.argv) > 1 else "lines"]()
