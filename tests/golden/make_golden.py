#!/usr/bin/env python3
"""Regenerates tests/golden/* by running the UNMODIFIED reference scripts.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

What it writes (all small, committed):
  c1.*          test-dir VCF -> real construct-graph.py -> GFA + svs_edges;
                synthetic GAF on that graph -> real filter-alignments.py ->
                informative_aln.json -> real predict-genotype.py -> genotype.vcf
  s2.* s3.* s4.* scaled-down C2/C3/C4 shapes: svjg.graphgen tables (asserted
                byte-equal to the real construct-graph.py on a random FASTA),
                synthetic GAF -> real filter / genotype outputs
  quirks.json   hand-made edge cases, each with the reference's exit code and output
  kat40.tsv     the 40 known-answer rows of test-dir/expected_genotype.vcf
  lik_random.tsv.gz   random (svtype, n0, n1, ms, e) -> reference likelihood() outputs

Nothing in the test-suite reads /root/reference; it only reads these files.
"""
import gzip
import hashlib
import importlib.util
import io
import json
import os
import subprocess
import sys
import tempfile
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "svjedi-graph_b200"))
from svjg import graphgen, synth  # noqa: E402


def run_ref(script, args, cwd):
    p = subprocess.run([sys.executable, os.path.join(REF, script)] + args, cwd=cwd,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    return p.returncode, p.stdout, p.stderr


def wgz(name, text):
    with gzip.GzipFile(os.path.join(HERE, name), "wb", mtime=0) as fh:
        fh.write(text.encode())


def wtxt(name, text):
    with open(os.path.join(HERE, name), "w") as fh:
        fh.write(text)


def ref_filter_and_genotype(tmp, prefix, vcf_text, gfa_text, edges_text, gaf_text, extra_gt_args=()):
    open(os.path.join(tmp, prefix + ".gfa"), "w").write(gfa_text)
    open(os.path.join(tmp, prefix + "_svs_edges.json"), "w").write(edges_text)
    open(os.path.join(tmp, prefix + ".gaf"), "w").write(gaf_text)
    open(os.path.join(tmp, prefix + ".vcf"), "w").write(vcf_text)
    rc, _, err = run_ref("filter-alignments.py", ["-a", prefix + ".gaf", "-g", prefix + ".gfa", "-p", prefix], tmp)
    assert rc == 0, err
    js = open(os.path.join(tmp, prefix + "_informative_aln.json")).read()
    rc, out, err = run_ref("predict-genotype.py", ["-d", prefix + "_informative_aln.json", "-v", prefix + ".vcf",
                                                   "-o", prefix + "_genotype.vcf", *extra_gt_args], tmp)
    assert rc == 0, err
    return js, open(os.path.join(tmp, prefix + "_genotype.vcf")).read(), out


def make_c1(tmp):
    vcf = open(os.path.join(REF, "test-dir/test.vcf")).read()
    rc, _, err = run_ref("construct-graph.py", ["-v", os.path.join(REF, "test-dir/test.vcf"), "-r",
                                                os.path.join(REF, "test-dir/reference_genome.fasta"), "-o", "c1.gfa"], tmp)
    assert rc == 0, err
    gfa = open(os.path.join(tmp, "c1.gfa")).read()
    edges = open(os.path.join(tmp, "c1_svs_edges.json")).read()
    # our builder must agree with the real one
    seqs, name = OrderedDict(), None
    for line in open(os.path.join(REF, "test-dir/reference_genome.fasta")):
        if line.startswith(">"):
            name = line[1:].split()[0]
            seqs[name] = []
        else:
            seqs[name].append(line.strip().upper())
    seqs = OrderedDict((k, "".join(v)) for k, v in seqs.items())
    g = graphgen.build_graph(OrderedDict((k, len(v)) for k, v in seqs.items()),
                             [l for l in vcf.splitlines() if not l.startswith("#")], seqs)
    buf = io.StringIO()
    g.write_gfa(buf)
    assert buf.getvalue() == gfa and g.edges_json() == edges
    gaf = synth.simulate_gaf(g, 3000, seed=1001, mean_len=8200, cg_frac=0.1)
    js, gt, out = ref_filter_and_genotype(tmp, "c1", vcf, gfa, edges, gaf)
    _, gt_ms1, _ = ref_filter_and_genotype(tmp, "c1", vcf, gfa, edges, gaf, ("-ms", "40", "-e", "0.001"))
    wtxt("c1.vcf", vcf)
    wgz("c1.gfa.gz", gfa)
    wtxt("c1_svs_edges.json", edges)
    wgz("c1.gaf.gz", gaf)
    wgz("c1_informative_aln.json.gz", js)
    wtxt("c1_genotype.vcf", gt)
    wtxt("c1_genotype_ms40_e1e-3.vcf", gt_ms1)
    wtxt("c1_stdout.txt", out)


def make_scaled(tmp, tag, name, scale, n_rec, seed, **gaf_kw):
    kind, n_sv, _, _, mean_len = synth.WORKLOADS[name]
    n_sv = int(n_sv * scale)
    chrom_len = synth.human_like_chroms(scale)
    rows = synth.catalogue(kind, n_sv, chrom_len, seed, ins_max=400, del_max=3000)
    vcf = synth.vcf_text(rows)
    g = graphgen.build_graph(chrom_len, rows)
    # real construct-graph on a random FASTA of the same lengths
    rng = np.random.Generator(np.random.PCG64(seed))
    seqs = OrderedDict()
    with open(os.path.join(tmp, tag + ".fa"), "w") as fh:
        for c, n in chrom_len.items():
            s = rng.choice(np.frombuffer(b"ACGT", dtype="S1"), size=n).tobytes().decode()
            seqs[c] = s
            fh.write(f">{c}\n{s}\n")
    open(os.path.join(tmp, tag + ".vcf"), "w").write(vcf)
    rc, _, err = run_ref("construct-graph.py", ["-v", tag + ".vcf", "-r", tag + ".fa", "-o", tag + "_ref.gfa"], tmp)
    assert rc == 0, err
    ref_edges = open(os.path.join(tmp, tag + "_ref_svs_edges.json")).read()
    assert ref_edges == g.edges_json(), tag
    g2 = graphgen.build_graph(chrom_len, rows, seqs)
    buf = io.StringIO()
    g2.write_gfa(buf)
    assert buf.getvalue() == open(os.path.join(tmp, tag + "_ref.gfa")).read(), tag
    assert g2.ignored_text() == open(os.path.join(tmp, tag + "_ref_ignored_svs.txt")).read(), tag
    buf = io.StringIO()
    g.write_gfa(buf)                      # '*' placeholders for reference nodes
    gfa = buf.getvalue()
    gaf = synth.simulate_gaf(g, n_rec, seed + 7, mean_len=mean_len * max(scale * 20, 0.1), **gaf_kw)
    js, gt, out = ref_filter_and_genotype(tmp, tag, vcf, gfa, g.edges_json(), gaf)
    wgz(tag + ".vcf.gz", vcf)
    wgz(tag + ".gfa.gz", gfa)
    wgz(tag + "_svs_edges.json.gz", g.edges_json())
    wgz(tag + ".gaf.gz", gaf)
    wtxt(tag + "_informative_aln.sha256", hashlib.sha256(js.encode()).hexdigest() + "\n")
    wgz(tag + "_counts.json.gz", json.dumps({k: [len(v[0]), len(v[1])] for k, v in json.loads(js).items()}, sort_keys=True))
    wgz(tag + "_genotype.vcf.gz", gt)
    wtxt(tag + "_stdout.txt", out)
    return len(json.loads(js))


# ---- hand-made edge cases -------------------------------------------------
def gaf_line(qid, path, tlen, ts, te, tags="tp:A:P\tcm:i:10", cols=None, nl="\n"):
    base = [qid, "5000", "0", "5000", "+", path, str(tlen), str(ts), str(te), "4500", "5000", "60"]
    if cols:
        for k, v in cols.items():
            base[k] = v
    return "\t".join(base) + ("\t" + tags if tags else "") + nl


def quirk_cases():
    E = {}   # shared edges
    E["11:1-5000@+@1:1-5000@-"] = [["11:DEL-5000-6000", 1]]
    E["11:1-5000@+@1:1-5000@+"] = [["11:DEL-5000-7000", 1]]
    E["A:1-1000@+@A:1001-2000@+"] = [["A:DEL-1000-1100", 0], ["A:INS-1000-1", 0]]
    E["A:1001-2000@-@A:1-1000@-"] = [["A:INV-900-2000", 1]]
    E["P:1-1000@+@P:1-1000@-"] = [["P:INV-1-1000", 1]]
    E["A:1-1000@+@A:1001.1@+"] = [["A:INS-1000-1", 1]]
    E["A:1001.1@+@A:1001-2000@+"] = [["A:INS-1000-1", 1]]
    E["A:1-1000@+@A:1-1000@+"] = [["A:DEL-rep", 1]]
    E["B:1-50@+@B:1-5000@+"] = [["B:DEL-pre", 1]]
    E["B:1-5000@+@B:1-50@+"] = [["B:DEL-pre2", 1]]
    E["C:1-600@-@C:601-1200@-"] = [["C:DEL-comma", 0]]
    E["C:601-1200@+@C:1-600@+"] = [["C:DEL-comma-rev", 1]]
    E["Z:1-1000@+@Z:1001-2000@+"] = []
    E["N:1-1000@+@N:1001-2000@+"] = [["nocolon", 1]]
    E["M:1-1000@+@M:1001.9@+"] = [["M:INS-1000-9", 1]]
    E["W:1-1000@+@W:abc@+"] = [["W:DEL-1", 1]]
    E["\u00e9:1-1000@+@\u00e9:1001-2000@+"] = [["\u00e9:DEL-1000-1100", 0]]
    gfa = "S\tA:1-1000\t*\nS\tA:1001.1\tACGTACGTAC" + "A" * 190 + "\nS\tA:1001.1\t" + "C" * 150 + "\nS\tQ:5.5\tAC\n"
    ok = {}
    ok["strand_first_occurrence"] = gaf_line("r1", ">11:1-5000>1:1-5000", 10000, 100, 9000)
    ok["repeated_node"] = gaf_line("r2", ">A:1-1000<A:1-1000", 2000, 100, 1500)
    ok["fwd_and_reverse_keys"] = gaf_line("r3", ">A:1-1000>A:1001-2000", 2000, 100, 1900) + \
        gaf_line("r4", "<A:1001-2000<A:1-1000", 2000, 100, 1900)
    ok["palindromic_link"] = gaf_line("r5", ">P:1-1000<P:1-1000", 2000, 10, 1990)
    ok["left_threshold"] = gaf_line("t900", ">A:1-1000>A:1001-2000", 2000, 900, 1900) + \
        gaf_line("t901", ">A:1-1000>A:1001-2000", 2000, 901, 1900)
    ok["right_threshold"] = gaf_line("e1099", ">A:1-1000>A:1001-2000", 2000, 100, 1099) + \
        gaf_line("e1098", ">A:1-1000>A:1001-2000", 2000, 100, 1098)
    ok["alt_node_len_last_wins"] = gaf_line("i1", ">A:1-1000>A:1001.1>A:1001-2000", 2150, 850, 1290) + \
        gaf_line("i2", ">A:1-1000>A:1001.1>A:1001-2000", 2150, 100, 1249) + \
        gaf_line("i3", ">A:1-1000>A:1001.1>A:1001-2000", 2150, 100, 1248)
    ok["cg_cut_and_no_final_newline"] = gaf_line("c1", ">A:1-1000>A:1001-2000", 2000, 100, 1900, tags="tp:A:P\tcg:Z:100M\tzz:i:1") + \
        gaf_line("cg:Z:name", ">A:1-1000>A:1001-2000", 2000, 100, 1900) + \
        gaf_line("c3", ">A:1-1000>A:1001-2000", 2000, 100, 1900, nl="")
    ok["prefix_token"] = gaf_line("p1", ">B:1-5000>B:1-50", 5050, 100, 5000) + gaf_line("p2", "<B:1-5000>B:1-50", 5050, 100, 5000) + \
        gaf_line("p3", ">B:1-50>B:1-5000", 5050, 0, 5000) + gaf_line("p4", ">B:1-50<B:1-5000", 5050, 0, 5000)
    ok["comma_path_leading_comma"] = gaf_line("g1", ",C:1-600+,C:601-1200+", 1200, 100, 1100)
    ok["single_node_forms"] = gaf_line("s1", "A:1-1000", 1000, 0, 900) + gaf_line("s2", ">A:1-1000", 1000, 0, 900) + \
        gaf_line("s3", "A:1-1000+", 1000, 0, 900)
    ok["empty_entry_list_bad_names_unused"] = gaf_line("z1", ">Z:1-1000>Z:1001-2000>Z:junk", 3000, 0, 2000)
    ok["non_ascii_and_escapes"] = gaf_line("r\u00e9ad\"\\x", ">\u00e9:1-1000>\u00e9:1001-2000", 2000, 100, 1900, tags="tp:A:P\tzz:Z:\u4e2d\U0001F600")
    ok["trailing_whitespace_columns"] = gaf_line("w1", ">A:1-1000>A:1001-2000", 2000, 100, 1900, tags="", nl="\t \n") + \
        gaf_line("w2", ">A:1-1000>A:1001-2000", 2000, 100, 1900, tags="tp:A:P \t", nl="\n")
    ok["signed_and_spaced_ints"] = gaf_line("n1", ">A:1-1000>A:1001-2000", "+2000", " 100", "1900 ") + \
        gaf_line("n2", ">A:1-1000>A:1001-2000", 2000, -5, 1900)
    ok["idf_tag"] = gaf_line("f1", ">A:1-1000>A:1001-2000", 2000, 100, 1900, tags="id:f:0.93\ttp:A:P", cols={10: "0"})
    ok["many_nodes"] = gaf_line("m1", "".join(f">K:{i * 100 + 1}-{i * 100 + 100}" for i in range(70)) + ">A:1-1000>A:1001-2000", 9000, 10, 8900)
    bad = {}
    bad["blank_line"] = gaf_line("b0", ">A:1-1000>A:1001-2000", 2000, 100, 1900) + "\n"
    bad["eleven_columns"] = "\t".join(["q", "1", "0", "1", "+", ">A:1-1000", "1000", "0", "1", "1", "1"]) + "\n"
    bad["twelfth_column_empty"] = "\t".join(["q", "1", "0", "1", "+", ">A:1-1000", "1000", "0", "1", "1", "1", ""]) + "\n"
    bad["non_integer_column"] = gaf_line("b1", "A:1-1000", 1000, 0, 900, cols={2: "x"})
    bad["float_integer_column"] = gaf_line("b1", "A:1-1000", 1000, 0, 900, cols={11: "60.0"})
    bad["alen_zero"] = gaf_line("b2", "A:1-1000", 1000, 0, 900, cols={10: "0"})
    bad["alt_node_missing"] = gaf_line("b3", ">M:1-1000>M:1001.9", 1200, 100, 1150)
    bad["bad_ref_node_name"] = gaf_line("b4", ">W:1-1000>W:abc", 2000, 100, 1900)
    bad["sv_id_without_colon"] = gaf_line("b5", ">N:1-1000>N:1001-2000", 2000, 100, 1900)
    bad["empty_path"] = gaf_line("b6", "", 2000, 100, 1900)
    bad["comma_path_two_nodes"] = gaf_line("b7", "C:1-600+,C:601-1200+", 1200, 100, 1100)
    return E, gfa, ok, bad


def make_quirks(tmp):
    E, gfa, ok, bad = quirk_cases()
    edges = json.dumps(E, sort_keys=True, indent=4)
    out = {"edges": edges, "gfa": gfa, "cases": []}
    for group, expect_ok in ((ok, True), (bad, False)):
        for name, gaf in group.items():
            for f, t in (("q.gfa", gfa), ("q_svs_edges.json", edges), ("q.gaf", gaf)):
                open(os.path.join(tmp, f), "w").write(t)
            js_path = os.path.join(tmp, "q_informative_aln.json")
            if os.path.exists(js_path):
                os.remove(js_path)
            rc, _, err = run_ref("filter-alignments.py", ["-a", "q.gaf", "-g", "q.gfa", "-p", "q"], tmp)
            assert (rc == 0) == expect_ok, (name, rc, err[-400:])
            out["cases"].append({"name": name, "gaf": gaf, "rc": rc,
                                 "json": open(js_path).read() if rc == 0 else None,
                                 "error": None if rc == 0 else err.strip().splitlines()[-1]})
    wtxt("quirks.json", json.dumps(out, indent=1, ensure_ascii=True))


def make_kat40():
    rows = []
    for line in open(os.path.join(REF, "test-dir/expected_genotype.vcf")):
        if line.startswith("#"):
            continue
        cols = line.rstrip("\n").split("\t")
        svtype = cols[7].split("SVTYPE=")[1].split(";")[0]
        sample = cols[-1]
        ad = sample.split(":")[2].split(",")
        c = [float(x) for x in ad]
        n0 = int(round(c[0] * 2)) if svtype == "DEL" else int(c[0])
        n1 = int(round(c[1] * 2)) if svtype == "INS" else int(c[1])
        rows.append(f"{cols[2]}\t{svtype}\t{n0}\t{n1}\t{sample}\n")
    assert len(rows) == 40
    wtxt("kat40.tsv", "".join(rows))


def make_lik_random():
    spec = importlib.util.spec_from_file_location("ref_pg", os.path.join(REF, "predict-genotype.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from decimal import getcontext
    getcontext().prec = 28
    rng = np.random.Generator(np.random.PCG64(4242))
    out = []
    types = ["DEL", "INS", "INV", "BND"]
    es = [0.00005, 0.00005, 0.00005, 0.001, 0.01, 1e-9, 0.3]
    n = 30000
    hi = np.where(rng.random(n) < 0.85, 80, np.where(rng.random(n) < 0.9, 2500, 60_000))
    a = (rng.random(n) * hi).astype(np.int64)
    b = (rng.random(n) * hi).astype(np.int64)
    zero = rng.random(n)
    a[zero < 0.08] = 0
    b[(zero > 0.08) & (zero < 0.16)] = 0
    for i in range(n):
        t = types[int(rng.integers(0, 4))]
        e = es[int(rng.integers(0, len(es)))]
        ms = int(rng.integers(0, 8))
        cnt = [int(a[i]), int(b[i])]
        geno, prob = mod.likelihood(cnt, t, ms, e)
        numbers = ",".join(str(y) for y in cnt)
        dp = str(round(sum(cnt), 3))
        out.append(f"{t}\t{int(a[i])}\t{int(b[i])}\t{ms}\t{e!r}\t{geno}\t{dp}\t{numbers}\t{','.join(prob)}\n")
    wgz("lik_random.tsv.gz", "".join(out))


def main():
    with tempfile.TemporaryDirectory() as tmp:
        make_kat40()
        make_lik_random()
        make_quirks(tmp)
        make_c1(tmp)
        n2 = make_scaled(tmp, "s2", "C2", 0.012, 6000, 1002)
        n3 = make_scaled(tmp, "s3", "C3", 0.004, 6000, 1003, idf=True)
        n4 = make_scaled(tmp, "s4", "C4", 0.015, 6000, 1004, cg_frac=0.2)
        print("sv keys with hits: s2", n2, "s3", n3, "s4", n4)


if __name__ == "__main__":
    main()
