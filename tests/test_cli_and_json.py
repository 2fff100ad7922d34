"""The drop-in command lines and the informative_aln.json reader.
CPU part: the JSON reader against json.load on the reference's golden output.
GPU part (-m gpu): the three front-ends run as subprocesses, byte-for-byte against the
outputs of the unmodified reference scripts (tests/golden, made by make_golden.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG, read_golden
from svjg import genotype


def test_aln_counts_reader_matches_json_load():
    text = read_golden("c1_informative_aln.json.gz")
    want = {k: [len(v[0]), len(v[1])] for k, v in json.loads(text).items()}
    c = genotype.AlnCounts.from_memory(text)
    assert c.num == len(want)
    got = {c.key(i): c.counts[i].tolist() for i in range(c.num)}
    assert got == want
    assert [c.key(i) for i in range(c.num)] == sorted(want)
    some = sorted(want)[3]
    assert c.find(some) == 3 and c.find("nope") is None


def test_aln_counts_reader_odd_values():
    c = genotype.AlnCounts.from_memory(
        '{"k1": [[], []], "k2": [["a\\"]", "b"], ["c"], "ignored"], "k3": 7, "k4": [[1]], "k5": "xy",'
        ' "k6": [{"a": 1, "b": 2}, "héé"], "k2b": [[], [[1, 2], {"x": []}]], "k1": [["last", "wins"], []]}')
    got = {c.key(i): c.counts[i].tolist() for i in range(c.num)}
    bad = [0xFFFFFFFF, 0xFFFFFFFF]
    assert got == {"k1": [2, 0], "k2": [2, 1], "k3": bad, "k4": bad, "k5": [1, 1], "k6": [2, 3], "k2b": [0, 2]}
    with pytest.raises(Exception):
        genotype.AlnCounts.from_memory('{"k": [[], []]')
    assert genotype.AlnCounts.from_memory("{}").num == 0


def _write_inputs(tmp_path, tag="c1"):
    p = str(tmp_path / "t")
    open(p + ".gaf", "w").write(read_golden(f"{tag}.gaf.gz"))
    open(p + ".gfa", "w").write(read_golden(f"{tag}.gfa.gz"))
    open(p + "_svs_edges.json", "w").write(read_golden("c1_svs_edges.json" if tag == "c1" else f"{tag}_svs_edges.json.gz"))
    open(p + ".vcf", "w").write(read_golden("c1.vcf" if tag == "c1" else f"{tag}.vcf.gz"))
    return p


def _run(script, *args, env=None, cwd=None):
    return subprocess.run([sys.executable, os.path.join(PKG, script), *args], capture_output=True, text=True, env=env, cwd=cwd)


@pytest.mark.gpu
def test_filter_and_genotype_front_ends_match_reference(tmp_path):
    p = _write_inputs(tmp_path)
    r = _run("filter-alignments.py", "-a", p + ".gaf", "-g", p + ".gfa", "-p", p)
    assert r.returncode == 0, r.stderr
    assert open(p + "_informative_aln.json").read() == read_golden("c1_informative_aln.json.gz")
    r = _run("predict-genotype.py", "-d", p + "_informative_aln.json", "-v", p + ".vcf", "-o", p + "_genotype.vcf")
    assert r.returncode == 0, r.stderr
    assert r.stdout == read_golden("c1_stdout.txt")
    assert open(p + "_genotype.vcf").read() == read_golden("c1_genotype.vcf")
    r = _run("predict-genotype.py", "-d", p + "_informative_aln.json", "-v", p + ".vcf", "-o", p + "_g2.vcf",
             "-ms", "40", "-e", "0.001")
    assert r.returncode == 0, r.stderr
    assert open(p + "_g2.vcf").read() == read_golden("c1_genotype_ms40_e1e-3.vcf")
    # -o <dir> prefixes the output path as the reference does (:83-84)
    os.makedirs(tmp_path / "out" / str(tmp_path).lstrip("/"), exist_ok=True)
    r = _run("filter-alignments.py", "-a", p + ".gaf", "-g", p + ".gfa", "-p", p, "-o", str(tmp_path / "out"))
    assert r.returncode == 0, r.stderr
    assert os.path.exists(str(tmp_path / "out") + "/" + p + "_informative_aln.json")


@pytest.mark.gpu
def test_front_end_failures_exit_1(tmp_path):
    p = _write_inputs(tmp_path)
    # no -p: the reference dies with UnboundLocalError (exit status 1)
    assert _run("filter-alignments.py", "-a", p + ".gaf", "-g", p + ".gfa").returncode == 1
    # -O: TypeError at the first overlap test
    assert _run("filter-alignments.py", "-a", p + ".gaf", "-g", p + ".gfa", "-p", p, "-O", "50").returncode == 1
    # malformed GAF line
    open(p + "_bad.gaf", "w").write(read_golden("c1.gaf.gz") + "only\tthree\tcolumns\n")
    r = _run("filter-alignments.py", "-a", p + "_bad.gaf", "-g", p + ".gfa", "-p", p)
    assert r.returncode == 1 and "reference raises" in r.stderr
    # VCF line with fewer than 8 columns
    open(p + "_aln.json", "w").write("{}")
    open(p + "_bad.vcf", "w").write("1\t100\tid\tN\t<DEL>\n")
    assert _run("predict-genotype.py", "-d", p + "_aln.json", "-v", p + "_bad.vcf", "-o", p + "_x.vcf").returncode == 1


@pytest.mark.gpu
def test_pipeline_front_end_with_stub_tools(tmp_path):
    """svjedi-graph.py end to end; construct-graph and minigraph (absent offline) are stubs that
    deliver the golden graph files and the golden GAF."""
    src = _write_inputs(tmp_path, "c1")
    bindir = tmp_path / "bin"
    bindir.mkdir()
    stub = bindir / "construct_stub.py"
    stub.write_text(
        "import shutil, sys\n"
        "out = sys.argv[sys.argv.index('-o') + 1]\n"
        f"shutil.copy({src + '.gfa'!r}, out)\n"
        f"shutil.copy({src + '_svs_edges.json'!r}, out[:-4] + '_svs_edges.json')\n")
    mg = bindir / "minigraph"
    mg.write_text(f"#!/bin/sh\ncat {src}.gaf\n")
    mg.chmod(0o755)
    env = dict(os.environ, PATH=f"{bindir}:{os.environ['PATH']}", SVJG_CONSTRUCT_GRAPH=str(stub))
    prefix = str(tmp_path / "run")
    open(tmp_path / "reads.fq", "w").write("")
    r = _run("svjedi-graph.py", "-v", src + ".vcf", "-r", "ref.fa", "-q", str(tmp_path / "reads.fq"), "-p", prefix, env=env)
    assert r.returncode == 0, r.stderr
    assert open(prefix + "_informative_aln.json").read() == read_golden("c1_informative_aln.json.gz")
    assert open(prefix + "_genotype.vcf").read() == read_golden("c1_genotype.vcf")
    assert r.stdout.endswith(read_golden("c1_stdout.txt"))
    assert "Constructing variation graph...\nMapping reads on graph...\nFiltering alignment file...\nGenotyping SVs...\n" in r.stdout


@pytest.mark.gpu
def test_error_rate_outside_0_1_fails_only_with_a_genotyped_record(tmp_path):
    """predict-genotype.py -e 2: math.log10 raises inside likelihood() (:295-299), so the reference stops only when
    some record passes the gate at :216; otherwise it writes "./." everywhere and exits 0 (checked on the
    unmodified script when this test was written)."""
    vcf = ("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n"
           "chr1\t1000\ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL;END=1200\n")
    (tmp_path / "in.vcf").write_text(vcf)
    (tmp_path / "empty.json").write_text("{}")
    (tmp_path / "one.json").write_text('{"chr1:DEL-1000-1200": [["x\\n", "y\\n", "z\\n"], []]}')
    r = _run("predict-genotype.py", "-d", str(tmp_path / "empty.json"), "-v", str(tmp_path / "in.vcf"), "-o", str(tmp_path / "o1.vcf"), "-e", "2")
    assert r.returncode == 0 and r.stdout == "Genotyped svs: 0\n", r.stderr
    assert open(tmp_path / "o1.vcf").read().endswith("SVTYPE=DEL;END=1200\tGT:DP:AD:PL\t./.:0:0,0:.,.,.\n")
    r = _run("predict-genotype.py", "-d", str(tmp_path / "one.json"), "-v", str(tmp_path / "in.vcf"), "-o", str(tmp_path / "o2.vcf"), "-e", "2")
    assert r.returncode == 1


@pytest.mark.gpu
def test_new_switches_min_overlap_and_min_identity(tmp_path):
    """The two switches the reference does not have, each against its oracle-side statement
    (oracle.filter_alignments d_over / min_identity): --min-overlap N is the threshold -O was meant to set
    (filter-alignments.py:56, :269; -O itself still stops the run), --min-identity X drops alignments by the
    identity :193-196 parses (id:f: tag, else Am / Alen).  Both leave the default run untouched."""
    from conftest import alt_len_from_gfa_text
    from oracle import svjg_oracle as O
    tag = "s3"
    p = _write_inputs(tmp_path, tag)
    edges = json.loads(read_golden(f"{tag}_svs_edges.json.gz"))
    alt = alt_len_from_gfa_text(read_golden(f"{tag}.gfa.gz"))
    lines = read_golden(f"{tag}.gaf.gz").splitlines(True)
    # identities: every third line gets an id:f: tag (some below, some above, odd spellings float() accepts)
    vals = ["0.5", "0.91", "0.899999", " 0.95 ", "9e-1", "1", ".97", "0.90000000000000002"]
    lines = [l[:-1] + f"\tid:f:{vals[i % len(vals)]}\n" if i % 3 == 0 else l for i, l in enumerate(lines)]
    with open(p + ".gaf", "w") as fh:
        fh.writelines(lines)
    for args, kw in ((["--min-overlap", "250"], {"d_over": 250}), (["--min-overlap", "0"], {"d_over": 0}),
                     (["--min-identity", "0.9"], {"min_identity": 0.9}), (["--min-identity", "0.9", "--min-overlap", "150"], {"min_identity": 0.9, "d_over": 150}),
                     ([], {})):
        r = _run("filter-alignments.py", "-a", p + ".gaf", "-g", p + ".gfa", "-p", p, *args)
        assert r.returncode == 0, r.stderr
        want = O.dumps_informative(O.filter_alignments(lines, edges, alt, **kw))
        assert open(p + "_informative_aln.json").read() == want, args
    base = O.filter_alignments(lines, edges, alt)
    assert O.filter_alignments(lines, edges, alt, min_identity=0.9) != base != O.filter_alignments(lines, edges, alt, d_over=250)
    # a value float() rejects on a line that is stored: exit status 1 (the reference raises on such a line at :194)
    bad = list(lines)
    k = next(i for i, l in enumerate(bad) if i % 3 == 0 and O.record_hits(l, edges, alt))
    bad[k] = bad[k].replace("id:f:", "id:f:x")
    with open(p + ".gaf", "w") as fh:
        fh.writelines(bad)
    r = _run("filter-alignments.py", "-a", p + ".gaf", "-g", p + ".gfa", "-p", p, "--min-identity", "0.9")
    assert r.returncode == 1 and "id:f:" in r.stderr
    # -O keeps the reference's behaviour
    r = _run("filter-alignments.py", "-a", p + ".gaf", "-g", p + ".gfa", "-p", p, "-O", "50")
    assert r.returncode == 1


@pytest.mark.gpu
def test_streamed_stages_on_the_device(tmp_path):
    """Row N3 on a GPU: a GAF piped into `filter-alignments.py -a /dev/stdin` (segments of whole lines read into
    page-locked buffers and filtered while the writer is still writing), and `SVJG_STREAM=1 svjedi-graph.py` with
    stub tools (mapping and filtering overlapped): the reference's files, byte for byte."""
    import hashlib
    import subprocess
    tag = "s3"
    p = _write_inputs(tmp_path, tag)
    with open(p + ".gaf", "rb") as fh:
        raw = fh.read()
    # the writer delivers the file in odd pieces with pauses, through a real pipe
    writer = tmp_path / "writer.py"
    writer.write_text("import sys, time\nraw = open(sys.argv[1], 'rb').read()\nstep = 1 + len(raw) // 37\n"
                      "for i in range(0, len(raw), step):\n    sys.stdout.buffer.write(raw[i:i + step]); sys.stdout.buffer.flush(); time.sleep(0.01)\n")
    cmd = f"{sys.executable} {writer} {p}.gaf | {sys.executable} {os.path.join(PKG, 'filter-alignments.py')} -a /dev/stdin -g {p}.gfa -p {p}"
    r = subprocess.run(cmd, shell=True, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    with open(p + "_informative_aln.json", "rb") as fh:
        assert hashlib.sha256(fh.read()).hexdigest() == read_golden(f"{tag}_informative_aln.sha256").strip()
    os.remove(p + "_informative_aln.json")
    # the orchestrator with mapping and filtering overlapped
    bindir = tmp_path / "bin"
    bindir.mkdir()
    stub = bindir / "construct_stub.py"
    stub.write_text(
        "import shutil, sys\n"
        "out = sys.argv[sys.argv.index('-o') + 1]\n"
        f"shutil.copy({p + '.gfa'!r}, out)\n"
        f"shutil.copy({p + '_svs_edges.json'!r}, out[:-4] + '_svs_edges.json')\n")
    mg = bindir / "minigraph"
    mg.write_text(f"#!/bin/sh\n{sys.executable} {writer} {p}.gaf\n")
    mg.chmod(0o755)
    env = dict(os.environ, PATH=f"{bindir}:{os.environ['PATH']}", SVJG_CONSTRUCT_GRAPH=str(stub), SVJG_STREAM="1")
    prefix = str(tmp_path / "run")
    open(tmp_path / "reads.fq", "w").write("")
    r = _run("svjedi-graph.py", "-v", p + ".vcf", "-r", "ref.fa", "-q", str(tmp_path / "reads.fq"), "-p", prefix, env=env)
    assert r.returncode == 0, r.stderr
    with open(prefix + "_informative_aln.json", "rb") as fh:
        assert hashlib.sha256(fh.read()).hexdigest() == read_golden(f"{tag}_informative_aln.sha256").strip()
    assert open(prefix + "_genotype.vcf").read() == read_golden(f"{tag}_genotype.vcf.gz")
    assert open(prefix + ".gaf", "rb").read() == raw
    assert r.stdout.endswith(read_golden(f"{tag}_stdout.txt"))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["s2", "s3", "s4"])
def test_front_ends_on_scaled_configs(tmp_path, tag):
    """The three scaled-down BASELINE.json shapes through the command lines (no tensor library in these
    processes): informative_aln.json by its SHA-256, genotype VCF and stdout byte for byte; the GAF and
    the VCF are handed over gzip-compressed for s3 (extension, svjg/gzio.py)."""
    import gzip
    import hashlib
    p = _write_inputs(tmp_path, tag)
    gaf, vcf = p + ".gaf", p + ".vcf"
    if tag == "s3":
        for f in (gaf, vcf):
            with open(f, "rb") as src, open(f + ".gz", "wb") as dst:
                dst.write(gzip.compress(src.read(), 1))
        gaf, vcf = gaf + ".gz", vcf + ".gz"
    r = _run("filter-alignments.py", "-a", gaf, "-g", p + ".gfa", "-p", p)
    assert r.returncode == 0, r.stderr
    with open(p + "_informative_aln.json", "rb") as fh:
        assert hashlib.sha256(fh.read()).hexdigest() == read_golden(f"{tag}_informative_aln.sha256").strip()
    r = _run("predict-genotype.py", "-d", p + "_informative_aln.json", "-v", vcf, "-o", p + "_genotype.vcf")
    assert r.returncode == 0, r.stderr
    assert r.stdout == read_golden(f"{tag}_stdout.txt")
    assert open(p + "_genotype.vcf").read() == read_golden(f"{tag}_genotype.vcf.gz")
