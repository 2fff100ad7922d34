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
