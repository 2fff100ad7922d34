"""Drop-in construct-graph.py (svjg/graphgen.py::construct_main) against the
unmodified reference graph constructor: tests/golden/fuzz_graph.json.gz holds, for 1500 small genomes
x catalogues (sound and damaged), the exit status, stdout and the SHA-256 of each file the reference
left behind (tests/golden/make_graph_fuzz.py ran it; nothing here reads /root/reference)."""
import gzip
import hashlib
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict
from contextlib import redirect_stderr, redirect_stdout

import pytest

from conftest import GOLDEN, PKG, read_golden
from svjg import graphgen

OUT_FILES = ("x.gfa", "x_svs_edges.json", "x_ignored_svs.txt")


def _sha(path):
    if not os.path.exists(path):
        return None
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def _run(argv):
    out, err = io.StringIO(), io.StringIO()
    rc, msg = 0, ""
    try:
        with redirect_stdout(out), redirect_stderr(err):
            graphgen.construct_main(argv)
    except SystemExit as exc:                       # sys.exit("text") is exit status 1
        rc, msg = (1, exc.code) if isinstance(exc.code, str) else (exc.code or 0, "")
    return rc, out.getvalue(), msg


def test_damaged_catalogues_match_the_reference(tmp_path, monkeypatch):
    with gzip.open(os.path.join(GOLDEN, "fuzz_graph.json.gz")) as fh:
        fix = json.loads(fh.read())
    assert len(fix["cases"]) == 1500
    monkeypatch.chdir(tmp_path)
    n_stopped = 0
    for i, case in enumerate(fix["cases"]):
        for f in OUT_FILES:
            if os.path.exists(f):
                os.remove(f)
        with open("x.fa", "w", newline="") as fh:
            fh.write(fix["fastas"][case["fa"]])
        with open("x.vcf", "w", newline="") as fh:
            fh.write(case["vcf"])
        rc, out, msg = _run(["-v", "x.vcf", "-r", "x.fa", "-o", "x.gfa"])
        assert rc == case["rc"], (i, msg, case["stderr_last"])
        assert out == case["stdout"], i
        for f in OUT_FILES:                         # complete, cut short or absent: as the reference left it
            assert _sha(f) == case["files"][f], (i, f)
        if rc:
            n_stopped += 1
            want = case["stderr_last"].split(":")[0]
            if want == "Error":                     # the reference's own sys.exit text
                assert msg == case["stderr_last"], i
            else:                                   # a traceback there: same exception class here
                assert msg.split(":")[1].strip() == want, (i, msg, case["stderr_last"])
    assert n_stopped > 300


@pytest.mark.parametrize("tag", ["c1", "s2", "s3", "s4"])
def test_edges_writer_is_json_dumps(tag):
    """write_edges_json() must be byte-equal to json.dumps(sort_keys=True, indent=4) — on the golden
    tables (which the reference wrote) and on keys that need escaping."""
    name = f"{tag}_svs_edges.json" + ("" if tag == "c1" else ".gz")
    text = read_golden(name)
    g = graphgen.Graph(OrderedDict())
    g.link_sv = {k: [tuple(e) for e in v] for k, v in json.loads(text).items()}
    buf = io.StringIO()
    g.write_edges_json(buf)
    assert buf.getvalue() == text
    g.link_sv = {'a"b\\c:1-2@+@é:3-4@+': [("é:DEL-2-3", 0)], "z": [], "\x01": [("x", 1), ("y", 0)]}
    buf = io.StringIO()
    g.write_edges_json(buf)
    assert buf.getvalue() == g.edges_json()
    g.link_sv = {}
    buf = io.StringIO()
    g.write_edges_json(buf)
    assert buf.getvalue() == "{}"


def test_fasta_loader(tmp_path):
    p = tmp_path / "a.fa"
    p.write_bytes(b"\n>one desc\r\nacgt\r\nNn\r\n>empty\n>two\tx\nAC\rGT\n\n>last")
    assert graphgen.load_fasta(str(p)) == OrderedDict([("one", "ACGTNN"), ("two", "ACGT"), ("last", "")])
    p.write_bytes(b"ACGT\n>a\nAC\n")
    with pytest.raises(UnboundLocalError):
        graphgen.load_fasta(str(p))
    p.write_bytes(b">\nAC\n")
    with pytest.raises(IndexError):
        graphgen.load_fasta(str(p))


def test_script_default_output_names(tmp_path):
    """No -o: variation_graph.gfa, svs_edges.json and ignored_svs.txt in the working directory
    (construct-graph.py:56-63); run as the script a user (or svjedi-graph.py:93) starts."""
    (tmp_path / "r.fa").write_text(">chr1\n" + "ACGT" * 100 + "\n")
    (tmp_path / "v.vcf").write_text("#h\nchr1\t100\ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL;END=150\n"
                                    "chr1\t200\tb\tNN\tACGT\t.\t.\tSVTYPE=INS\n")
    p = subprocess.run([sys.executable, os.path.join(PKG, "construct-graph.py"), "-v", "v.vcf", "-r", "r.fa"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout == "", p.stderr
    gfa = (tmp_path / "variation_graph.gfa").read_text()
    assert gfa.startswith("#chr1\tDEL-100-150\nS\tchr1:1-100\t")
    assert gfa.endswith("L\tchr1:1-100\t+\tchr1:151-400\t+\t0M\n")
    edges = json.loads((tmp_path / "svs_edges.json").read_text())
    assert edges["chr1:1-100@+@chr1:151-400@+"] == [["chr1:DEL-100-150", 1]]
    assert (tmp_path / "ignored_svs.txt").read_text().endswith("wrong format\nchr1\t200\tb\tNN\tACGT\t.\t.\tSVTYPE=INS")
    p = subprocess.run([sys.executable, os.path.join(PKG, "construct-graph.py"), "-v", "v.vcf", "-r", "missing.fa"],
                       cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 1


def test_compressed_inputs_give_the_same_files(tmp_path, monkeypatch):
    """Extension (row N4): a gzip VCF / FASTA gives the bytes the plain files give."""
    import gzip as gz
    monkeypatch.chdir(tmp_path)
    with gz.open(os.path.join(GOLDEN, "fuzz_graph.json.gz")) as fh:
        fix = json.loads(fh.read())
    case = next(c for c in fix["cases"] if c["rc"] == 0 and c["vcf"].count("\n") > 20)
    fa, vcf = fix["fastas"][case["fa"]].encode(), case["vcf"].encode()
    (tmp_path / "x.fa.gz").write_bytes(gz.compress(fa))
    (tmp_path / "x.vcf.gz").write_bytes(gz.compress(vcf))
    rc, out, msg = _run(["-v", "x.vcf.gz", "-r", "x.fa.gz", "-o", "x.gfa"])
    assert rc == 0 and out == case["stdout"], msg
    for f in OUT_FILES:
        assert _sha(f) == case["files"][f], f
