"""oracle/svjg_oracle.c (the C restatement used as the checker at full size) pinned on the outputs of the
unmodified reference — the same fixtures that pin the Python oracle — and held against the Python
oracle on generated workloads.  CPU only."""
import hashlib
import importlib.util
import io
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, alt_len_from_gfa_text, read_golden, revisit_walks
from oracle import svjg_oracle as O


@pytest.fixture(scope="module")
def CO():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    from oracle import c_oracle
    return c_oracle


def _informative(CO, tables, gaf_bytes, **kw):
    """The reference's dictionary rebuilt from the C oracle's hit tuples (file order)."""
    counts, stats, sv2, off, ln = CO.filter_counts(tables, gaf_bytes, want_hits=True, **kw)
    d = {}
    for s, o, n in zip(sv2.tolist(), off.tolist(), ln.tolist()):
        slot = d.setdefault(tables.sv_ids[s >> 1], [[], []])
        slot[s & 1].append(O.kept_text(gaf_bytes[o:o + n].decode()))
    return d, counts, stats


@pytest.mark.parametrize("tag", ["c1", "s2", "s3", "s4"])
def test_golden_outputs_of_the_reference(CO, tag):
    edges = json.loads(read_golden("c1_svs_edges.json" if tag == "c1" else f"{tag}_svs_edges.json.gz"))
    alt = alt_len_from_gfa_text(read_golden(f"{tag}.gfa.gz"))
    gaf = read_golden(f"{tag}.gaf.gz").encode()
    t = CO.Tables(edges, alt)
    for threads in (1, 5):
        d, counts, stats = _informative(CO, t, gaf, threads=threads)
        js = O.dumps_informative(d)
        if tag == "c1":
            assert js == read_golden("c1_informative_aln.json.gz")
        else:
            assert hashlib.sha256(js.encode()).hexdigest() == read_golden(f"{tag}_informative_aln.sha256").strip()
            assert {k: list(v) for k, v in CO.counts_dict(t, counts).items()} == json.loads(read_golden(f"{tag}_counts.json.gz"))
        assert stats["n_hits"] == sum(len(v[0]) + len(v[1]) for v in d.values())
    assert CO.filter_counts(t, b"")[1]["n_hits"] == 0


def test_quirk_cases(CO, quirks):
    edges = json.loads(quirks["edges"])
    t = CO.Tables(edges, alt_len_from_gfa_text(quirks["gfa"]))
    n = 0
    for case in quirks["cases"]:
        gaf = "".join(O.text_mode_lines(case["gaf"])).encode()
        try:
            d, _c, _s = _informative(CO, t, gaf)
        except CO.OracleRaises:
            assert case["rc"] != 0, case["name"]
        except CO.OracleUnsupported:
            continue
        else:
            assert case["rc"] == 0 and O.dumps_informative(d) == case["json"], case["name"]
        n += 1
    assert n >= 20


def test_damaged_lines_match_the_reference(CO):
    edges = json.loads(read_golden("c1_svs_edges.json"))
    t = CO.Tables(edges, alt_len_from_gfa_text(read_golden("c1.gfa.gz")))
    fx = json.loads(read_golden("fuzz_lines.json.gz"))
    n_unsupported = 0
    for c in fx["cases"]:
        gaf = "".join(O.text_mode_lines(c["line"])).encode()
        try:
            counts, _s = CO.filter_counts(t, gaf)
        except CO.OracleRaises:
            assert c["rc"] == 1, c["line"]
            continue
        except CO.OracleUnsupported:
            n_unsupported += 1
            continue
        assert c["rc"] == 0, c["line"]
        assert {k: list(v) for k, v in CO.counts_dict(t, counts).items()} == c["counts"], c["line"]
    assert n_unsupported <= 25
    # the CR LF file
    gaf = "".join(O.text_mode_lines(fx["crlf"]["gaf"])).encode()
    d, _c, _s = _informative(CO, t, gaf)
    assert O.dumps_informative(d) == fx["crlf"]["json"]


def test_damaged_link_tables_match_the_reference(CO):
    spec = importlib.util.spec_from_file_location("make_fuzz", os.path.join(os.path.dirname(__file__), "golden", "make_fuzz.py"))
    mf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mf)
    alt = alt_len_from_gfa_text(read_golden("c1.gfa.gz"))
    gaf = "".join(mf.edges_gaf_lines()).encode()
    want = json.loads(read_golden("fuzz_edges.json"))
    n_raise = 0
    for text, c in zip(mf.damaged_edges(len(want)), want):
        t = CO.Tables(json.loads(text), alt)
        try:
            d, _c, _s = _informative(CO, t, gaf)
        except CO.OracleRaises:
            assert c["rc"] == 1
            n_raise += 1
            continue
        assert c["rc"] == 0
        assert hashlib.sha256(O.dumps_informative(d).encode()).hexdigest() == c["sha256"]
    assert n_raise > 20


@pytest.mark.parametrize("name,scale", [("C2", 0.02), ("C3", 0.004), ("C4", 0.02), ("C5", 0.0002)])
def test_generated_workloads_match_the_python_oracle(CO, name, scale):
    """The BASELINE.json shapes at the sizes of tests/test_gpu_parity.py::test_named_workloads_scaled_match_oracle
    (10 % of the records carry a cg:Z: tail): counters and informative_aln.json against the Python oracle."""
    from svjg import synth
    g, _vcf, gaf_text = synth.make_workload(name, scale=scale, cg_frac=0.1)
    buf = io.StringIO()
    g.write_gfa(buf)
    edges, alt = json.loads(g.edges_json()), alt_len_from_gfa_text(buf.getvalue())
    gaf = gaf_text.encode()
    t = CO.Tables(edges, alt)
    d, counts, stats = _informative(CO, t, gaf, threads=3)
    want = O.filter_alignments(gaf_text.splitlines(True), edges, alt)
    assert CO.counts_dict(t, counts) == O.hit_counts(want)
    assert O.dumps_informative(d) == O.dumps_informative(want)
    assert stats["n_hits"] > 100 and stats["n_multi"] > 0


def test_full_c2_batch_has_the_counts_the_gpu_reported(CO):
    """The whole BASELINE.json C2 batch (3 M records, 0.5 GB) through the C oracle on the CPU: the hit and
    multi-node record counts are the ones the CUDA path printed in every bench.py run of round 1
    (profiles/r1/bench_final_c2.json: hits_per_gpu, multi_node_records) — and stay so if the generator or
    either side changes."""
    from svjg import synth
    g, _vcf, gaf_text = synth.make_workload("C2", scale=1.0)
    buf = io.StringIO()
    g.write_gfa(buf)
    t = CO.Tables(json.loads(g.edges_json()), alt_len_from_gfa_text(buf.getvalue()))
    gaf = gaf_text.encode()
    del gaf_text
    counts, stats = CO.filter_counts(t, gaf)
    assert len(gaf) == 512_635_132                              # gaf_bytes_per_gpu of that run
    assert stats == {"n_hits": 1_108_909, "n_records": 3_000_000, "n_multi": 789_043}
    assert int(counts.sum()) == 1_108_909


def test_paths_that_revisit_nodes_equal_the_python_oracle(CO):
    """conftest.revisit_walks: the C restatement (the checker of the full-size GPU tests, where such paths are
    resolved in the fast route) against the line-by-line Python restatement."""
    edges, lines = revisit_walks()
    want = O.filter_alignments(lines, edges, {})
    assert sum(len(v[0]) + len(v[1]) for v in want.values()) > 500
    assert sum(len(set(re.findall(r"chr7:\d+-\d+", l.split("\t")[5]))) < l.split("\t")[5].count("chr7") for l in lines) > 300
    t = CO.Tables(edges, {})
    d, counts, stats = _informative(CO, t, "".join(lines).encode())
    assert d == want
