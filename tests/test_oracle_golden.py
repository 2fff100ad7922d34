"""Pins oracle/svjg_oracle.py to outputs of the unmodified reference
(tests/golden/, produced by tests/golden/make_golden.py)."""
import hashlib
import json

import pytest

from conftest import alt_len_from_gfa_text, read_golden
from oracle import svjg_oracle as O


def test_kat40_expected_genotype_rows():
    rows = read_golden("kat40.tsv").splitlines()
    assert len(rows) == 40
    for row in rows:
        _vid, svtype, n0, n1, sample = row.split("\t")
        gt, dp, ad, pl = O.genotype_counts(int(n0), int(n1), svtype)
        assert f"{gt}:{dp}:{ad}:{','.join(pl)}" == sample, row


def test_likelihood_random_vectors():
    n = 0
    for row in read_golden("lik_random.tsv.gz").splitlines():
        t, a, b, ms, e, geno, dp, numbers, prob = row.split("\t")
        gt, dp2, ad, pl = O.genotype_counts(int(a), int(b), t, int(ms), float(e))
        assert (gt, dp2, ad, ",".join(pl)) == (geno, dp, numbers, prob), row
        n += 1
    assert n == 30000


@pytest.mark.parametrize("case_idx", range(32))
def test_quirk_cases(quirks, case_idx):
    if case_idx >= len(quirks["cases"]):
        pytest.skip("no such case")
    case = quirks["cases"][case_idx]
    edges = json.loads(quirks["edges"])
    alt = alt_len_from_gfa_text(quirks["gfa"])
    lines = case["gaf"].splitlines(True)
    if case["rc"] == 0:
        got = O.dumps_informative(O.filter_alignments(lines, edges, alt))
        assert got == case["json"], case["name"]
    else:
        with pytest.raises(O.OracleInputError):
            O.filter_alignments(lines, edges, alt)


def test_quirk_case_count(quirks):
    assert 25 <= len(quirks["cases"]) <= 32


def test_c1_filter_and_genotype_byte_equal():
    edges = json.loads(read_golden("c1_svs_edges.json"))
    alt = alt_len_from_gfa_text(read_golden("c1.gfa.gz"))
    assert len(edges) == 118 and len(alt) == 11
    lines = read_golden("c1.gaf.gz").splitlines(True)
    d = O.filter_alignments(lines, edges, alt)
    assert O.dumps_informative(d) == read_golden("c1_informative_aln.json.gz")
    vcf_lines = read_golden("c1.vcf").splitlines(True)
    text, n = O.genotype_vcf(O.hit_counts(d), vcf_lines)
    assert text == read_golden("c1_genotype.vcf")
    assert f"Genotyped svs: {n}\n" == read_golden("c1_stdout.txt")
    text2, _ = O.genotype_vcf(O.hit_counts(d), vcf_lines, 40, 0.001)
    assert text2 == read_golden("c1_genotype_ms40_e1e-3.vcf")


@pytest.mark.parametrize("tag", ["s2", "s3", "s4"])
def test_scaled_configs_byte_equal(tag):
    edges = json.loads(read_golden(f"{tag}_svs_edges.json.gz"))
    alt = alt_len_from_gfa_text(read_golden(f"{tag}.gfa.gz"))
    lines = read_golden(f"{tag}.gaf.gz").splitlines(True)
    d = O.filter_alignments(lines, edges, alt)
    js = O.dumps_informative(d)
    assert hashlib.sha256(js.encode()).hexdigest() == read_golden(f"{tag}_informative_aln.sha256").strip()
    assert {k: list(v) for k, v in O.hit_counts(d).items()} == json.loads(read_golden(f"{tag}_counts.json.gz"))
    text, n = O.genotype_vcf(O.hit_counts(d), read_golden(f"{tag}.vcf.gz").splitlines(True))
    assert text == read_golden(f"{tag}_genotype.vcf.gz")
    assert f"Genotyped svs: {n}\n" == read_golden(f"{tag}_stdout.txt")


def test_oracle_on_damaged_lines_matches_the_reference():
    """tests/golden/fuzz_lines.json.gz: 2500 randomly damaged GAF lines, each run alone as a file through
    the unmodified reference filter (tests/golden/make_fuzz.py).  The oracle must raise exactly where
    the reference exits with status 1, and count the same hits elsewhere.  A carriage return is a line
    end for the reference (text mode): O.text_mode_lines."""
    import json
    edges = json.loads(read_golden("c1_svs_edges.json"))
    alt = alt_len_from_gfa_text(read_golden("c1.gfa.gz"))
    fx = json.loads(read_golden("fuzz_lines.json.gz"))
    cases = fx["cases"]
    assert len(cases) == 2500 and 500 < sum(c["rc"] for c in cases) < 2000
    for c in cases:
        try:
            got = {}
            for line in O.text_mode_lines(c["line"]):
                for sv, allele in O.record_hits(line, edges, alt):
                    got.setdefault(sv, [0, 0])[allele] += 1
            raised = False
        except Exception:
            raised = True
        assert raised == bool(c["rc"]), c["line"]
        if not raised:
            assert got == c["counts"], c["line"]
    # a file with CR LF line ends: same JSON as the reference, "\n" stored
    d = O.filter_alignments(O.text_mode_lines(fx["crlf"]["gaf"]), edges, alt)
    assert O.dumps_informative(d) == fx["crlf"]["json"]


def test_oracle_on_damaged_vcfs_matches_the_reference():
    """tests/golden/fuzz_vcf.json.gz: 800 small VCFs with one damaged body line each, run through the
    unmodified predict-genotype.py with the c1 informative_aln.json.  Exit status, output text and the
    'Genotyped svs' line must agree."""
    import json
    d = json.loads(read_golden("c1_informative_aln.json.gz"))
    counts = {k: (len(v[0]), len(v[1])) for k, v in d.items()}
    cases = json.loads(read_golden("fuzz_vcf.json.gz"))
    assert len(cases) == 800 and 100 < sum(c["rc"] for c in cases) < 600
    for c in cases:
        try:
            text, n = O.genotype_vcf(counts, O.text_mode_lines(c["vcf"]), c["ms"])
            raised = False
        except Exception:
            raised = True
        assert raised == bool(c["rc"]), c["vcf"][-400:]
        if not raised:
            assert text == c["out"], c["vcf"][-400:]
            assert f"Genotyped svs: {n}\n" == c["stdout"]


def test_oracle_on_damaged_link_tables_matches_the_reference():
    """tests/golden/fuzz_edges.json: 300 svs_edges.json variants with damaged entries (odd allele values,
    sv ids without ':', wrong shapes; regenerated here from the same seed by make_fuzz.damaged_edges),
    each run through the unmodified reference filter on 150 c1 lines: exit status and JSON hash."""
    import hashlib
    import importlib.util
    import json
    import os
    spec = importlib.util.spec_from_file_location("make_fuzz", os.path.join(os.path.dirname(__file__), "golden", "make_fuzz.py"))
    mf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mf)
    alt = alt_len_from_gfa_text(read_golden("c1.gfa.gz"))
    lines = mf.edges_gaf_lines()
    want = json.loads(read_golden("fuzz_edges.json"))
    texts = mf.damaged_edges(len(want))
    assert len(want) == 300 and 50 < sum(c["rc"] for c in want) < 290
    for text, c in zip(texts, want):
        try:
            js = O.dumps_informative(O.filter_alignments(lines, json.loads(text), alt))
            raised = False
        except Exception:
            raised = True
        assert raised == bool(c["rc"])
        if not raised:
            assert hashlib.sha256(js.encode()).hexdigest() == c["sha256"]


def test_oracle_gfa_loader_on_damaged_gfas_matches_the_reference(tmp_path):
    """tests/golden/fuzz_gfa.json.gz (see tests/test_capi_host.py): the oracle's loader must raise where
    the reference's does and build the same alt_node_len dictionary elsewhere."""
    import json
    cases = json.loads(read_golden("fuzz_gfa.json.gz"))
    p = tmp_path / "g.gfa"
    for c in cases:
        with open(p, "w", newline="") as fh:
            fh.write(c["gfa"])
        try:
            got = O.load_alt_node_len(str(p))
            raised = False
        except Exception:
            raised = True
        assert raised == bool(c["rc"]), c["gfa"]
        if not raised:
            assert got == c["alt"], c["gfa"]
