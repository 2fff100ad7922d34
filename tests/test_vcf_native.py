"""csrc/vcf.cpp (svjg_vcf_*): the VCF keys, gate inputs and output text of the genotype stage, built by
the library for the whole file.  No GPU here: the likelihood kernel is replaced by the oracle's
genotype_counts() (test stand-in only), so what is checked is the host side around the kernel —
against the reference's own outputs (golden files and the 800 damaged VCFs of fuzz_vcf.json.gz) and
against the line-by-line Python statement of the same rules (svjg/genotype.py::parse_vcf / format_vcf)."""
import json

import numpy as np
import pytest

from conftest import read_golden
from oracle import svjg_oracle as O
from svjg import capi, genotype

TYPES = ("DEL", "INS", "INV", "BND")


def oracle_arrays(v, counts, ms=3, e=0.00005):
    """gt / flags / ad2 / pl as the genotype kernel returns them, from the oracle (stand-in)."""
    n = v.n
    gt, fl = np.full(n, 3, np.uint8), np.zeros(n, np.uint8)
    ad, pl = np.zeros((n, 2), np.uint32), np.zeros((n, 3), np.int64)
    ty = v.svtype
    for i in range(n):
        key = v.key(i)
        if ty[i] == 255 or ty[i] & 0x80 or key not in counts:
            continue
        g, _dp, a, p = O.genotype_counts(counts[key][0], counts[key][1], TYPES[ty[i] & 3], ms, e)
        a1, a2 = a.split(",")
        fl[i] = capi.GT_GENOTYPED | (capi.GT_HALVED_0 if "." in a1 else 0) | (capi.GT_HALVED_1 if "." in a2 else 0)
        gt[i] = ("0/0", "0/1", "1/1", "./.").index(g)
        ad[i] = (round(float(a1) * 2), round(float(a2) * 2))
        pl[i] = [int(x) for x in p]
    return gt, fl, ad, pl


@pytest.mark.parametrize("tag", ["c1", "s2", "s3", "s4"])
def test_golden_outputs_of_the_reference(tag):
    vcf = read_golden("c1.vcf" if tag == "c1" else f"{tag}.vcf.gz")
    if tag == "c1":
        counts = O.hit_counts(json.loads(read_golden("c1_informative_aln.json.gz")))
    else:
        counts = {k: tuple(v) for k, v in json.loads(read_golden(f"{tag}_counts.json.gz")).items()}
    want = read_golden("c1_genotype.vcf" if tag == "c1" else f"{tag}_genotype.vcf.gz")
    for src in (vcf.encode(), vcf.replace("\n", "\r\n").encode(), vcf.splitlines(True), np.frombuffer(vcf.encode(), np.uint8)):
        v = genotype.NativeVcf.from_input(src)
        text, n = v.format(*oracle_arrays(v, counts))
        assert text == want
    header, recs = genotype.parse_vcf(vcf.splitlines(True))
    assert [v.key(i) for i in range(v.n)] == [r[2] for r in recs]
    assert v.svtype.tolist() == [r[1] for r in recs]


def test_damaged_vcfs_match_the_reference():
    """tests/golden/fuzz_vcf.json.gz: 800 small VCFs with a damaged body line, run through the unmodified
    predict-genotype.py with the c1 informative_aln.json (tests/golden/make_fuzz.py vcf): the library
    must refuse exactly the files the reference exits 1 on and print the same text for the others."""
    d = json.loads(read_golden("c1_informative_aln.json.gz"))
    counts = O.hit_counts(d)
    aln = genotype.AlnCounts.from_memory(read_golden("c1_informative_aln.json.gz"))
    cases = json.loads(read_golden("fuzz_vcf.json.gz"))
    n_native = 0
    import io
    for c in cases:
        data = c["vcf"].encode()
        as_lines = list(io.StringIO(c["vcf"], newline=None))          # what open(path).readlines() gives
        try:
            v = genotype.NativeVcf.from_input(data)
        except genotype.VcfError:
            assert c["rc"] == 1, c["vcf"][-300:]
            with pytest.raises(genotype.VcfError):
                genotype.NativeVcf.from_input(as_lines)
            n_native += 1
            continue
        v2 = genotype.NativeVcf.from_input(as_lines)
        assert (v is None) == (v2 is None)
        if v is not None:
            assert v2.n == v.n and v2.svtype.tolist() == v.svtype.tolist() and [v2.key(i) for i in range(v.n)] == [v.key(i) for i in range(v.n)]
        if v is None:                                # declined: the Python statement takes over
            assert not c["vcf"].isascii() or any(len(x) > 18 for x in c["vcf"].replace(";", "\t").replace("=", "\t").split("\t") if x.isdigit())
            continue
        n_native += 1
        assert c["rc"] == 0, c["vcf"][-300:]
        text, n = v.format(*oracle_arrays(v, counts, c["ms"]))
        assert text == c["out"], c["vcf"][-300:]
        assert f"Genotyped svs: {n}\n" == c["stdout"]
        # the index against an informative_aln.json: present keys get bit 6
        idx, ty = v.index_counts(aln)
        for i in range(v.n):
            k = v.key(i)
            j = None if k is None else aln.find(k)
            assert idx[i] == (capi.NO_SV if j is None else j)
            assert ty[i] == (v.svtype[i] | (0x40 if j is not None and v.svtype[i] != 255 else 0))
    assert n_native > 780


def test_native_and_python_statements_agree():
    """Same keys, codes, errors and text on spellings picked to separate the two if they differed."""
    head = "##fileformat=VCFv4.2\n##FORMAT=<ID=XX>\n##x=1\n#CHROM\tPOS\n"
    bodies = [
        "1\t100\ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL;END=149\n",                   # 49: short
        "1\t100\ta\tN\t<DEL>\t.\t.\tEND=150;SVTYPE=DEL\tGT\t0/1\n",          # extra columns cut
        "1\t 1_00 \ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL;END=+0200\n",              # int() spellings
        "1\t100\ta\tN\t<INV>\t.\t.\tX=1;END=40;SVTYPE=INV;Y\n",             # negative length
        "1\t100\ta\tN\tACGT\t.\t.\tSVTYPE=INS\n2\t100\tb\tN\t" + "A" * 50 + "\t.\t.\tSVTYPE=INS;END=5\n",   # k per POS string
        "1\t0100\ta\tN\t" + "A" * 60 + "\t.\t.\tSVTYPE=INS\n",
        "1\t7\ta\tN\tN[2:55[\t.\t.\tSVTYPE=BND\n1\t7\ta\tN\t]2:55]N\t.\t.\tSVTYPE=BND\n1\t7\ta\tN\t[2:55[N\t.\t.\tSVTYPE=BND\n",
        "1\t7\ta\tN\tN]2:55]\t.\t.\tSVTYPE=BND\n1\t7\ta\tN\t<BND>\t.\t.\tSVTYPE=BND\n1\t7\ta\tN\tN[[x\t.\t.\tSVTYPE=BND\n",
        "1\t7\ta\tN\t<DUP>\t.\t.\tSVTYPE=DUP;END=9\n1\t7\ta\tN\t<X>\t.\t.\tEND=9\n",
        "1\t7\ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL;SVTYPE=INS;END=99\n",           # text between the first two SVTYPE=
        "1\t7\ta\tN\t<DEL>\t.\t.\tEND=99;SVTYPE=DEL;X;SVTYPE=DEL\n",
        "1\t7\ta\tN\t<DEL>\t.\t.\tEND=99;XEND=5;SVTYPE=DEL",                 # no final newline
        "#odd\n", "\n", "1\t7\ta\tN\t<DEL>\t.\t.\n", "1\t7\ta\tN\t<DEL>\t.\t.\tSVTYPE\n", "1\t7\ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL\n",
        "1\tx\ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL;END=9\n", "1\t7\ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL;END=9_\n", "1\t7\ta\tN\tN[\t.\t.\tSVTYPE=BND\n",
        "1\t7\ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL;END=1__0\n", "1\t-7\ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL;END=-60\n",
    ]
    rng = np.random.default_rng(5)
    for body in bodies:
        text = head + body + "##tail=1\n"
        lines = genotype._as_lines(text.encode())
        try:
            header, recs = genotype.parse_vcf(lines)
            err = None
        except genotype.VcfError as exc:
            err = exc
        for src in (text.encode(), lines):
            if err is not None:
                with pytest.raises(genotype.VcfError):
                    genotype.NativeVcf.from_input(src)
                continue
            v = genotype.NativeVcf.from_input(src)
            assert v is not None and v.n == len(recs), body
            assert [v.key(i) for i in range(v.n)] == [r[2] for r in recs], body
            assert v.svtype.tolist() == [r[1] for r in recs], body
            n = v.n
            gt = rng.integers(0, 4, n).astype(np.uint8)
            fl = rng.integers(0, 8, n).astype(np.uint8)
            ad = rng.integers(0, 5000, (n, 2)).astype(np.uint32)
            pl = rng.integers(-10, 10**15, (n, 3)).astype(np.int64)
            assert v.format(gt, fl, ad, pl) == genotype.format_vcf(header, recs, gt, fl, ad, pl), body


def test_big_vcf_text_filled_by_several_threads():
    """svjg_vcf_format fills texts of more than 4 MB with several threads (record ranges of equal bytes): the
    same bytes as the line-by-line statement, header lines between and behind the records included."""
    rng = np.random.default_rng(11)
    parts = ["##fileformat=VCFv4.2\n##FORMAT=<ID=XX>\n#CHROM\tPOS\n"]
    for i in range(6000):
        if i in (1, 2500, 5999):
            parts.append(f"##odd header in front of record {i}\n")
        if i % 3 == 0:
            parts.append(f"chr{i % 7}\t{1000 + i}\tv{i}\tN\t{'ACGT' * int(rng.integers(200, 1200))}\t.\t.\tSVTYPE=INS;END={1001 + i}\n")
        else:
            parts.append(f"chr{i % 7}\t{1000 + i}\tv{i}\tN\t<DEL>\t.\t.\tSVTYPE=DEL;END={1100 + i};NOTE={'x' * int(rng.integers(0, 300))}\n")
    parts.append("##behind the last record\n##and one more\n")
    text = "".join(parts)
    assert len(text) > (5 << 20)
    lines = genotype._as_lines(text.encode())
    header, recs = genotype.parse_vcf(lines)
    v = genotype.NativeVcf.from_input(text.encode())
    n = v.n
    assert n == 6000 == len(recs)
    gt = rng.integers(0, 4, n).astype(np.uint8)
    fl = rng.integers(0, 8, n).astype(np.uint8)
    ad = rng.integers(0, 5000, (n, 2)).astype(np.uint32)
    pl = rng.integers(-10, 10**15, (n, 3)).astype(np.int64)
    assert v.format(gt, fl, ad, pl) == genotype.format_vcf(header, recs, gt, fl, ad, pl)


def test_declined_spellings_fall_to_the_python_statement():
    head = "#CHROM\n"
    assert genotype.NativeVcf.from_input((head + "é\t7\ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL;END=99\n").encode()) is None
    assert genotype.NativeVcf.from_input((head + "1\t7\ta\tN\t<DEL>\t.\t.\tSVTYPE=DEL;END=" + "9" * 19 + "\n").encode()) is None
    assert genotype.NativeVcf.from_input(["a\nb\n"]) is None            # not lines as readlines() gives them
    assert genotype.NativeVcf.from_input(["a", "b\n"]) is None
    assert genotype.NativeVcf.from_input(["##h\n", ""]) is None            # an empty string is a (bad) line for the line-by-line rules
    assert genotype.NativeVcf.from_input([]).n == 0
    assert genotype.NativeVcf.from_input(b"").format(np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros((0, 2), np.uint32),
                                                     np.zeros((0, 3), np.int64)) == ("", 0)


def stand_in_genotype_host(counts, sv_index, svtype, min_support=3, e=0.00005):
    """svjg_genotype_host's contract (include/svjg.h) computed by the oracle: for the CPU test of the
    command-line wiring only."""
    n = len(sv_index)
    gt, fl = np.full(n, 3, np.uint8), np.zeros(n, np.uint8)
    ad, pl = np.zeros((n, 2), np.uint32), np.zeros((n, 3), np.int64)
    for i in range(n):
        ty = int(svtype[i])
        if ty == 255 or (ty & 0x3F) > 3 or ty & 0x80 or sv_index[i] == capi.NO_SV:
            continue
        n0, n1 = (int(x) for x in counts[sv_index[i]])
        if not (ty & 0x40) and n0 == 0 and n1 == 0:
            continue
        g, _dp, a, p = O.genotype_counts(n0, n1, TYPES[ty & 3], min_support, e)
        a1, a2 = a.split(",")
        fl[i] = capi.GT_GENOTYPED | (capi.GT_HALVED_0 if "." in a1 else 0) | (capi.GT_HALVED_1 if "." in a2 else 0)
        gt[i] = ("0/0", "0/1", "1/1", "./.").index(g)
        ad[i] = (round(float(a1) * 2), round(float(a2) * 2))
        pl[i] = [int(x) for x in p]
    return gt, fl, ad, pl


@pytest.mark.parametrize("compressed", [False, True])
def test_predict_genotype_command_line_wiring(tmp_path, monkeypatch, capsys, compressed):
    """predict-genotype.py's front-end around the kernel, with the kernel call replaced by the stand-in
    above and the device start-up skipped: files, flags and the 'Genotyped svs' line as the reference
    (golden c1 outputs, defaults and -ms 40 -e 0.001)."""
    import gzip
    from svjg import cli
    monkeypatch.setattr(cli, "_start_device", lambda: (lambda: None))
    monkeypatch.setattr(genotype, "genotype_host", stand_in_genotype_host)
    monkeypatch.chdir(tmp_path)
    (tmp_path / "a.json").write_text(read_golden("c1_informative_aln.json.gz"))
    vcf = read_golden("c1.vcf").encode()
    name = "in.vcf.gz" if compressed else "in.vcf"
    (tmp_path / name).write_bytes(gzip.compress(vcf) if compressed else vcf)
    assert cli.genotype_main(["-d", "a.json", "-v", name, "-o", "out.vcf"]) == 0
    assert (tmp_path / "out.vcf").read_text() == read_golden("c1_genotype.vcf")
    assert capsys.readouterr().out == read_golden("c1_stdout.txt")
    assert cli.genotype_main(["-d", "a.json", "-v", name, "-ms", "40", "-e", "0.001"]) == 0
    assert (tmp_path / "genotype_results.txt").read_text() == read_golden("c1_genotype_ms40_e1e-3.vcf")
    # a line the reference raises on: exit status 1, the output file exists and is empty (opened first, :92)
    (tmp_path / "bad.vcf").write_bytes(vcf + b"chr1\t5\n")
    with pytest.raises(SystemExit) as exc:
        cli.genotype_main(["-d", "a.json", "-v", "bad.vcf", "-o", "bad_out.vcf"])
    assert exc.value.code == 1 and (tmp_path / "bad_out.vcf").read_bytes() == b""
    # non-ASCII VCF: the Python statement of the rules writes the same kind of file
    (tmp_path / "u.vcf").write_bytes(vcf.replace(b"##fileformat", "##note=é\n##fileformat".encode(), 1))
    assert cli.genotype_main(["-d", "a.json", "-v", "u.vcf", "-o", "u_out.vcf"]) == 0
    assert (tmp_path / "u_out.vcf").read_text() == "##note=é\n" + read_golden("c1_genotype.vcf")


def test_damaged_informative_aln_json_matches_the_reference():
    """tests/golden/fuzz_json.json: 600 informative_aln.json variants (odd values, escapes in keys, repeated
    keys, byte damage; regenerated here from the seed by make_fuzz.damaged_jsons), each run through the
    unmodified predict-genotype.py with c1.vcf.  The library's JSON reader + VCF side (kernel: stand-in)
    must stop exactly where the reference exits 1 and write the same file elsewhere."""
    import hashlib
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_fuzz", os.path.join(os.path.dirname(__file__), "golden", "make_fuzz.py"))
    mf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mf)
    want = json.loads(read_golden("fuzz_json.json"))
    texts = mf.damaged_jsons(len(want))
    vcf = read_golden("c1.vcf").encode()
    n_stop = 0
    for text, c in zip(texts, want):
        assert hashlib.sha256(text.encode()).hexdigest()[:12] == c["in"]          # same inputs as when the fixture was made
        try:
            aln = genotype.AlnCounts.from_memory(text)
            v = genotype.NativeVcf.from_input(vcf)
            idx, ty = v.index_counts(aln)
            counts = aln.counts if aln.num else np.zeros((1, 2), np.uint32)
            out = v.format(*stand_in_genotype_host(counts, idx, ty))
            failed = False
        except (capi.SvjgError, genotype.VcfError):
            failed = True
        assert failed == bool(c["rc"]), text[:300]
        n_stop += failed
        if not failed:
            assert hashlib.sha256(out[0].encode()).hexdigest() == c["sha256"], text[:300]
            assert f"Genotyped svs: {out[1]}\n" == c["stdout"]
    assert 100 < n_stop < 500


# ---- the same fixtures through the real genotype kernel (torch-free path of the command lines) ----
@pytest.mark.gpu
def test_golden_and_damaged_vcfs_on_the_device():
    aln = genotype.AlnCounts.from_memory(read_golden("c1_informative_aln.json.gz"))
    vcf = read_golden("c1.vcf")
    for src in (vcf.encode(), vcf.splitlines(True)):
        assert genotype.genotype_vcf_from_json(aln, src) == (read_golden("c1_genotype.vcf"), int(read_golden("c1_stdout.txt").split()[-1]))
        assert genotype.genotype_vcf_from_json(aln, src, 40, 0.001)[0] == read_golden("c1_genotype_ms40_e1e-3.vcf")
    cases = json.loads(read_golden("fuzz_vcf.json.gz"))
    n_err = 0
    for c in cases:
        try:
            text, n = genotype.genotype_vcf_from_json(aln, c["vcf"].encode(), c["ms"], 0.00005)
            failed = False
        except (genotype.VcfError, capi.SvjgError, ValueError):
            failed = True
        assert failed == bool(c["rc"]), c["vcf"][-300:]
        n_err += failed
        if not failed:
            assert text == c["out"], c["vcf"][-300:]
            assert f"Genotyped svs: {n}\n" == c["stdout"]
    assert n_err > 100


@pytest.mark.gpu
def test_damaged_informative_aln_json_on_the_device():
    import hashlib
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_fuzz", os.path.join(os.path.dirname(__file__), "golden", "make_fuzz.py"))
    mf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mf)
    want = json.loads(read_golden("fuzz_json.json"))
    vcf = read_golden("c1.vcf").encode()
    for text, c in zip(mf.damaged_jsons(len(want)), want):
        try:
            out = genotype.genotype_vcf_from_json(genotype.AlnCounts.from_memory(text), vcf)
            failed = False
        except (capi.SvjgError, genotype.VcfError):
            failed = True
        assert failed == bool(c["rc"]), text[:300]
        if not failed:
            assert hashlib.sha256(out[0].encode()).hexdigest() == c["sha256"], text[:300]
            assert f"Genotyped svs: {out[1]}\n" == c["stdout"]
