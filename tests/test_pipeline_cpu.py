"""svjedi-graph.py's wiring end to end on the CPU: the drop-in construct-graph.py runs for real, minigraph
is a stub that prints a synthetic GAF, and the two GPU calls (filter, genotype) are replaced by oracle
stand-ins — so every file the orchestrator handles (svjedi-graph.py:84-128) is produced and checked
against the oracle's own run of the same stages.  The GPU version of this test is
tests/test_cli_and_json.py::test_pipeline_front_end_with_stub_tools."""
import io
import json
import os
from collections import OrderedDict

import numpy as np
import pytest

from conftest import PKG, alt_len_from_gfa_text
from oracle import svjg_oracle as O
from svjg import alnfilter, cli, genotype, graphgen, synth
from test_vcf_native import stand_in_genotype_host


@pytest.mark.parametrize("streamed", [False, True])
def test_orchestrator_files_and_order(tmp_path, monkeypatch, capfd, streamed):
    rng = np.random.Generator(np.random.PCG64(3))
    chrom_len = OrderedDict((("chr1", 60000), ("chr2", 40000)))
    seqs = OrderedDict((c, rng.choice(np.frombuffer(b"ACGT", dtype="S1"), size=n).tobytes().decode()) for c, n in chrom_len.items())
    (tmp_path / "ref.fa").write_text("".join(f">{c} test\n{s}\n" for c, s in seqs.items()))
    rows = synth.catalogue("mix", 60, chrom_len, 11, ins_max=300, del_max=900)
    vcf = synth.vcf_text(rows)
    (tmp_path / "in.vcf").write_text(vcf)
    g = graphgen.build_graph(chrom_len, rows)
    gaf = synth.simulate_gaf(g, 1500, seed=5, mean_len=4000)
    (tmp_path / "reads.gaf").write_text(gaf)
    bindir = tmp_path / "bin"
    bindir.mkdir()
    mg = bindir / "minigraph"
    mg.write_text(f"#!/bin/sh\ncat {tmp_path}/reads.gaf\n")
    mg.chmod(0o755)
    monkeypatch.setenv("PATH", f"{bindir}:{os.environ['PATH']}")
    monkeypatch.delenv("SVJG_CONSTRUCT_GRAPH", raising=False)
    (tmp_path / "reads.fq").write_text("")
    prefix = str(tmp_path / "run")

    def load_tables(pfx, gfa_file, device_ready=None):               # host half of cli._load_tables
        return alnfilter.Tables.load(pfx + "_svs_edges.json", gfa_file)

    def filter_host(tables, gaf_bytes, d_over=100, **kw):             # svjg_filter_host's contract, by the oracle
        edges = json.load(open(prefix + "_svs_edges.json"))
        alt = alt_len_from_gfa_text(open(prefix + ".gfa").read())
        counts = np.zeros((tables.num_sv, 2), np.uint32)
        sv2, off, ln = [], [], []
        pos = 0
        for line in bytes(alnfilter._as_u8(gaf_bytes)).decode().splitlines(True):
            for sv, allele in O.record_hits(line, edges, alt):
                i = tables.find_sv(sv)
                counts[i, allele] += 1
                sv2.append(2 * i + allele)
                off.append(pos)
                ln.append(len(line))
            pos += len(line)
        stats = {"n_hits": len(sv2), "n_checks": 0}
        return alnfilter.FilterResult(counts, stats, np.array(sv2, np.uint32), np.array(off, np.uint64), np.array(ln, np.uint32))

    monkeypatch.setattr(cli, "_load_tables", load_tables)
    monkeypatch.setattr(alnfilter, "filter_host", filter_host)
    monkeypatch.setattr(alnfilter, "filter_json_begin", lambda tables, gaf, **kw: None)   # no device: the host route
    monkeypatch.setattr(alnfilter, "read_file_pinned", lambda path: np.fromfile(path, dtype=np.uint8))
    monkeypatch.setattr(genotype, "genotype_host", stand_in_genotype_host)
    monkeypatch.chdir(tmp_path)
    if streamed:                                              # SVJG_STREAM=1: mapping and filtering overlapped (row N3)
        monkeypatch.setenv("SVJG_STREAM", "1")
    else:
        monkeypatch.delenv("SVJG_STREAM", raising=False)
    assert cli.pipeline_main(PKG, ["-v", "in.vcf", "-r", "ref.fa", "-q", "reads.fq", "-p", prefix]) == 0
    out = capfd.readouterr().out

    # the graph stage wrote the reference's three files (this builder, with sequences)
    g2 = graphgen.build_graph(chrom_len, rows, seqs)
    buf = io.StringIO()
    g2.write_gfa(buf)
    assert open(prefix + ".gfa").read() == buf.getvalue()
    assert open(prefix + "_svs_edges.json").read() == g2.edges_json()
    assert open(prefix + "_ignored_svs.txt").read() == g2.ignored_text()
    assert open(prefix + ".gaf").read() == gaf
    # stages 3-4 against the oracle's own run
    d = O.filter_alignments(gaf.splitlines(True), g2.link_sv_as_json(), alt_len_from_gfa_text(buf.getvalue()))
    assert open(prefix + "_informative_aln.json").read() == O.dumps_informative(d)
    text, n = O.genotype_vcf(O.hit_counts(d), vcf.splitlines(True))
    assert n > 10
    assert open(prefix + "_genotype.vcf").read() == text
    assert out == ("Constructing variation graph...\nMapping reads on graph...\nFiltering alignment file...\n"
                   f"Genotyping SVs...\nGenotyped svs: {n}\n")


def test_filter_command_line_wiring(tmp_path, monkeypatch):
    """filter-alignments.py's front-end around the GPU call (stand-in): -p/-o file naming, gzip input,
    CR LF line ends read as text mode reads them, and the two flag defects kept from the reference."""
    import gzip
    import pytest
    from conftest import read_golden
    edges_text, gfa_text = read_golden("c1_svs_edges.json"), read_golden("c1.gfa.gz")
    edges, alt = json.loads(edges_text), alt_len_from_gfa_text(gfa_text)
    lines = [l for l in read_golden("c1.gaf.gz").splitlines(True) if "cg:Z:" not in l][:300]
    want = O.dumps_informative(O.filter_alignments(lines, edges, alt))
    (tmp_path / "p_svs_edges.json").write_text(edges_text)
    (tmp_path / "p.gfa").write_text(gfa_text)
    (tmp_path / "p.gaf").write_text("".join(lines))
    (tmp_path / "crlf.gaf.gz").write_bytes(gzip.compress("".join(lines).replace("\n", "\r\n").encode()))
    (tmp_path / "out").mkdir()

    def filter_host(tables, gaf_bytes, d_over=100, **kw):
        counts = np.zeros((tables.num_sv, 2), np.uint32)
        sv2, off, ln = [], [], []
        pos = n_checks = 0
        for line in bytes(alnfilter._as_u8(gaf_bytes)).decode().splitlines(True):
            for sv, allele in O.record_hits(line, edges, alt):
                i = tables.find_sv(sv)
                counts[i, allele] += 1
                sv2.append(2 * i + allele)
                off.append(pos)
                ln.append(len(line))
                n_checks += 1
            pos += len(line)
        return alnfilter.FilterResult(counts, {"n_hits": len(sv2), "n_checks": n_checks}, np.array(sv2, np.uint32),
                                      np.array(off, np.uint64), np.array(ln, np.uint32))

    monkeypatch.setattr(cli, "_start_device", lambda: (lambda: None))
    monkeypatch.setattr(cli, "_load_tables", lambda pfx, gfa, ready=None: alnfilter.Tables.load(pfx + "_svs_edges.json", gfa))
    monkeypatch.setattr(alnfilter, "filter_host", filter_host)
    monkeypatch.setattr(alnfilter, "filter_json_begin", lambda tables, gaf, **kw: None)   # no device: the host route
    # page-locking needs the CUDA runtime: keep the wrapper, skip the registration
    monkeypatch.setattr(alnfilter.RegisteredBytes, "__init__", lambda self, array: (setattr(self, "array", array), setattr(self, "_reg", False)) and None)
    monkeypatch.chdir(tmp_path)
    assert cli.filter_main(["-a", "p.gaf", "-g", "p.gfa", "-p", "p"]) == 0
    assert (tmp_path / "p_informative_aln.json").read_text() == want
    assert cli.filter_main(["-a", "crlf.gaf.gz", "-g", "p.gfa", "-p", "p", "-o", "out"]) == 0
    assert (tmp_path / "out" / "p_informative_aln.json").read_text() == want
    # the same records through a pipe: filtered as a stream (row N3)
    import threading
    fifo = str(tmp_path / "gaf.fifo")
    os.mkfifo(fifo)
    th = threading.Thread(target=lambda: open(fifo, "wb").write("".join(lines).replace("\n", "\r\n").encode()))
    th.start()
    (tmp_path / "piped_svs_edges.json").write_text(edges_text)
    assert cli.filter_main(["-a", fifo, "-g", "p.gfa", "-p", "piped"]) == 0
    th.join()
    assert (tmp_path / "piped_informative_aln.json").read_text() == want
    with pytest.raises(SystemExit) as exc:                       # no -p: the reference never finds its link table (:95)
        cli.filter_main(["-a", "p.gaf", "-g", "p.gfa"])
    assert exc.value.code == 1
    with pytest.raises(SystemExit) as exc:                       # -O: TypeError at the first overlap test (:269)
        cli.filter_main(["-a", "p.gaf", "-g", "p.gfa", "-p", "p", "-O", "50"])
    assert exc.value.code == 1
    with pytest.raises(SystemExit) as exc:
        cli.filter_main(["-a", "missing.gaf", "-g", "p.gfa", "-p", "p"])
    assert exc.value.code == 1


def test_streamed_input_equals_the_whole_file(monkeypatch):
    """alnfilter.filter_stream (row N3): segments of whole lines, CR LF / CR translated across read
    boundaries, offsets moved to their place — against one pass over the whole (translated) input.  The GPU
    call is the oracle stand-in; what is tested is the cutting and the bookkeeping."""
    import io
    import threading
    from conftest import read_golden
    edges, alt = json.loads(read_golden("c1_svs_edges.json")), alt_len_from_gfa_text(read_golden("c1.gfa.gz"))
    lines = [l for l in read_golden("c1.gaf.gz").splitlines(True) if "cg:Z:" not in l][:400]
    text = "".join(l if i % 3 else l.replace("\n", "\r\n") for i, l in enumerate(lines))
    text = text[:-1]                                            # no line end after the last record
    raw = text.encode()
    t = alnfilter.Tables.from_memory(read_golden("c1_svs_edges.json"), read_golden("c1.gfa.gz"))

    def filter_host(tables, gaf_bytes, d_over=100, **kw):
        counts = np.zeros((tables.num_sv, 2), np.uint32)
        sv2, off, ln = [], [], []
        pos = 0
        for line in bytes(alnfilter._as_u8(gaf_bytes)).decode().splitlines(True):
            for sv, allele in O.record_hits(line, edges, alt):
                i = tables.find_sv(sv)
                counts[i, allele] += 1
                sv2.append(2 * i + allele)
                off.append(pos)
                ln.append(len(line))
            pos += len(line)
        return alnfilter.FilterResult(counts, {"n_hits": len(sv2), "n_records": pos and 1}, np.array(sv2, np.uint32),
                                      np.array(off, np.uint64), np.array(ln, np.uint32))

    monkeypatch.setattr(alnfilter, "filter_host", filter_host)

    monkeypatch.setattr(alnfilter, "filter_json_begin", lambda tables, gaf, **kw: None)   # no device: the host route
    whole_bytes = bytes(alnfilter._as_u8(alnfilter.translate_newlines(np.frombuffer(raw, np.uint8))))
    whole = filter_host(t, whole_bytes)
    assert whole.stats["n_hits"] > 50
    want = sorted(zip(whole.hit_off.tolist(), whole.hit_sv2.tolist(), whole.hit_len.tolist()))
    for chunk in (7, 64, 1000, 5000, 1 << 20):
        res, gaf = alnfilter.filter_stream(t, io.BytesIO(raw), chunk_bytes=chunk)
        assert gaf.tobytes() == whole_bytes, chunk
        assert (res.counts == whole.counts).all() and res.n_hits == whole.stats["n_hits"], chunk
        assert sorted(zip(res.hit_off.tolist(), res.hit_sv2.tolist(), res.hit_len.tolist())) == want, chunk
    # a real pipe that delivers odd-sized pieces
    r, w = os.pipe()

    def feed():
        with os.fdopen(w, "wb", buffering=0) as fh:
            for i in range(0, len(raw), 777):
                fh.write(raw[i:i + 777])
    th = threading.Thread(target=feed)
    th.start()
    with os.fdopen(r, "rb", buffering=0) as fh:                 # raw reads: short counts are normal
        res, gaf = alnfilter.filter_stream(t, fh, chunk_bytes=4096)
    th.join()
    assert gaf.tobytes() == whole_bytes and (res.counts == whole.counts).all()
    assert sorted(zip(res.hit_off.tolist(), res.hit_sv2.tolist(), res.hit_len.tolist())) == want
    empty, gaf = alnfilter.filter_stream(t, io.BytesIO(b""))
    assert gaf.size == 0 and empty.n_hits == 0 and not empty.counts.any()


def test_streamed_orchestrator_appends_and_reports_like_the_reference(tmp_path, monkeypatch, capfd):
    """SVJG_STREAM=1 edge semantics: <prefix>.gaf is appended to (what it already holds is filtered too,
    svjedi-graph.py:100-104), every FASTQ of a comma list is mapped in turn, and exit status 1 of the LAST
    mapper stops the run with the reference's message (:107) after the file has been written."""
    from conftest import read_golden
    edges_text, gfa_text = read_golden("c1_svs_edges.json"), read_golden("c1.gfa.gz")
    edges, alt = json.loads(edges_text), alt_len_from_gfa_text(gfa_text)
    lines = [l for l in read_golden("c1.gaf.gz").splitlines(True) if "cg:Z:" not in l]
    old, new = "".join(lines[:40]), "".join(lines[40:200])
    prefix = str(tmp_path / "run")
    stub = tmp_path / "construct_stub.py"                       # the graph files of c1 (its FASTA is not in this repository)
    (tmp_path / "c1.gfa").write_text(gfa_text)
    (tmp_path / "c1_edges.json").write_text(edges_text)
    stub.write_text("import shutil, sys\nout = sys.argv[sys.argv.index('-o') + 1]\n"
                    f"shutil.copy({str(tmp_path / 'c1.gfa')!r}, out)\n"
                    f"shutil.copy({str(tmp_path / 'c1_edges.json')!r}, out[:-4] + '_svs_edges.json')\n")
    (tmp_path / "new.gaf").write_text(new)
    bindir = tmp_path / "bin"
    bindir.mkdir()
    mg = bindir / "minigraph"
    mg.write_text(f"#!/bin/sh\ncat {tmp_path}/new.gaf\nexit ${{MG_RC:-0}}\n")
    mg.chmod(0o755)
    monkeypatch.setenv("PATH", f"{bindir}:{os.environ['PATH']}")
    monkeypatch.setenv("SVJG_CONSTRUCT_GRAPH", str(stub))
    monkeypatch.setenv("SVJG_STREAM", "1")
    (tmp_path / "in.vcf").write_text(read_golden("c1.vcf"))

    def filter_host(tables, gaf_bytes, d_over=100, **kw):
        counts = np.zeros((tables.num_sv, 2), np.uint32)
        sv2, off, ln = [], [], []
        pos = 0
        for line in bytes(alnfilter._as_u8(gaf_bytes)).decode().splitlines(True):
            for sv, allele in O.record_hits(line, edges, alt):
                i = tables.find_sv(sv)
                counts[i, allele] += 1
                sv2.append(2 * i + allele)
                off.append(pos)
                ln.append(len(line))
            pos += len(line)
        return alnfilter.FilterResult(counts, {"n_hits": len(sv2)}, np.array(sv2, np.uint32), np.array(off, np.uint64),
                                      np.array(ln, np.uint32))

    monkeypatch.setattr(cli, "_load_tables", lambda pfx, gfa, ready=None: alnfilter.Tables.load(pfx + "_svs_edges.json", gfa))
    monkeypatch.setattr(alnfilter, "filter_host", filter_host)
    monkeypatch.setattr(alnfilter, "filter_json_begin", lambda tables, gaf, **kw: None)   # no device: the host route
    monkeypatch.setattr(genotype, "genotype_host", stand_in_genotype_host)
    monkeypatch.chdir(tmp_path)
    (tmp_path / "run.gaf").write_text(old)                      # left over from an earlier run
    assert cli.pipeline_main(PKG, ["-v", "in.vcf", "-r", "ref.fa", "-q", "a.fq,b.fq", "-p", prefix]) == 0
    capfd.readouterr()
    assert (tmp_path / "run.gaf").read_text() == old + new + new
    want = O.filter_alignments((old + new + new).splitlines(True), edges, alt)
    assert (tmp_path / "run_informative_aln.json").read_text() == O.dumps_informative(want)
    assert (tmp_path / "run_genotype.vcf").read_text() == O.genotype_vcf(O.hit_counts(want), read_golden("c1.vcf").splitlines(True))[0]
    # the last mapper fails: the file is complete, the run stops with the reference's message
    monkeypatch.setenv("MG_RC", "1")
    (tmp_path / "run.gaf").write_text("")
    with pytest.raises(SystemExit) as exc:
        cli.pipeline_main(PKG, ["-v", "in.vcf", "-r", "ref.fa", "-q", "a.fq", "-p", prefix])
    assert str(exc.value.code).startswith("Failed to map the reads on the graph.")
    assert (tmp_path / "run.gaf").read_text() == new
    # a line the reference raises on: mapping completes, then the filter's failure is reported
    monkeypatch.setenv("MG_RC", "0")
    (tmp_path / "new.gaf").write_text(new + "only\tthree\tcolumns\n" + new)
    (tmp_path / "run.gaf").write_text("")
    monkeypatch.setattr(alnfilter, "filter_host", lambda *a, **k: (_ for _ in ()).throw(alnfilter.InputError("GAF line: fewer than 12 columns")))
    monkeypatch.setattr(alnfilter, "filter_json_begin", lambda tables, gaf, **kw: None)   # no device: the host route
    with pytest.raises(SystemExit) as exc:
        cli.pipeline_main(PKG, ["-v", "in.vcf", "-r", "ref.fa", "-q", "a.fq", "-p", prefix])
    assert str(exc.value.code).startswith("Failed to filter the alignments.")
    assert (tmp_path / "run.gaf").read_text() == new + "only\tthree\tcolumns\n" + new
