"""CPU-only checks of libsvjg.so: it loads, exports every symbol include/svjg.h
declares, builds the host tables like the reference's loaders, and writes
informative_aln.json byte-for-byte (hits supplied by the oracle here; the GPU
tests supply them from the kernel)."""
import json
import os
import re

import numpy as np
import pytest

from conftest import ROOT, alt_len_from_gfa_text, read_golden
from oracle import svjg_oracle as O
from svjg import alnfilter, capi


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "svjg.h")).read()
    declared = set(re.findall(r"\b(svjg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(capi.lib, name), f"{name} declared in svjg.h but not exported"
    assert declared == set(capi.EXPORTS)
    assert b"sm_100a" in capi.lib.svjg_version()


@pytest.mark.parametrize("tag", ["c1", "s2", "s3", "s4"])
def test_tables_match_reference_loaders(tag):
    edges_text = read_golden("c1_svs_edges.json" if tag == "c1" else f"{tag}_svs_edges.json.gz")
    gfa_text = read_golden(f"{tag}.gfa.gz")
    t = alnfilter.Tables.from_memory(edges_text, gfa_text)
    d = json.loads(edges_text)
    want_ids = sorted({sv for ents in d.values() for sv, _ in ents})
    assert t.num_links == len(d)
    assert t.sv_ids == want_ids
    assert t.num_alt_nodes == len(alt_len_from_gfa_text(gfa_text))
    assert t.find_sv(want_ids[-1]) == len(want_ids) - 1
    assert t.find_sv("nope:DEL-1-2") is None


def test_tables_reject_bad_json():
    with pytest.raises(capi.SvjgError):
        alnfilter.Tables.from_memory("[1, 2]", "")
    with pytest.raises(capi.SvjgError):
        alnfilter.Tables.from_memory('{"a@+@b@+": [["x:DEL-1-2", 0]', "")


def _oracle_hits(lines, edges, alt, ids):
    idx = {s: i for i, s in enumerate(ids)}
    sv2, off, ln = [], [], []
    pos = 0
    for line in lines:
        nbytes = len(line.encode())
        for sv, allele in O.record_hits(line, edges, alt):
            sv2.append(idx[sv] * 2 + (allele if allele >= 0 else allele + 2))
            off.append(pos)
            ln.append(nbytes)
        pos += nbytes
    return (np.array(sv2, np.uint32), np.array(off, np.uint64), np.array(ln, np.uint32))


def test_json_emitter_byte_equal_c1(tmp_path):
    edges_text = read_golden("c1_svs_edges.json")
    gfa_text = read_golden("c1.gfa.gz")
    gaf = read_golden("c1.gaf.gz")
    t = alnfilter.Tables.from_memory(edges_text, gfa_text)
    sv2, off, ln = _oracle_hits(gaf.splitlines(True), json.loads(edges_text), alt_len_from_gfa_text(gfa_text), t.sv_ids)
    # shuffled: the emitter must restore file order inside each list
    perm = np.random.default_rng(1).permutation(len(sv2))
    res = alnfilter.FilterResult(None, {"n_hits": len(sv2)}, sv2[perm], off[perm], ln[perm])
    out = tmp_path / "c1_informative_aln.json"
    alnfilter.write_informative_json(t, gaf.encode(), res, str(out))
    assert out.read_text() == read_golden("c1_informative_aln.json.gz")


@pytest.mark.parametrize("threads,batch", [(1, 1 << 30), (5, 1 << 30), (3, 40_000), (16, 1)])
def test_json_emitter_threads_and_batches(tmp_path, monkeypatch, threads, batch):
    """The keys are rendered by several threads, a batch at a time: any split must give the same bytes."""
    edges_text = read_golden("c1_svs_edges.json")
    gfa_text = read_golden("c1.gfa.gz")
    gaf = read_golden("c1.gaf.gz")
    t = alnfilter.Tables.from_memory(edges_text, gfa_text)
    sv2, off, ln = _oracle_hits(gaf.splitlines(True), json.loads(edges_text), alt_len_from_gfa_text(gfa_text), t.sv_ids)
    res = alnfilter.FilterResult(None, {"n_hits": len(sv2)}, sv2, off, ln)
    monkeypatch.setenv("SVJG_JSON_THREADS", str(threads))
    monkeypatch.setenv("SVJG_JSON_BATCH", str(batch))
    out = tmp_path / "x.json"
    alnfilter.write_informative_json(t, gaf.encode(), res, str(out))
    assert out.read_text() == read_golden("c1_informative_aln.json.gz")
    empty = alnfilter.FilterResult(None, {"n_hits": 0}, sv2[:0], off[:0], ln[:0])
    alnfilter.write_informative_json(t, gaf.encode(), empty, str(out))
    assert out.read_text() == "{}"


def test_json_emitter_quirks_and_empty(tmp_path, quirks):
    edges = json.loads(quirks["edges"])
    alt = alt_len_from_gfa_text(quirks["gfa"])
    t = alnfilter.Tables.from_memory(quirks["edges"], quirks["gfa"])
    n = 0
    for case in quirks["cases"]:
        if case["rc"] != 0:
            continue
        sv2, off, ln = _oracle_hits(case["gaf"].splitlines(True), edges, alt, t.sv_ids)
        res = alnfilter.FilterResult(None, {"n_hits": len(sv2)}, sv2, off, ln)
        out = tmp_path / "q.json"
        alnfilter.write_informative_json(t, case["gaf"].encode(), res, str(out))
        assert out.read_text() == case["json"], case["name"]
        n += 1
    assert n >= 15


def test_translate_newlines_is_text_mode():
    """Carriage returns end lines for the reference (text-mode open()): the host mirror translates them
    exactly as Python does, and leaves buffers without one untouched (same object)."""
    import io
    plain = b"a\tb\nc\n"
    assert alnfilter.translate_newlines(plain) is plain
    assert alnfilter.translate_newlines(b"") == b""
    for text in ("a\r\nb\r\n", "a\rb\nc", "\r\r\n\n\r", "x\ty\r"):
        want = "".join(io.StringIO(text, newline=None))
        assert bytes(alnfilter.translate_newlines(text.encode())) == want.encode()


def test_gfa_loader_on_damaged_gfas_matches_the_reference():
    """tests/golden/fuzz_gfa.json.gz: 1200 small GFAs with damaged lines (columns dropped, odd white
    space, CR / CR LF line ends, repeated nodes), each loaded by the unmodified reference filter
    (tests/golden/make_fuzz.py gfa; its alt_node_len dictionary read from the frame of main()).  The C
    loader must refuse exactly the files the reference raises on and hold the same length under every
    name, in the tables the kernels probe (svjg_tables_alt_node_len)."""
    cases = json.loads(read_golden("fuzz_gfa.json.gz"))
    assert len(cases) == 1200 and 50 < sum(c["rc"] for c in cases) < 600
    n_names = 0
    for c in cases:
        try:
            t = alnfilter.Tables.from_memory("{}", c["gfa"])
        except capi.SvjgError as exc:
            assert c["rc"] == 1, (c["gfa"], str(exc))
            continue
        assert c["rc"] == 0, c["gfa"]
        assert t.num_alt_nodes == len(c["alt"]), c["gfa"]
        for name, n in c["alt"].items():
            assert t.alt_node_len(name) == n, (c["gfa"], name)
            n_names += 1
        assert t.alt_node_len("chr1:1-2") is None and t.alt_node_len("") is None
        t.close()
    assert n_names > 1500


def test_table_images_are_reproducible(quirks):
    """svjg_tables_image_hash: the host image (hash tables, name blob, entries, sv ids) of the golden
    catalogues, as built when the GPU suite last ran in full (round 2: chrom length in the plain-node key, 32-bit link hash).  A builder change that is meant to keep
    the uploaded bytes (a faster loader) must keep these; a layout change updates them on purpose."""
    want = {"c1": 8916239499250450647, "s2": 13201132800516666657, "s3": 14863090804196457325, "s4": 1219435239302125404}
    for tag, h in want.items():
        t = alnfilter.Tables.from_memory(read_golden("c1_svs_edges.json" if tag == "c1" else f"{tag}_svs_edges.json.gz"),
                                         read_golden(f"{tag}.gfa.gz"))
        assert t.image_hash == h, tag
    assert alnfilter.Tables.from_memory(quirks["edges"], quirks["gfa"]).image_hash == 15319934271598952324
    # the order of the keys in the file does not matter for what is found, but the ascending-order shortcut
    # of the reader must not change the outcome of repeated keys: the last one wins, at the first one's place
    a = alnfilter.Tables.from_memory('{"a:1-2@+@a:3-4@+": [["a:X", 0]], "b:1-2@+@b:3-4@+": [["b:Y", 1]], "a:1-2@+@a:3-4@+": [["a:Z", 1]]}', "")
    b = alnfilter.Tables.from_memory('{"a:1-2@+@a:3-4@+": [["a:Z", 1]], "b:1-2@+@b:3-4@+": [["b:Y", 1]]}', "")
    assert a.image_hash == b.image_hash and a.sv_ids == ["a:Z", "b:Y"]


def test_large_link_table_parsed_in_pieces_equals_the_sequential_parse():
    """A svs_edges.json beyond 32 MiB in the layout construct-graph.py:554 writes (json.dumps(indent=4)) is cut at
    top-level members and parsed side by side (csrc/tables.cpp: parse_edges_pieces); the same dictionary in any
    other layout goes through the sequential parser.  Same tables either way; a misleading line in the layout, a
    duplicate key across the cut and a damaged file fall back to the sequential parser and its verdict."""
    import os
    from svjg import alnfilter, capi
    if (os.cpu_count() or 1) < 2 or len(os.sched_getaffinity(0)) < 2:
        pytest.skip("one CPU: every file is parsed sequentially")
    n = 230_000
    d = {}
    for i in range(n):
        a, b = f"chr{i % 22 + 1}:{i * 100 + 1}-{i * 100 + 100}", f"chr{i % 22 + 1}:{i * 100 + 101}-{i * 100 + 200}"
        d[f"{a}@+@{b}@+"] = [[f"chr{i % 22 + 1}:DEL-{i * 100 + 100}-{i * 100 + 190}", i & 1]]
        if i % 7 == 0:
            d[f"{a}@+@{b}@+"].append([f"chr{i % 22 + 1}:INS-{i * 100 + 100}-1", 1])
    pretty = json.dumps(d, sort_keys=True, indent=4)
    compact = json.dumps(d, sort_keys=True)
    assert len(pretty) > (33 << 20) and "\n    \"" in pretty and "\n    \"" not in compact
    want = alnfilter.Tables.from_memory(compact, "")
    got = alnfilter.Tables.from_memory(pretty, "")
    assert got.image_hash == want.image_hash and got.num_links == want.num_links == n and got.sv_ids == want.sv_ids
    # a nested line dressed up as a top-level member start: valid JSON, same dictionary
    k = pretty.index("\n            \"chr", len(pretty) // 2)
    trap = pretty[:k] + "\n    \"" + pretty[k + len("\n            \""):]
    assert json.loads(trap) == d
    assert alnfilter.Tables.from_memory(trap, "").image_hash == want.image_hash
    # the first key once more at the very end (the last one wins, as in a dict): across the pieces
    first_key = min(d)
    dup = pretty[:pretty.rindex("\n}")] + ",\n    " + json.dumps(first_key) + ": [\n        [\n            \"chr1:DEL-7-77\",\n            1\n        ]\n    ]\n}"
    d2 = json.loads(dup)
    assert d2[first_key] == [["chr1:DEL-7-77", 1]]
    assert alnfilter.Tables.from_memory(dup, "").image_hash == alnfilter.Tables.from_memory(json.dumps(d2, sort_keys=True), "").image_hash
    # damaged in the second half: the sequential parser's error
    cut = pretty[: (len(pretty) * 3) // 4]
    with pytest.raises(capi.SvjgError):
        alnfilter.Tables.from_memory(cut, "")
    with pytest.raises(capi.SvjgError):
        alnfilter.Tables.from_memory(pretty + " x", "")
