"""informative_aln.json rendered on the device (svjg_filter_json_host, csrc/json.cu) against the reference's
own files (tests/golden, written by the unmodified filter-alignments.py) and against the host emitter
(svjg_emit_informative_json) on generated inputs: byte for byte."""
import hashlib
import io
import json
import os

import numpy as np
import pytest

from conftest import read_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from svjg import alnfilter, capi
    try:
        capi.check(capi.lib.svjg_device_init(0))
    except Exception as exc:  # pragma: no cover
        pytest.skip(f"no CUDA device: {exc}")
    return alnfilter, capi


def _tables(alnfilter, tag):
    edges = read_golden("c1_svs_edges.json" if tag == "c1" else f"{tag}_svs_edges.json.gz")
    return alnfilter.Tables.from_memory(edges, read_golden(f"{tag}.gfa.gz")).to_device(0)


@pytest.mark.parametrize("tag", ["c1", "s2", "s3", "s4"])
def test_device_json_equals_the_reference_file(gpu, tag):
    alnfilter, capi = gpu
    t = _tables(alnfilter, tag)
    gaf = read_golden(f"{tag}.gaf.gz").encode()
    res, text = alnfilter.filter_json_host(t, gaf)
    assert text is not None
    if tag == "c1":
        assert bytes(text) == read_golden("c1_informative_aln.json.gz").encode()
    else:
        assert hashlib.sha256(text).hexdigest() == read_golden(f"{tag}_informative_aln.sha256").strip()
    host = alnfilter.filter_host(t, gaf)
    assert (res.counts == host.counts).all() and res.stats == host.stats


def test_empty_inputs_and_no_hits(gpu):
    alnfilter, capi = gpu
    t = _tables(alnfilter, "s2")
    res, text = alnfilter.filter_json_host(t, b"")
    assert bytes(text) == b"{}" and res.n_hits == 0
    one = b"r\t100\t0\t100\t+\t>chrZ:1-50\t50\t0\t50\t50\t50\t60\n"
    res, text = alnfilter.filter_json_host(t, one)
    assert bytes(text) == b"{}" and res.stats["n_records"] == 1


def test_escapes_cg_cut_and_chunk_seams(gpu, tmp_path):
    """Lines with a "cg:Z:" tag (the text stops in front of it, :166), with quotes, backslashes and control
    bytes in the read name and the tags, several 64 MiB upload chunks with line ends at odd byte offsets, a key
    with both lists and keys with one empty list -- against the host emitter."""
    alnfilter, capi = gpu
    names = [f"chrQ:{i * 100 + 1}-{(i + 1) * 100}" for i in range(40)]
    edges = {}
    for i in range(39):
        edges[f"{names[i]}@+@{names[i + 1]}@+"] = [[f"chrQ:DEL-{(i + 1) * 100}-{(i + 1) * 100 + 60}", i % 2], [f"chrQ:INS-{i}-1", 1 - i % 2]] \
            if i % 5 == 0 else [[f"chrQ:DEL-{(i + 1) * 100}-{(i + 1) * 100 + 60}", i % 2]]
    t = alnfilter.Tables.from_memory(json.dumps(edges), "").to_device(0)
    rng = np.random.default_rng(5)
    odd = ['plain', 'with"quote', 'back\\slash', 'ctl\x01\x1f', 'form\x0cfeed', 'bs\x08', 'del\x7f', "cr\rinside"]
    lines = []
    for k in range(260_000):
        i = int(rng.integers(0, 38))
        n = int(rng.integers(2, 4))
        path = "".join(">" + names[i + j] for j in range(n) if i + j < 40)
        tl = 100 * path.count(">")
        name = odd[k % len(odd)] + str(k)
        tags = "tp:A:P\tcm:i:5" + ("\tcg:Z:50M2D48M\ttail:Z:x" if k % 3 == 0 else "") + ("\tzz:Z:" + "y" * int(rng.integers(0, 400)))
        lines.append(f"{name}\t{tl}\t0\t{tl}\t+\t{path}\t{tl}\t0\t{tl}\t{tl - 3}\t{tl}\t60\t{tags}\n")
    gaf = "".join(lines).encode()
    assert len(gaf) > (64 << 20) + 1000                    # more than one upload chunk
    res, text = alnfilter.filter_json_host(t, gaf)
    assert text is not None and res.n_hits > 100_000
    host = alnfilter.filter_host(t, gaf)
    out = os.path.join(tmp_path, "h.json")
    alnfilter.write_informative_json(t, gaf, host, out)
    with open(out, "rb") as fh:
        want = fh.read()
    assert hashlib.sha256(bytes(text)).hexdigest() == hashlib.sha256(want).hexdigest()
    assert (res.counts == host.counts).all() and res.stats == host.stats
    d = json.loads(bytes(text))                              # and it is JSON a reader accepts
    assert sum(len(a) + len(b) for a, b in d.values()) == res.n_hits


def test_declines_non_ascii_lines(gpu):
    alnfilter, capi = gpu
    t = _tables(alnfilter, "s2")
    gaf = read_golden("s2.gaf.gz").encode()
    res, text = alnfilter.filter_json_host(t, gaf)
    first_hit_line = bytes(text).split(b'"')[3].split(b"\\t")[0]          # a read name that is stored
    bad = gaf.replace(first_hit_line, first_hit_line[:-1] + "é".encode(), 1)
    res2, text2 = alnfilter.filter_json_host(t, bad)
    assert text2 is None and (res2.counts == res.counts).all()


def test_full_c2_batch_equals_host_emitter(gpu, tmp_path):
    from svjg import synth
    alnfilter, capi = gpu
    g, vcf, gaf_text = synth.make_workload("C2", scale=float(os.environ.get("SVJG_TEST_FULL_SCALE", "1.0")))
    buf = io.StringIO()
    g.write_gfa(buf)
    t = alnfilter.Tables.from_memory(g.edges_json(), buf.getvalue()).to_device(0)
    gaf = gaf_text.encode()
    res, text = alnfilter.filter_json_host(t, gaf)
    host = alnfilter.filter_host(t, gaf)
    out = os.path.join(tmp_path, "h.json")
    alnfilter.write_informative_json(t, gaf, host, out)
    h = hashlib.sha256()
    with open(out, "rb") as fh:
        for block in iter(lambda: fh.read(1 << 24), b""):
            h.update(block)
    assert text is not None and hashlib.sha256(text).hexdigest() == h.hexdigest()


@pytest.mark.parametrize("tag,slice_bytes", [("c1", 0), ("c1", 1), ("c1", 4096), ("s2", 1 << 16), ("s3", 1 << 20), ("s4", 3000)])
def test_sliced_file_equals_the_text_in_one_piece(gpu, tmp_path, tag, slice_bytes):
    """svjg_filter_json_write: the text rendered key range by key range into slices (of one key each at
    slice_bytes = 1) and written by the library -- the same bytes as the one-piece rendering and as the
    reference's file."""
    alnfilter, capi = gpu
    t = _tables(alnfilter, tag)
    gaf = read_golden(f"{tag}.gaf.gz").encode()
    _, text = alnfilter.filter_json_host(t, gaf)
    whole = bytes(text)
    res = alnfilter.filter_json_begin(t, gaf)
    out = tmp_path / "sliced.json"
    n = alnfilter.filter_json_write(t, str(out), slice_bytes)
    got = out.read_bytes()
    assert n == len(got) == len(whole) and got == whole
    if tag == "c1":
        assert got == read_golden("c1_informative_aln.json.gz").encode()
    assert res.n_hits > 0


def test_sliced_file_corner_cases(gpu, tmp_path):
    alnfilter, capi = gpu
    t = _tables(alnfilter, "s2")
    out = tmp_path / "x.json"
    # nothing appended at all: "{}"
    alnfilter.filter_json_begin(t, b"")
    assert alnfilter.filter_json_write(t, str(out), 64) == 2 and out.read_bytes() == b"{}"
    # no second half without a first
    with pytest.raises(capi.SvjgError):
        alnfilter.filter_json_write(t, str(out))
    # a path that cannot be written: an error, no file
    gaf = read_golden("s2.gaf.gz").encode()
    alnfilter.filter_json_begin(t, gaf)
    with pytest.raises(capi.SvjgError):
        alnfilter.filter_json_write(t, str(tmp_path / "no_such_dir" / "x.json"))
    # a stored line with a non-ASCII byte: the renderer declines and writes nothing
    lines = gaf.split(b"\n")
    res0 = alnfilter.filter_host(t, gaf)
    first = int(np.sort(res0.hit_off)[0])
    k = gaf[:first].count(b"\n")
    lines[k] = lines[k] + b"\txx:Z:\xc3\xa9"
    odd = b"\n".join(lines)
    out2 = tmp_path / "y.json"
    assert alnfilter.filter_json_begin(t, odd) is not None
    assert alnfilter.filter_json_write(t, str(out2)) is None and not out2.exists()
