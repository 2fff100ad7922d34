"""BASELINE.json's configurations at FULL size bit for bit: every counter and every hit tuple of the CUDA
path against the C restatement of the reference filter (oracle/svjg_oracle.c, pinned on the reference's
outputs by tests/test_c_oracle.py), the informative_aln.json written from them, and the genotype VCF.
C2 (25 k SVs, 3 M records), C4 (20 k BND, 3 M records), C3 (100 k clustered SVs, 6 M records, paths of
hundreds of nodes) whole; C5 with its whole 1 M-SV catalogue (a 459 MB table image, beyond L2) and a
500 k-record shard of its 200 M records.  Runs last (file name) and loads no tensor library.
SVJG_TEST_FULL_SCALE shrinks the record counts for a quick run."""
import io
import json
import os

import numpy as np
import pytest

from conftest import alt_len_from_gfa_text
from oracle import svjg_oracle as O


# (name, record scale, catalogue scale): C5's catalogue is whole, its records a shard (the full 200 M are 66 GB)
CASES = [("C2", 1.0, None), ("C4", 1.0, None), ("C3", 1.0, None), ("C5", 0.0025, 1.0)]


@pytest.mark.gpu
@pytest.mark.parametrize("name,rec_scale,cat_scale", CASES)
def test_full_size_exact_against_the_c_oracle(name, rec_scale, cat_scale):
    from svjg import alnfilter, genotype, synth
    from oracle import c_oracle as CO
    CO.ensure_built()
    scale = rec_scale * float(os.environ.get("SVJG_TEST_FULL_SCALE", "1.0"))
    g, vcf, gaf_text = synth.make_workload(name, scale=scale, catalogue_scale=cat_scale)
    buf = io.StringIO()
    g.write_gfa(buf)
    edges_text, gfa_text = g.edges_json(), buf.getvalue()
    gaf = gaf_text.encode()
    del gaf_text
    # checker
    ct = CO.Tables(json.loads(edges_text), alt_len_from_gfa_text(gfa_text))
    want_counts, want_stats, w_sv2, w_off, w_len = CO.filter_counts(ct, gaf, want_hits=True)
    # CUDA path through the C ABI, host buffers
    t = alnfilter.Tables.from_memory(edges_text, gfa_text).to_device(0)
    assert t.sv_ids == ct.sv_ids
    res = alnfilter.filter_host(t, gaf)
    assert res.stats["n_records"] == want_stats["n_records"] == gaf.count(b"\n")
    assert res.stats["n_multi"] == want_stats["n_multi"]
    assert res.n_hits == want_stats["n_hits"] > 1000 * scale / rec_scale
    assert (res.counts == want_counts).all()
    # the hit tuples: the kernels append in any order, the reference in file order
    got = np.stack([res.hit_off.astype(np.uint64), res.hit_sv2.astype(np.uint64), res.hit_len.astype(np.uint64)])
    want = np.stack([w_off.astype(np.uint64), w_sv2.astype(np.uint64), w_len.astype(np.uint64)])
    got = got[:, np.lexsort(got[::-1])]
    want = want[:, np.lexsort(want[::-1])]
    assert (got == want).all()
    # informative_aln.json of the whole batch: the library's emitter against json.dumps of the reference's dictionary
    # rebuilt from the checker's hit tuples (file order)
    import hashlib
    import tempfile
    d = {}
    view = memoryview(gaf)
    for s2, o, n in zip(w_sv2.tolist(), w_off.tolist(), w_len.tolist()):
        d.setdefault(ct.sv_ids[s2 >> 1], [[], []])[s2 & 1].append(O.kept_text(str(view[o:o + n], "ascii")))
    want_sha = hashlib.sha256(O.dumps_informative(d).encode()).hexdigest()
    del d
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "informative_aln.json")
        alnfilter.write_informative_json(t, gaf, res, out)
        h = hashlib.sha256()
        with open(out, "rb") as fh:
            for block in iter(lambda: fh.read(1 << 24), b""):
                h.update(block)
    assert h.hexdigest() == want_sha
    # genotypes from those counters: the VCF text against the line-by-line oracle
    text, n = genotype.genotype_vcf(t, res.counts, vcf.encode())
    assert (text, n) == O.genotype_vcf(CO.counts_dict(ct, want_counts), vcf.splitlines(True))
    assert n > 100 * scale / rec_scale


@pytest.mark.gpu
def test_offsets_beyond_4_gib():  # noqa: C901
    """A per-GPU shard of C5 is 8 GB: byte offsets into it do not fit 32 bits.  A 25 MB block of records is
    laid out ~190 times (4.6 GiB); every counter must be that multiple of the block's, every hit tuple the
    block's moved by a multiple of the block length (checker: the C oracle on ONE block), and the JSON text
    rendered on the device from the resident 4.6 GiB must hold every list that many times over."""
    import hashlib
    from svjg import alnfilter, synth
    from oracle import c_oracle as CO
    CO.ensure_built()
    g, vcf, gaf_text = synth.make_workload("C2", scale=0.05)
    buf = io.StringIO()
    g.write_gfa(buf)
    edges_text, gfa_text = g.edges_json(), buf.getvalue()
    block = np.frombuffer(gaf_text.encode(), dtype=np.uint8)
    L = int(block.size)
    K = int(float(os.environ.get("SVJG_TEST_BIG_GIB", "4.3")) * (1 << 30)) // L + 1
    ct = CO.Tables(json.loads(edges_text), alt_len_from_gfa_text(gfa_text))
    b_counts, b_stats, b_sv2, b_off, b_len = CO.filter_counts(ct, block.tobytes(), want_hits=True)
    nb = int(b_stats["n_hits"])
    assert nb > 1000
    big = np.tile(block, K)
    assert big.size == K * L and (big.size > (1 << 32) or "SVJG_TEST_BIG_GIB" in os.environ)
    pinned = alnfilter.RegisteredBytes(big)
    t = alnfilter.Tables.from_memory(edges_text, gfa_text).to_device(0)
    # chunked route (64 MB pieces cut at line ends, hits with absolute offsets)
    res = alnfilter.filter_host(t, big)
    assert res.stats["n_records"] == K * b_stats["n_records"]
    assert res.n_hits == K * nb
    assert (res.counts.astype(np.uint64) == b_counts.astype(np.uint64) * K).all()
    order = np.lexsort((res.hit_sv2, res.hit_off))
    off = res.hit_off[order].astype(np.uint64)
    sv2 = res.hit_sv2[order]
    ln = res.hit_len[order]
    # the block's tuples in file order (several tuples may share a line: sort both sides alike)
    bo = np.lexsort((b_sv2, b_off))
    w_off = (b_off[bo].astype(np.uint64)[None, :] + (np.arange(K, dtype=np.uint64) * np.uint64(L))[:, None]).ravel()
    assert int(w_off.max()) > (1 << 32) or "SVJG_TEST_BIG_GIB" in os.environ
    assert (off == w_off).all()
    assert (sv2 == np.tile(b_sv2[bo], K)).all()
    assert (ln == np.tile(b_len[bo], K)).all()
    del res, order, off, sv2, ln, w_off
    # whole-file-resident route with the text rendered on the device
    res2, text = alnfilter.filter_json_host(t, big)
    assert text is not None
    assert (res2.counts.astype(np.uint64) == b_counts.astype(np.uint64) * K).all()
    d = {}
    view = memoryview(block.tobytes())
    for s2, o, n in zip(b_sv2.tolist(), b_off.tolist(), b_len.tolist()):
        d.setdefault(ct.sv_ids[s2 >> 1], [[], []])[s2 & 1].append(O.kept_text(str(view[o:o + n], "ascii")))

    def pieces(times):
        """json.dumps(d, sort_keys=True, indent=4) with every list `times` times over, key by key."""
        def lst(elems):
            if not elems:
                return "        []"
            return "        [\n" + ",\n".join(["            " + json.dumps(e) for e in elems] * times) + "\n        ]"
        yield "{\n" if d else "{"
        for i, key in enumerate(sorted(d)):
            yield ("" if i == 0 else ",\n") + "    " + json.dumps(key) + ": [\n" + lst(d[key][0]) + ",\n" + lst(d[key][1]) + "\n    ]"
        yield "\n}" if d else "}"
    assert "".join(pieces(1)) == O.dumps_informative(d)                  # the helper writes what json.dumps writes
    h = hashlib.sha256()
    for piece in pieces(K):
        h.update(piece.encode())
    got = hashlib.sha256()
    for i in range(0, len(text), 1 << 26):
        got.update(text[i:i + (1 << 26)])
    assert got.hexdigest() == h.hexdigest()
    # ... and the same text written slice by slice (svjg_filter_json_write), never whole in memory
    import tempfile
    del text
    assert alnfilter.filter_json_begin(t, big) is not None
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "sliced.json")
        n = alnfilter.filter_json_write(t, out)
        assert n == os.path.getsize(out)
        got = hashlib.sha256()
        with open(out, "rb") as fh:
            for piece in iter(lambda: fh.read(1 << 24), b""):
                got.update(piece)
    assert got.hexdigest() == h.hexdigest()
    del pinned
