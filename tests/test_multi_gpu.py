"""One GAF file on several GPUs of a node (the stage boundary svjedi-graph.py:113-128 with the records
sharded): byte ranges cut at line ends, one device each, counters summed, hits merged in range order ->
ONE informative_aln.json and ONE genotype.vcf, byte-equal to the reference's own outputs (tests/golden) and to
the oracle.  Skipped on a box with one GPU (`gpurun --gpus 2`)."""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG, alt_len_from_gfa_text, read_golden
from oracle import svjg_oracle as O

pytestmark = pytest.mark.gpu


def _devices():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.fixture(scope="module")
def n_dev():
    n = _devices()
    if n < 2:
        pytest.skip("needs two GPUs")
    return min(n, 4)


@pytest.mark.parametrize("tag", ["s2", "s3", "s4"])
def test_one_file_on_n_devices_equals_the_reference(n_dev, tag, tmp_path):
    from svjg import alnfilter, genotype
    edges, gfa = read_golden(f"{tag}_svs_edges.json.gz"), read_golden(f"{tag}.gfa.gz")
    gaf_text = read_golden(f"{tag}.gaf.gz")
    gaf = gaf_text.encode()
    t0 = alnfilter.Tables.from_memory(edges, gfa).to_device(0)
    tables = [t0] + [t0.clone().to_device(d) for d in range(1, n_dev)]
    res = alnfilter.filter_host_multi(tables, gaf)
    # the oracle: the reference's loop on the CPU
    want = O.filter_alignments(gaf_text.splitlines(True), json.loads(edges), alt_len_from_gfa_text(gfa))
    got = {t0.sv_ids[i]: (int(res.counts[i, 0]), int(res.counts[i, 1])) for i in range(t0.num_sv) if res.counts[i].any()}
    assert got == O.hit_counts(want) and res.n_hits == sum(a + b for a, b in got.values())
    one = alnfilter.filter_host(t0, gaf)
    assert (one.counts == res.counts).all() and {k: v for k, v in one.stats.items() if k != "n_exact"} == \
        {k: v for k, v in res.stats.items() if k != "n_exact"}
    out = tmp_path / "x_informative_aln.json"
    alnfilter.write_informative_json(t0, gaf, res, str(out))
    assert hashlib.sha256(out.read_bytes()).hexdigest() == read_golden(f"{tag}_informative_aln.sha256").strip()
    text, n = genotype.genotype_vcf(t0, res.counts, read_golden(f"{tag}.vcf.gz").encode())
    assert text == read_golden(f"{tag}_genotype.vcf.gz")


def test_command_line_on_two_gpus(n_dev, tmp_path):
    """SVJG_GPUS=2 filter-alignments.py: the same file as with one GPU, the reference's file."""
    tag = "s3"
    import gzip
    for name, src in ((f"{tag}.gaf", f"{tag}.gaf.gz"), (f"{tag}.gfa", f"{tag}.gfa.gz"), (f"{tag}_svs_edges.json", f"{tag}_svs_edges.json.gz")):
        (tmp_path / name).write_text(read_golden(src))
    env = dict(os.environ, SVJG_GPUS="2")
    r = subprocess.run([sys.executable, os.path.join(PKG, "filter-alignments.py"), "-a", f"{tag}.gaf", "-g", f"{tag}.gfa", "-p", tag],
                       cwd=tmp_path, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = (tmp_path / f"{tag}_informative_aln.json").read_bytes()
    assert hashlib.sha256(got).hexdigest() == read_golden(f"{tag}_informative_aln.sha256").strip()
    # a damaged line in the second half: exit status 1 and the byte offset of the line in the FILE
    lines = read_golden(f"{tag}.gaf.gz").splitlines(True)
    k = (3 * len(lines)) // 4
    lines[k] = "short\tline\n"
    (tmp_path / "bad.gaf").write_text("".join(lines))
    os.symlink(f"{tag}_svs_edges.json", tmp_path / "bad_svs_edges.json")
    r = subprocess.run([sys.executable, os.path.join(PKG, "filter-alignments.py"), "-a", "bad.gaf", "-g", f"{tag}.gfa", "-p", "bad"],
                       cwd=tmp_path, env=env, capture_output=True, text=True)
    assert r.returncode == 1 and f"byte {sum(len(x) for x in lines[:k])}:" in r.stderr, r.stderr


@pytest.mark.parametrize("tag", ["s2", "s3", "s4"])
def test_text_rendered_on_device_0_from_all_ranges(n_dev, tag, tmp_path):
    """filter_json_multi_begin + filter_json_finish / filter_json_write: every range stays on its device, device 0
    renders the whole text and reads the other ranges' lines over peer access -- the reference's file."""
    from svjg import alnfilter
    edges, gfa = read_golden(f"{tag}_svs_edges.json.gz"), read_golden(f"{tag}.gfa.gz")
    gaf = read_golden(f"{tag}.gaf.gz").encode()
    t0 = alnfilter.Tables.from_memory(edges, gfa).to_device(0)
    tables = [t0] + [t0.clone().to_device(d) for d in range(1, n_dev)]
    want_sha = read_golden(f"{tag}_informative_aln.sha256").strip()
    one = alnfilter.filter_host(t0, gaf)
    res = alnfilter.filter_json_multi_begin(tables, gaf)
    assert res is not None, "no peer access on this box?"
    assert (res.counts == one.counts).all()
    assert {k: v for k, v in res.stats.items() if k != "n_exact"} == {k: v for k, v in one.stats.items() if k != "n_exact"}
    text = alnfilter.filter_json_finish(t0)
    assert text is not None and hashlib.sha256(text).hexdigest() == want_sha
    # again, to a file in small slices; the workspace is reused (hit arrays grown by the first gather)
    assert alnfilter.filter_json_multi_begin(tables, gaf) is not None
    out = tmp_path / "sliced.json"
    n = alnfilter.filter_json_write(t0, str(out), 50_000)
    assert n == out.stat().st_size and hashlib.sha256(out.read_bytes()).hexdigest() == want_sha
    # a single device afterwards: the same handle renders its own file again
    res1, text1 = alnfilter.filter_json_host(t0, gaf)
    assert hashlib.sha256(text1).hexdigest() == want_sha


def test_damaged_line_on_the_multi_device_json_route(n_dev):
    from svjg import alnfilter
    tag = "s3"
    edges, gfa = read_golden(f"{tag}_svs_edges.json.gz"), read_golden(f"{tag}.gfa.gz")
    lines = read_golden(f"{tag}.gaf.gz").splitlines(True)
    k = (3 * len(lines)) // 4
    lines[k] = "short\tline\n"
    t0 = alnfilter.Tables.from_memory(edges, gfa).to_device(0)
    tables = [t0, t0.clone().to_device(1)]
    with pytest.raises(alnfilter.InputError) as exc:
        alnfilter.filter_json_multi_begin(tables, "".join(lines).encode())
    assert f"byte {sum(len(x) for x in lines[:k])}:" in str(exc.value)
