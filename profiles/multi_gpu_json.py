#!/usr/bin/env python3
"""One GAF file on the N GPUs of the box, host bytes in -> informative_aln.json file out: the round-1 route (hit lists to
the host, host emitter) against the device route (ranges stay on their devices, device 0 renders the text and reads
the other ranges' lines over NVLink, slices written as they come back).
usage: multi_gpu_json.py [workload=C2] [scale=1.0]"""
import hashlib
import io
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "svjedi-graph_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from svjg import alnfilter, synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
n_dev = torch.cuda.device_count()
g, vcf, gaf_text = synth.make_workload(name, scale=scale)
buf = io.StringIO()
g.write_gfa(buf)
gaf = alnfilter.PinnedBytes(len(gaf_text))
gaf.array[:] = np.frombuffer(gaf_text.encode(), dtype=np.uint8)
del gaf_text
t0 = alnfilter.Tables.from_memory(g.edges_json(), buf.getvalue()).to_device(0)
tables = [t0] + [t0.clone().to_device(d) for d in range(1, n_dev)]
print(f"{name} x{scale}: {gaf.array.size / 1e6:.1f} MB GAF, {t0.num_sv} sv keys, {n_dev} GPUs")


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as fh:
        for piece in iter(lambda: fh.read(1 << 24), b""):
            h.update(piece)
    return h.hexdigest()


with tempfile.TemporaryDirectory() as tmp:
    out = os.path.join(tmp, "x.json")
    for label, n in (("1 GPU ", 1), (f"{n_dev} GPUs", n_dev)):
        ts = tables[:n]
        for rep in range(3):
            a = time.perf_counter()
            res = alnfilter.filter_host_multi(ts, gaf.array) if n > 1 else alnfilter.filter_host(t0, gaf.array)
            b = time.perf_counter()
            alnfilter.write_informative_json(t0, gaf.array, res, out)
            c = time.perf_counter()
        want = sha(out)
        print(f"{label} hit lists to the host + host emitter: filter {1e3 * (b - a):7.1f} ms, JSON file {1e3 * (c - b):7.1f} ms")
        for rep in range(3):
            a = time.perf_counter()
            res = alnfilter.filter_json_multi_begin(ts, gaf.array) if n > 1 else alnfilter.filter_json_begin(t0, gaf.array)
            b = time.perf_counter()
            nbytes = alnfilter.filter_json_write(t0, out)
            c = time.perf_counter()
        print(f"{label} text rendered on device 0, written in slices: filter {1e3 * (b - a):7.1f} ms, JSON file {1e3 * (c - b):7.1f} ms "
              f"({nbytes / 1e6:.1f} MB, same bytes: {sha(out) == want})")
