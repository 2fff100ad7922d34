import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = os.path.join(ROOT, "svjedi-graph_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def read_golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".gz"):
        with gzip.open(path, "rb") as fh:
            return fh.read().decode()
    with open(path, "r") as fh:
        return fh.read()


@pytest.fixture(scope="session")
def golden():
    return read_golden


@pytest.fixture(scope="session")
def quirks():
    return json.loads(read_golden("quirks.json"))


def alt_len_from_gfa_text(text):
    """alt-node table straight from GFA text (test helper; mirrors what the
    product's loader must produce)."""
    out = {}
    for line in text.splitlines(True):
        if line.startswith("S"):
            cols = line.split("\t")
            if "." in cols[1].split(":")[-1]:
                out[cols[1]] = len(line.rstrip().split("\t")[2])
    return out


def revisit_walks(n_lines=1500, seed=77):
    """(edges, GAF lines) of random walks WITH LOOPS on a chain graph with jump links, inversion links and
    breakends whose two ends lie in one node: a name that comes twice takes its strand from its first occurrence
    (filter-alignments.py:206) and its place in the overlap sums from its first index (:269-271)."""
    import numpy as np
    rng = np.random.Generator(np.random.PCG64(seed))
    n_nodes, step = 60, 131
    names = [f"chr7:{i * step + 1}-{(i + 1) * step}" for i in range(n_nodes)]
    edges = {}
    for i in range(n_nodes - 1):
        edges[f"{names[i]}@+@{names[i + 1]}@+"] = [[f"chr7:DEL-{(i + 1) * step}-{(i + 1) * step + 60}", 0]]
    for i in range(0, n_nodes - 2, 3):
        edges[f"{names[i]}@+@{names[i + 2]}@+"] = [[f"chr7:DEL-{(i + 1) * step}-{(i + 2) * step}", 1]]
    for i in range(0, n_nodes - 1, 5):
        edges[f"{names[i]}@+@{names[i]}@+"] = [[f"chr7:BND-{i}", 1]]
        edges[f"{names[i]}@+@{names[i + 1]}@-"] = [[f"chr7:INV-{i}", 1]]
    lines = []
    for k in range(n_lines):
        n = int(rng.integers(2, 12))
        walk = [int(rng.integers(0, n_nodes - 12))]
        for _ in range(n - 1):
            r = rng.random()
            walk.append(walk[-1] if r < 0.15 else max(0, walk[-1] - int(rng.integers(1, 4))) if r < 0.3
                        else walk[-1] + int(rng.integers(1, 3)))
        orient = [">" if rng.random() < 0.8 else "<" for _ in walk]
        path = "".join(o + names[i] for o, i in zip(orient, walk))
        tlen = len(walk) * step
        ts = int(rng.integers(0, step))
        te = tlen - int(rng.integers(0, step))
        lines.append(f"r{k}\t{tlen}\t0\t{tlen}\t+\t{path}\t{tlen}\t{ts}\t{te}\t{tlen - 5}\t{tlen}\t60\ttp:A:P\n")
    return edges, lines
