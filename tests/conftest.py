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
