import sys
sys.path.insert(0, "."); sys.path.insert(0, "svjedi-graph_b200"); sys.path.insert(0, "tests")
import numpy as np
from conftest import read_golden
from svjg import alnfilter, genotype
for tag in ("c1", "s3", "s4"):
    t = alnfilter.Tables.from_memory(read_golden("c1_svs_edges.json" if tag == "c1" else f"{tag}_svs_edges.json.gz"), read_golden(f"{tag}.gfa.gz")).to_device(0)
    gaf = read_golden(f"{tag}.gaf.gz").encode()
    res = alnfilter.filter_host(t, gaf)
    r2, text = alnfilter.filter_json_host(t, gaf)
    assert text is not None and (r2.counts == res.counts).all()
    vcf = read_golden("c1.vcf" if tag == "c1" else f"{tag}.vcf.gz")
    out, n = genotype.genotype_vcf(t, res.counts, vcf.encode())
    print(tag, res.n_hits, len(text), n)
    # the text once more slice by slice (key ranges rendered into slices of about 20 kB)
    import tempfile, os
    assert alnfilter.filter_json_begin(t, gaf) is not None
    with tempfile.TemporaryDirectory() as tmp:
        nb = alnfilter.filter_json_write(t, os.path.join(tmp, "x.json"), 20_000)
        assert open(os.path.join(tmp, "x.json"), "rb").read() == bytes(text) and nb == len(text)
