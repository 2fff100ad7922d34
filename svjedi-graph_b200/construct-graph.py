#!/usr/bin/env python3
"""Drop-in replacement of the reference's construct-graph.py: same command line, same three files
(<out>.gfa, <prefix>_svs_edges.json, <prefix>_ignored_svs.txt), same bytes; node lookups are
dictionary probes over the sorted breakpoints instead of one scan over every node per SV
(svjg/graphgen.py).  Host-only: the graph is built once per catalogue, there is nothing here
for the GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from svjg import graphgen  # noqa: E402

if __name__ == "__main__":
    sys.exit(graphgen.construct_main())
