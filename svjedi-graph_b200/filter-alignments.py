#!/usr/bin/env python3
"""Drop-in replacement of the reference's filter-alignments.py: same command line, same files, same
exit statuses; the hot path runs on the GPU through libsvjg.so (svjg/cli.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from svjg import cli  # noqa: E402

if __name__ == "__main__":
    cli.leave(cli.filter_main())
