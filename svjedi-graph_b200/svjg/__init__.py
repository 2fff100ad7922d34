"""svjg — B200-native hot path of SVJedi-graph (filter + count + genotype).

Host-side mirror of the reference's post-mapping stage
(filter-alignments.py / predict-genotype.py); all compute goes through the
C-ABI library ``libsvjg.so`` (``include/svjg.h``), loaded by :mod:`svjg.capi`.
"""
__version__ = "0.1.0"
