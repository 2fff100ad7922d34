"""ctypes front of oracle/svjg_oracle.c — TEST INFRASTRUCTURE ONLY (see the header of that file).

The tables are the Python oracle's own (``load_link_table`` / ``load_alt_node_len``: plain dicts as
``json.load`` and the GFA scan give them); this module flattens them for the C code, shards the
GAF at line ends over threads (ctypes releases the GIL) and folds the results back."""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

OK, RAISES, UNSUPPORTED, HITS_FULL, NOMEM = range(5)
POISON, POISON_ALWAYS = 0xFFFFFFFF, 0xFFFFFFFE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libsvjg_oracle.so")
_lib = None


def ensure_built():
    """Builds oracle/_build/libsvjg_oracle.so if it is not there (gcc, a second); a library that travelled
    with the tree is used as it is."""
    if not os.path.exists(LIB_PATH):
        import subprocess
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        ensure_built()
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        _lib.svjg_oracle_filter.restype = C.c_int
        _lib.svjg_oracle_filter.argtypes = [vp, C.c_uint64, C.c_uint64, vp, vp, C.c_uint32, vp, vp, vp, vp, vp, C.c_uint32,
                                            C.c_int64, vp, vp, vp, vp, C.c_uint64, vp]
    return _lib


class OracleRaises(Exception):
    """The reference raises on the line at ``offset`` (exit status 1)."""

    def __init__(self, offset):
        super().__init__(f"the reference raises on the line at byte {offset}")
        self.offset = offset


class OracleUnsupported(Exception):
    """A spelling the C restatement does not cover (non-ASCII bytes, integers beyond 18 digits)."""

    def __init__(self, offset):
        super().__init__(f"line at byte {offset}: not covered by the C restatement")
        self.offset = offset


def _pairs(entry):
    """(sv_id, allele) as ``for sv_id, allele in d_link_sv[key]`` unpacks one entry, or None where that raises."""
    try:
        sv_id, allele = entry
    except (TypeError, ValueError):
        return None
    return sv_id, allele


class Tables:
    """Flat copy of (d_link_sv, alt_node_len) with the per-entry behaviour of filter-alignments.py
    :153-166 decided once: a valid entry is ``2 * rank(sv_id) + list index``; POISON raises once the
    overlap test has passed; POISON_ALWAYS raises as soon as the entry is reached."""

    def __init__(self, d_link_sv, alt_len):
        ok_ids = set()
        decoded = {}
        for key, ents in d_link_sv.items():
            out = []
            try:
                it = list(ents)
            except TypeError:                       # `for ... in 5`
                decoded[key] = [POISON_ALWAYS]
                continue
            for e in it:
                p = _pairs(e)
                if p is None:
                    out.append(POISON_ALWAYS)
                    continue
                sv_id, allele = p
                good = isinstance(sv_id, str) and ":" in sv_id and not isinstance(allele, float) \
                    and isinstance(allele, int) and allele in (0, 1, -1, -2)
                if good:
                    ok_ids.add(sv_id)
                    out.append((sv_id, int(allele) % 2))            # [[], []][-1] is the alt list, [-2] the ref list
                else:
                    out.append(POISON)
            decoded[key] = out
        self.sv_ids = sorted(ok_ids)
        rank = {s: i for i, s in enumerate(self.sv_ids)}
        keys = list(decoded)
        kb = [k.encode() for k in keys]
        self.key_off = np.zeros(len(kb) + 1, np.uint64)
        self.key_off[1:] = np.cumsum([len(b) for b in kb], dtype=np.uint64) if kb else []
        self.keys = b"".join(kb) or b"\0"
        ent, begin = [], [0]
        for k in keys:
            for e in decoded[k]:
                ent.append(e if isinstance(e, int) else 2 * rank[e[0]] + e[1])
            begin.append(len(ent))
        self.ent = np.array(ent or [0], np.uint32)
        self.ent_begin = np.array(begin, np.uint32)
        self.n_keys = len(keys)
        names = list(alt_len)
        nb = [n.encode() for n in names]
        self.alt_off = np.zeros(len(nb) + 1, np.uint64)
        self.alt_off[1:] = np.cumsum([len(b) for b in nb], dtype=np.uint64) if nb else []
        self.alts = b"".join(nb) or b"\0"
        self.alt_len = np.array([alt_len[n] for n in names] or [0], np.int64)
        self.n_alt = len(names)


def _cuts(buf, parts):
    n = len(buf)
    cuts = [0]
    for r in range(1, parts):
        p = max(cuts[-1], n * r // parts)
        j = buf.find(b"\n", p)
        cuts.append(n if j < 0 else j + 1)
    cuts.append(n)
    return cuts


def filter_counts(tables, gaf, d_over=100, want_hits=False, threads=None):
    """(counts uint32 [num_sv, 2], stats dict[, hit_sv2, hit_off, hit_len]) for the GAF bytes ``gaf``.
    Raises OracleRaises / OracleUnsupported for the first line (lowest offset) that does not pass."""
    gaf = bytes(gaf) if not isinstance(gaf, bytes) else gaf
    threads = threads or min(32, os.cpu_count() or 1)
    parts = max(1, min(threads, len(gaf) // (1 << 20) + 1))
    cuts = _cuts(gaf, parts)
    num = max(1, len(tables.sv_ids))
    base = C.cast(C.c_char_p(gaf), C.c_void_p).value

    def run(r):
        lo, hi = cuts[r], cuts[r + 1]
        counts = np.zeros(2 * num, np.uint32)
        stats = np.zeros(4, np.uint64)
        cap = max(1024, (hi - lo) // 16) if want_hits else 0
        while True:
            hs = np.empty(cap, np.uint32) if want_hits else None
            ho = np.empty(cap, np.uint64) if want_hits else None
            hl = np.empty(cap, np.uint32) if want_hits else None
            counts[:] = 0
            rc = lib().svjg_oracle_filter(
                base + lo, hi - lo, lo, tables.keys, tables.key_off.ctypes.data, tables.n_keys, tables.ent_begin.ctypes.data,
                tables.ent.ctypes.data, tables.alts, tables.alt_off.ctypes.data, tables.alt_len.ctypes.data, tables.n_alt,
                int(d_over), counts.ctypes.data, hs.ctypes.data if want_hits else None, ho.ctypes.data if want_hits else None,
                hl.ctypes.data if want_hits else None, cap, stats.ctypes.data)
            if rc == HITS_FULL:
                cap *= 4
                continue
            n = int(stats[0])
            return rc, counts, stats, (hs[:n], ho[:n], hl[:n]) if want_hits and rc == OK else None

    if parts == 1:
        res = [run(0)]
    else:
        with ThreadPoolExecutor(parts) as pool:
            res = list(pool.map(run, range(parts)))
    for rc, _c, stats, _h in res:                    # shards are in file order: the first failure is the reference's
        if rc == RAISES:
            raise OracleRaises(int(stats[3]))
        if rc == UNSUPPORTED:
            raise OracleUnsupported(int(stats[3]))
        if rc != OK:
            raise MemoryError("svjg_oracle_filter")
    counts = np.sum([c for _rc, c, _s, _h in res], axis=0, dtype=np.uint64).astype(np.uint32).reshape(num, 2)[:len(tables.sv_ids)]
    stats = {"n_hits": sum(int(s[0]) for _rc, _c, s, _h in res), "n_records": sum(int(s[1]) for _rc, _c, s, _h in res),
             "n_multi": sum(int(s[2]) for _rc, _c, s, _h in res)}
    if not want_hits:
        return counts, stats
    return (counts, stats, np.concatenate([h[0] for *_x, h in res]), np.concatenate([h[1] for *_x, h in res]),
            np.concatenate([h[2] for *_x, h in res]))


def counts_dict(tables, counts):
    """sv_id -> (n_ref, n_alt) for the ids with a hit: what ``hit_counts(filter_alignments(...))`` gives."""
    return {tables.sv_ids[i]: (int(counts[i, 0]), int(counts[i, 1])) for i in np.nonzero(counts.any(axis=1))[0]}
