"""Host mirror of filter-alignments.py (reference :27-175): graph tables, the
filter over a GAF, and the informative_aln.json writer — all through libsvjg.so.
PyTorch is used only to own device / pinned buffers and streams."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

D_OVER = 100    # filter-alignments.py:56


class InputError(Exception):
    """The reference raises on this input (exit status 1)."""


class Tables:
    """link key -> SV entries and alt-node lengths (filter-alignments.py:95-113)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        self.num_sv = int(capi.lib.svjg_tables_num_sv(self._h))
        self.num_links = int(capi.lib.svjg_tables_num_links(self._h))
        self.num_alt_nodes = int(capi.lib.svjg_tables_num_alt_nodes(self._h))
        self.device = None
        self._ids = None

    @classmethod
    def load(cls, svs_edges_path, gfa_path):
        h = C.c_void_p()
        capi.check(capi.lib.svjg_tables_load(os.fsencode(svs_edges_path), os.fsencode(gfa_path), C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_memory(cls, edges_json, gfa_text):
        ej = edges_json.encode() if isinstance(edges_json, str) else bytes(edges_json)
        gt = gfa_text.encode() if isinstance(gfa_text, str) else bytes(gfa_text)
        h = C.c_void_p()
        capi.check(capi.lib.svjg_tables_from_memory(ej, len(ej), gt, len(gt), C.byref(h)))
        return cls(h.value)

    def to_device(self, device=0):
        capi.check(capi.lib.svjg_tables_to_device(self._h, int(device)))
        self.device = int(device)
        return self

    def clone(self):
        """A second handle with the same host image and no device image (one handle per GPU)."""
        h = C.c_void_p()
        capi.check(capi.lib.svjg_tables_clone(self._h, C.byref(h)))
        return Tables(h.value)

    def set_flags(self, flags):
        capi.check(capi.lib.svjg_tables_set_flags(self._h, int(flags)))
        return self

    @property
    def device_bytes(self):
        return int(capi.lib.svjg_tables_device_bytes(self._h))

    def sv_id(self, i):
        n = C.c_uint32()
        p = capi.lib.svjg_tables_sv_id(self._h, i, C.byref(n))
        if not p:
            raise IndexError(i)
        return C.string_at(p, n.value).decode("utf-8")

    @property
    def sv_ids(self):
        if self._ids is None:
            self._ids = [self.sv_id(i) for i in range(self.num_sv)]
        return self._ids

    def find_sv(self, key):
        b = key.encode("utf-8")
        i = capi.lib.svjg_tables_find_sv(self._h, b, len(b))
        return None if i == capi.NO_SV else int(i)

    @property
    def image_hash(self):
        return int(capi.lib.svjg_tables_image_hash(self._h))

    def alt_node_len(self, name):
        """``alt_node_len.get(name)`` of filter-alignments.py:103-113, read from the tables the kernels probe."""
        b = name.encode("utf-8")
        n = int(capi.lib.svjg_tables_alt_node_len(self._h, b, len(b)))
        if n == -2:
            raise capi.SvjgError("node tables disagree on " + repr(name))
        return None if n < 0 else n

    def close(self):
        if self._h:
            capi.lib.svjg_tables_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FilterResult:
    def __init__(self, counts, stats, hit_sv2=None, hit_off=None, hit_len=None):
        self.counts = counts          # np.uint32 [num_sv, 2]  (REF, ALT hit multiplicities)
        self.stats = stats
        self.hit_sv2, self.hit_off, self.hit_len = hit_sv2, hit_off, hit_len

    @property
    def n_hits(self):
        return self.stats["n_hits"]


def _as_u8(buf):
    """numpy uint8 view (no copy) of bytes / bytearray / memoryview / ndarray / pinned torch tensor."""
    if isinstance(buf, (PinnedBytes, RegisteredBytes)):
        a = buf.array
    elif isinstance(buf, np.ndarray):
        a = buf
    elif hasattr(buf, "numpy") and hasattr(buf, "is_pinned"):
        a = buf.numpy()
    else:
        a = np.frombuffer(buf, dtype=np.uint8)
    if a.dtype != np.uint8 or not a.flags.c_contiguous:
        raise TypeError("GAF buffer must be contiguous bytes")
    return a


def translate_newlines(buf):
    """The reference reads the GAF in text mode (filter-alignments.py:123): "\r\n" and a lone "\r" are
    line ends too and become "\n" in the stored lines.  The kernels split at "\n" only, so a buffer
    that contains a carriage return is translated first (one copy, one pass: svjg_translate_newlines);
    any other buffer is returned as it is.  One memchr over the bytes."""
    a = _as_u8(buf)
    if a.size == 0:
        return buf
    libc = C.CDLL(None)
    libc.memchr.restype = C.c_void_p
    libc.memchr.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
    if not libc.memchr(a.ctypes.data, 13, a.size):
        return buf
    out = a.copy()                                          # the caller's bytes stay as they are
    n = int(capi.lib.svjg_translate_newlines(out.ctypes.data, out.size))
    return out[:n]


def _raise_input(stats, base=0):
    reason = capi.BAD_REASONS.get(stats["status"], "malformed input")
    err = InputError(f"GAF line at byte {stats['err_offset'] + base}: {reason} (the reference raises here)")
    err.stats = dict(stats)
    raise err


class HostBuffers:
    """Pinned host arrays for the results of filter_host(): the counters and the hit list come back at
    PCIe speed instead of through the driver's staging copies.  Reusable: every filter_host(out=...)
    call overwrites them, so the result of the previous call must have been consumed."""

    def __init__(self, tables, hit_cap):
        import torch
        self.hit_cap = int(hit_cap)
        self._t = [torch.zeros((max(1, tables.num_sv), 2), dtype=torch.int32).pin_memory(),
                   torch.empty(max(1, hit_cap), dtype=torch.int32).pin_memory(),
                   torch.empty(max(1, hit_cap), dtype=torch.int64).pin_memory(),
                   torch.empty(max(1, hit_cap), dtype=torch.int32).pin_memory()]
        self.counts = self._t[0].numpy().view(np.uint32)[: tables.num_sv]
        self.sv2 = self._t[1].numpy().view(np.uint32)
        self.off = self._t[2].numpy().view(np.uint64)
        self.len = self._t[3].numpy().view(np.uint32)


def filter_host(tables, gaf, d_over=D_OVER, want_hits=True, hit_cap=None, out=None):
    """The per-line loop of filter-alignments.py:123-166 over GAF bytes in HOST
    memory (pinned memory copies fastest).  Returns counts, stats and, if
    ``want_hits``, the hit list with absolute byte offsets.  ``out``: HostBuffers
    to receive them (pinned, reused); otherwise fresh numpy arrays."""
    if tables.device is None:
        raise RuntimeError("tables.to_device() first")
    a = _as_u8(gaf)
    n = int(a.size)
    counts = out.counts if out is not None else np.zeros((tables.num_sv, 2), dtype=np.uint32)
    stats = capi.FilterStats()
    if out is not None:
        hit_cap = out.hit_cap if want_hits else 0
    elif hit_cap is None:
        hit_cap = max(1024, n // 64) if want_hits else 0
    while True:
        if hit_cap and out is not None:
            sv2, off, ln = out.sv2, out.off, out.len
            ptrs = (sv2.ctypes.data, off.ctypes.data, ln.ctypes.data)
        elif hit_cap:
            sv2 = np.empty(hit_cap, dtype=np.uint32)
            off = np.empty(hit_cap, dtype=np.uint64)
            ln = np.empty(hit_cap, dtype=np.uint32)
            ptrs = (sv2.ctypes.data, off.ctypes.data, ln.ctypes.data)
        else:
            sv2 = off = ln = None
            ptrs = (None, None, None)
        rc = capi.lib.svjg_filter_host(tables._h, a.ctypes.data if n else None, n, int(d_over),
                                       counts.ctypes.data, *ptrs, hit_cap, C.byref(stats))
        st = stats.as_dict()
        if rc == capi.E_HITS_OVERFLOW:
            if out is not None:
                raise RuntimeError(f"HostBuffers hold {out.hit_cap} hits, the batch has {st['n_hits']}")
            hit_cap = int(st["n_hits"]) + 16
            continue
        if rc == capi.E_INPUT:
            _raise_input(st)
        capi.check(rc)
        break
    nh = st["n_hits"] if hit_cap else 0
    if hit_cap:
        return FilterResult(counts, st, sv2[:nh], off[:nh], ln[:nh])
    return FilterResult(counts, st)


def apply_min_identity(tables, gaf, result, min_identity):
    """Extension (off by default): drops the hits whose alignment has Aid < ``min_identity``, Aid as
    filter-alignments.py:193-196 parses it (float() behind the last "id:f:", else Am / Alen); the counters follow.
    InputError where float() raises on a stored line."""
    a = _as_u8(gaf)
    sv2 = np.ascontiguousarray(result.hit_sv2, dtype=np.uint32).copy()
    off = np.ascontiguousarray(result.hit_off, dtype=np.uint64).copy()
    ln = np.ascontiguousarray(result.hit_len, dtype=np.uint32).copy()
    counts = np.ascontiguousarray(result.counts, dtype=np.uint32).copy()
    n = C.c_uint64(int(sv2.size))
    rc = capi.lib.svjg_hits_min_identity(a.ctypes.data if a.size else None, int(a.size), sv2.ctypes.data, off.ctypes.data, ln.ctypes.data,
                                         C.byref(n), float(min_identity), counts.ctypes.data, tables.num_sv)
    if rc == capi.E_INPUT:
        raise InputError(capi.lib.svjg_last_error().decode("utf-8", "replace"))
    capi.check(rc)
    k = int(n.value)
    stats = dict(result.stats, n_hits=k)
    return FilterResult(counts, stats, sv2[:k], off[:k], ln[:k])


def filter_host_multi(tables_by_device, gaf, d_over=D_OVER):
    """The per-line loop over ONE buffer on several GPUs of the node (the stage boundary svjedi-graph.py:113-118
    with the records sharded): the buffer is cut by bytes at line ends (shard.shard_cuts), device k filters range k
    from the host (svjg_filter_host on a thread of its own; the tables are replicated, one handle per device), the
    counters are summed and the hit lists concatenated in range order = file order, offsets made absolute.
    The result equals filter_host on one device."""
    from concurrent.futures import ThreadPoolExecutor
    from . import shard
    a = _as_u8(gaf)
    world = len(tables_by_device)
    cuts = shard.shard_cuts(a, world)
    def one(k):
        try:
            return filter_host(tables_by_device[k], a[cuts[k]:cuts[k + 1]], d_over=d_over)
        except InputError as exc:                         # offsets of a shard are its own: make them the file's
            _raise_input(exc.stats, base=cuts[k])

    with ThreadPoolExecutor(world) as pool:
        parts = list(pool.map(one, range(world)))          # in range order: the first failure is the reference's
    counts = np.zeros((tables_by_device[0].num_sv, 2), dtype=np.uint32)
    stats = {}
    for k, r in enumerate(parts):
        counts += r.counts
        for key, v in r.stats.items():
            if key == "err_offset":
                continue
            stats[key] = stats.get(key, 0) + v
    stats["status"] = 0
    stats["err_offset"] = parts[0].stats["err_offset"]
    sv2 = np.concatenate([r.hit_sv2 for r in parts])
    off = np.concatenate([r.hit_off.astype(np.uint64) + np.uint64(cuts[k]) for k, r in enumerate(parts)])
    ln = np.concatenate([r.hit_len for r in parts])
    return FilterResult(counts, stats, sv2, off, ln)


def filter_json_multi_begin(tables_by_device, gaf, d_over=D_OVER):
    """:func:`filter_json_begin` for ONE buffer on several GPUs of the node: range k (shard.shard_cuts) is uploaded
    to and filtered on device k (svjg_filter_json_begin_at on a thread of its own) and stays there; the hit tuples
    are gathered on device 0 and the counters summed (svjg_filter_json_gather), so that
    ``filter_json_finish(tables_by_device[0])`` / ``filter_json_write(tables_by_device[0], path)`` render the whole
    text there, reading the lines of the other ranges from their devices over NVLink.  Returns the FilterResult
    (summed counters and stats), or None where this route is not available (a range that does not fit its device,
    no peer access): use :func:`filter_host_multi`."""
    from concurrent.futures import ThreadPoolExecutor
    from . import shard
    a = _as_u8(gaf)
    world = len(tables_by_device)
    cuts = shard.shard_cuts(a, world)
    num_sv = tables_by_device[0].num_sv

    def one(k):
        t = tables_by_device[k]
        part = a[cuts[k]:cuts[k + 1]]
        counts = np.zeros((num_sv, 2), dtype=np.uint32)
        stats = capi.FilterStats()
        rc = capi.lib.svjg_filter_json_begin_at(t._h, part.ctypes.data if part.size else None, int(part.size), int(cuts[k]),
                                                int(d_over), counts.ctypes.data, C.byref(stats))
        st = stats.as_dict()
        if rc == capi.E_INPUT:
            _raise_input(st)                             # err_offset is already the file's
        if rc == capi.E_UNSUPPORTED:
            return None
        capi.check(rc)
        return counts, st

    with ThreadPoolExecutor(world) as pool:
        parts = list(pool.map(one, range(world)))          # in range order: the first failure is the reference's
    if any(p is None for p in parts):
        return None
    counts = np.zeros((num_sv, 2), dtype=np.uint32)
    stats = {}
    for c, st in parts:
        counts += c
        for key, v in st.items():
            if key != "err_offset":
                stats[key] = stats.get(key, 0) + v
    stats["status"] = 0
    stats["err_offset"] = parts[0][1]["err_offset"]
    handles = (C.c_void_p * world)(*[t._h.value for t in tables_by_device])
    rc = capi.lib.svjg_filter_json_gather(handles, world, counts.ctypes.data)
    if rc == capi.E_UNSUPPORTED:
        return None
    capi.check(rc)
    return FilterResult(counts, stats)


def filter_json_host(tables, gaf, d_over=D_OVER, counts=None):
    """svjg_filter_json_host: the filter over GAF bytes in HOST memory with ``informative_aln.json``
    (filter-alignments.py:160-175) rendered on the device.  Returns (FilterResult without a hit list, memoryview
    of the JSON text: it lives in a buffer of ``tables`` and is valid until the next call), or None in place
    of the text when the device renderer declines (non-ASCII bytes in a stored line, a list beyond 64 Ki
    entries): the caller then runs filter_host + write_informative_json."""
    if tables.device is None:
        raise RuntimeError("tables.to_device() first")
    a = _as_u8(gaf)
    n = int(a.size)
    counts = counts if counts is not None else np.zeros((tables.num_sv, 2), dtype=np.uint32)
    stats = capi.FilterStats()
    p, ln = C.c_void_p(), C.c_uint64()
    rc = capi.lib.svjg_filter_json_host(tables._h, a.ctypes.data if n else None, n, int(d_over), counts.ctypes.data,
                                        C.byref(stats), C.byref(p), C.byref(ln))
    st = stats.as_dict()
    if rc == capi.E_INPUT:
        _raise_input(st)
    if rc == capi.E_UNSUPPORTED:
        return FilterResult(counts, st), None
    capi.check(rc)
    text = memoryview((C.c_char * ln.value).from_address(p.value)) if ln.value else memoryview(b"")
    return FilterResult(counts, st), text


def filter_json_begin(tables, gaf, d_over=D_OVER, counts=None):
    """First half of :func:`filter_json_host` (svjg_filter_json_begin): upload + filter; returns the FilterResult
    (counters, stats) as soon as they are on the host, or None where the file does not fit the device (use
    :func:`filter_host`).  The caller may genotype from the counters while :func:`filter_json_finish` -- on
    another thread if it likes -- renders the text and waits for it."""
    if tables.device is None:
        raise RuntimeError("tables.to_device() first")
    a = _as_u8(gaf)
    n = int(a.size)
    counts = counts if counts is not None else np.zeros((tables.num_sv, 2), dtype=np.uint32)
    stats = capi.FilterStats()
    rc = capi.lib.svjg_filter_json_begin(tables._h, a.ctypes.data if n else None, n, int(d_over), counts.ctypes.data, C.byref(stats))
    st = stats.as_dict()
    if rc == capi.E_INPUT:
        _raise_input(st)
    if rc == capi.E_UNSUPPORTED:                      # the file does not fit the device: filter_host streams it
        return None
    capi.check(rc)
    return FilterResult(counts, st)


def filter_json_finish(tables):
    """Second half: the JSON text (memoryview into a buffer of ``tables``), or None where the device renderer declines."""
    p, ln = C.c_void_p(), C.c_uint64()
    rc = capi.lib.svjg_filter_json_finish(tables._h, C.byref(p), C.byref(ln))
    if rc == capi.E_UNSUPPORTED:
        return None
    capi.check(rc)
    return memoryview((C.c_char * ln.value).from_address(p.value)) if ln.value else memoryview(b"")


def filter_json_write(tables, path, slice_bytes=0):
    """Second half, to a file (svjg_filter_json_write): the text leaves the device in slices of whole keys that are
    written while the next one is rendered and copied; neither side holds the whole text.  Returns the bytes
    written, or None where the device renderer declines (nothing is written then)."""
    ln = C.c_uint64()
    rc = capi.lib.svjg_filter_json_write(tables._h, os.fsencode(path), int(slice_bytes), C.byref(ln))
    if rc == capi.E_UNSUPPORTED:
        return None
    capi.check(rc)
    return int(ln.value)


def filter_stream(tables, fileobj, chunk_bytes=16 << 20, d_over=D_OVER):
    """the filter fed from a pipe (``minigraph ... | filter-alignments.py -a
    /dev/stdin``) while the mapper is still writing.  The stream is read straight into PAGE-LOCKED buffers
    (three of them in turn: one being read into, one being filtered, one spare) and cut into segments of whole
    lines of about ``chunk_bytes``; each segment goes through :func:`filter_host` on a worker thread (ctypes
    releases the GIL) while the next one is being read, its counters are added up and its hit offsets
    moved to their place in the whole input.  Text-mode line ends are translated per segment exactly as
    for a file (a "\r" at the end of a read is held back until the next byte is known).  Returns
    (FilterResult, the whole translated GAF as one uint8 array) — what the JSON writer needs."""
    from concurrent.futures import ThreadPoolExecutor
    segments, results = [], []
    pool = ThreadPoolExecutor(1)
    jobs = []                                              # (future, ring slot) in flight, oldest first
    base = 0
    cap = 2 * chunk_bytes                                  # a segment is cut once chunk_bytes are there; a longer line: grow()
    keep, bufs = [], [None, None, None]

    def alloc(n):
        try:
            keep.append(PinnedBytes(n))                    # page-locked; alive until the jobs that read it are done
            return keep[-1].array
        except capi.SvjgError:                             # no device (the CPU tests of the cutting logic)
            return np.empty(n, np.uint8)

    def buffer(k):
        """ring buffer k, made when first needed (a short stream needs one)"""
        if bufs[k] is None or bufs[k].size < cap:
            bufs[k] = alloc(cap)
        return bufs[k]

    def grow():
        """a line longer than a buffer: twice the room from here on, what is there moves along"""
        nonlocal cap
        old = bufs[slot]
        cap *= 2
        bufs[slot] = alloc(cap)
        bufs[slot][:fill] = old[:fill]
    slot, fill = 0, 0                                      # bufs[slot][:fill]: bytes read and not yet submitted
    held_cr = False

    def one_segment(arr, b):
        try:
            return b, filter_host(tables, arr, d_over)
        except InputError as exc:
            raise InputError(f"{exc} [offset within the segment that starts at byte {b} of the stream]") from None

    def reap(upto):
        while len(jobs) > upto:
            results.append(jobs.pop(0)[0].result())        # raises InputError where the reference raises

    def submit(n):
        """bufs[slot][:n] is a segment; what lies behind it moves to the head of the next free buffer"""
        nonlocal slot, fill, base
        cur = buffer(slot)
        segments.append(cur[:n].copy())                    # the JSON writer needs the whole text at the end
        reap(1)                                            # at most two segments in flight: the third buffer is free
        jobs.append((pool.submit(one_segment, cur[:n], base), slot))
        base += n
        nxt = (slot + 1) % 3
        rest = fill - n
        if rest:
            buffer(nxt)[:rest] = cur[n:fill]
        slot, fill = nxt, rest

    try:
        eof = False
        readinto = getattr(fileobj, "readinto", None)
        while not eof:
            cur = buffer(slot)
            room = memoryview(cur)[fill + (1 if held_cr else 0):min(cap, fill + chunk_bytes)]
            if readinto is not None:
                got = readinto(room) or 0
            else:
                block = fileobj.read(chunk_bytes)
                got = len(block)
                room[:got] = block
            eof = got == 0
            new0 = fill
            if held_cr:                                    # the "\r" held back belongs in front of these bytes (room was left)
                cur[fill] = 13
                got += 1
                held_cr = False
            fill += got
            if not eof and fill > new0 and cur[fill - 1] == 13:     # "\r\n" may straddle two reads
                fill -= 1
                held_cr = True
            if fill > new0 and (cur[new0:fill] == 13).any():
                t = np.frombuffer(bytes(cur[new0:fill]).replace(b"\r\n", b"\n").replace(b"\r", b"\n"), dtype=np.uint8)
                cur[new0:new0 + t.size] = t
                fill = new0 + t.size
            if eof:
                cut = fill
            else:
                nl = np.flatnonzero(cur[:fill][::-1] == 10)
                cut = fill - int(nl[0]) if nl.size else 0   # whole lines only; the rest waits for more bytes
            if cut and (eof or fill >= chunk_bytes):
                submit(cut)
            if fill >= cap - 1:
                grow()
        reap(0)
    finally:
        pool.shutdown(wait=True)
    counts = np.zeros((tables.num_sv, 2), dtype=np.uint32)
    stats = {}
    sv2, off, ln = [], [], []
    for b, r in results:
        counts += r.counts
        for k, v in r.stats.items():
            stats[k] = stats.get(k, 0) + v if k not in ("status", "err_offset") else 0
        sv2.append(r.hit_sv2)
        off.append(r.hit_off.astype(np.uint64) + np.uint64(b))
        ln.append(r.hit_len)
    cat = (lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt))
    gaf = np.concatenate(segments) if segments else np.zeros(0, np.uint8)
    for k in ("n_hits", "n_records", "n_multi", "n_checks", "status", "err_offset", "n_generic", "n_exact"):
        stats.setdefault(k, 0)
    return FilterResult(counts, stats, cat(sv2, np.uint32), cat(off, np.uint64), cat(ln, np.uint32)), gaf


class DeviceFilter:
    """Device-resident variant: GAF shard already in HBM (torch uint8 tensor).
    Buffers are torch tensors; launches go to torch's current stream."""

    def __init__(self, tables, hit_cap=1 << 20, device=None):
        import torch
        self.torch = torch
        self.tables = tables
        self.dev = torch.device("cuda", tables.device if device is None else device)
        self.counts = torch.zeros((max(1, tables.num_sv), 2), dtype=torch.int32, device=self.dev)
        self.stats = torch.zeros(8, dtype=torch.int64, device=self.dev)
        self._alloc_hits(hit_cap)

    def _alloc_hits(self, cap):
        t = self.torch
        self.hit_cap = int(cap)
        self.hit_sv2 = t.empty(max(1, cap), dtype=t.int32, device=self.dev)
        self.hit_off = t.empty(max(1, cap), dtype=t.int32, device=self.dev)
        self.hit_len = t.empty(max(1, cap), dtype=t.int32, device=self.dev)

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def reset(self):
        capi.check(capi.lib.svjg_filter_reset(self.counts.data_ptr(), self.tables.num_sv, self.stats.data_ptr(),
                                              self._stream()))

    def run(self, d_gaf, n_bytes=None, base_offset=0, d_over=D_OVER):
        """Asynchronous: accumulates into self.counts / self.stats / hit arrays."""
        n = int(d_gaf.numel() if n_bytes is None else n_bytes)
        capi.check(capi.lib.svjg_filter_device(self.tables._h, d_gaf.data_ptr(), n, int(base_offset), int(d_over),
                                               self.counts.data_ptr(), self.hit_sv2.data_ptr(), self.hit_off.data_ptr(),
                                               self.hit_len.data_ptr(), self.hit_cap, self.stats.data_ptr(),
                                               self._stream()))

    def read_stats(self):
        v = self.stats.cpu().numpy().astype(np.uint64)
        names = [n for n, _ in capi.FilterStats._fields_]
        return {k: int(x) for k, x in zip(names, v)}

    def result(self):
        """Synchronises and copies everything back (hits as absolute offsets
        within the shard passed to run())."""
        st = self.read_stats()
        if st["status"]:
            _raise_input(st)
        nh = st["n_hits"]
        if nh > self.hit_cap:
            # the cursor counts every hit, the arrays hold the first hit_cap of them: a list cut short
            # would print an informative_aln.json with alignments missing
            raise capi.SvjgError(capi.E_HITS_OVERFLOW, f"hit buffers too small: {nh} hits, room for {self.hit_cap} "
                                                       "(DeviceFilter(hit_cap=...); counters and stats are complete)")
        counts = self.counts.cpu().numpy().view(np.uint32)[: self.tables.num_sv]
        return FilterResult(counts, st, self.hit_sv2[:nh].cpu().numpy().view(np.uint32),
                            self.hit_off[:nh].cpu().numpy().view(np.uint32).astype(np.uint64),
                            self.hit_len[:nh].cpu().numpy().view(np.uint32))


def write_informative_json(tables, gaf, result, out_path):
    """filter-alignments.py:174-175, byte for byte."""
    a = _as_u8(gaf)
    nh = int(result.hit_sv2.size)
    sv2 = np.ascontiguousarray(result.hit_sv2, dtype=np.uint32)
    off = np.ascontiguousarray(result.hit_off, dtype=np.uint64)
    ln = np.ascontiguousarray(result.hit_len, dtype=np.uint32)
    capi.check(capi.lib.svjg_emit_informative_json(
        tables._h, a.ctypes.data if a.size else None, int(a.size), sv2.ctypes.data if nh else None,
        off.ctypes.data if nh else None, ln.ctypes.data if nh else None, nh, os.fsencode(out_path)))


class JsonText:
    """``informative_aln.json`` in memory (svjg_emit_informative_json_mem): ``.view`` is a memoryview of the
    text, valid until the object goes away."""

    def __init__(self, tables, gaf, result):
        a = _as_u8(gaf)
        nh = int(result.hit_sv2.size)
        sv2 = np.ascontiguousarray(result.hit_sv2, dtype=np.uint32)
        off = np.ascontiguousarray(result.hit_off, dtype=np.uint64)
        ln = np.ascontiguousarray(result.hit_len, dtype=np.uint32)
        self._p, n = C.c_void_p(), C.c_uint64()
        capi.check(capi.lib.svjg_emit_informative_json_mem(
            tables._h, a.ctypes.data if a.size else None, int(a.size), sv2.ctypes.data if nh else None,
            off.ctypes.data if nh else None, ln.ctypes.data if nh else None, nh, C.byref(self._p), C.byref(n)))
        self.nbytes = int(n.value)
        self.view = memoryview((C.c_char * self.nbytes).from_address(self._p.value)) if self.nbytes else memoryview(b"")

    def __del__(self):
        try:
            if self._p:
                self.view = None
                capi.lib.svjg_buffer_free(self._p)
                self._p = None
        except Exception:
            pass


class PinnedBytes:
    """Page-locked host bytes from libsvjg (cudaHostAlloc) -- no tensor library needed.  ``.array`` is the
    numpy uint8 view; the memory is released with the object."""

    def __init__(self, n):
        self._p = C.c_void_p()
        capi.check(capi.lib.svjg_host_alloc(max(int(n), 1), C.byref(self._p)))
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_uint8)), shape=(max(int(n), 1),))[:int(n)]

    def __del__(self):
        try:
            if self._p:
                capi.lib.svjg_host_free(self._p)
                self._p = None
        except Exception:
            pass


class RegisteredBytes:
    """A numpy uint8 array page-locked in place (cudaHostRegister) for the lifetime of the object."""

    def __init__(self, array):
        self.array = array
        self._reg = False
        if array.size:
            capi.check(capi.lib.svjg_host_register(array.ctypes.data, array.size))
            self._reg = True

    def __del__(self):
        try:
            if self._reg:
                capi.lib.svjg_host_unregister(self.array.ctypes.data)
                self._reg = False
        except Exception:
            pass


def read_file_pinned(path):
    """Whole file into page-locked host memory; returns a PinnedBytes (pass ``.array`` on)."""
    n = os.path.getsize(path)
    buf = PinnedBytes(n)
    with open(path, "rb", buffering=0) as fh:
        view = memoryview(buf.array)
        got = 0
        while got < n:
            k = fh.readinto(view[got:])
            if not k:
                break
            got += k
    return buf
