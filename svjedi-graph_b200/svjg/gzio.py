"""Transparent gzip / BGZF input for the front-ends.

An extension, not reference behaviour: the reference opens its inputs as text and stops with a
UnicodeDecodeError on a compressed file.  Here a file that starts with the gzip magic is inflated
on the way in — BGZF (bgzip: independent blocks of at most 64 KiB that carry their size in a 'BC'
extra field) block-parallel on the host threads (zlib releases the GIL), plain gzip (one or more
members) as one stream — and everything behind the reader sees the bytes of the uncompressed file."""
from __future__ import annotations

import os
import stat
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

MAGIC = b"\x1f\x8b"


class GzipError(OSError):
    """Damaged compressed input (the front-ends exit with status 1)."""


def is_gzip(head):
    return bytes(head[:2]) == MAGIC


def _bgzf_blocks(buf):
    """[(payload offset, payload length, inflated size)] if ``buf`` is BGZF from end to end, else None."""
    n = len(buf)
    pos, out = 0, []
    while pos < n:
        if n - pos < 18 or buf[pos:pos + 4] != b"\x1f\x8b\x08\x04":
            return None
        xlen = struct.unpack_from("<H", buf, pos + 10)[0]
        x, xend, bsize = pos + 12, pos + 12 + xlen, None
        if xend > n:
            return None
        while x + 4 <= xend:
            si1, si2, slen = buf[x], buf[x + 1], struct.unpack_from("<H", buf, x + 2)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", buf, x + 4)[0] + 1
            x += 4 + slen
        if bsize is None or bsize < xlen + 20 or pos + bsize > n:
            return None
        isize = struct.unpack_from("<I", buf, pos + bsize - 4)[0]
        if isize > 65536:
            return None
        out.append((xend, pos + bsize - 8 - xend, isize))
        pos += bsize
    return out


def _inflate_bgzf(buf, blocks, threads):
    sizes = np.fromiter((b[2] for b in blocks), dtype=np.int64, count=len(blocks))
    starts = np.concatenate(([0], np.cumsum(sizes)))
    out = np.empty(int(starts[-1]), dtype=np.uint8)
    view = memoryview(out)
    src = memoryview(buf)
    step = max(1, min(256, len(blocks) // (threads * 4) or 1))

    def work(lo):
        for i in range(lo, min(lo + step, len(blocks))):
            off, ln, isize = blocks[i]
            data = zlib.decompress(src[off:off + ln], -15, isize or 1)
            if len(data) != isize or zlib.crc32(data) != struct.unpack_from("<I", src, off + ln)[0]:
                raise GzipError(f"BGZF block {i}: length or CRC-32 does not match the block trailer")
            view[starts[i]:starts[i] + isize] = data

    try:
        if threads <= 1 or len(blocks) < 8:
            for lo in range(0, len(blocks), step):
                work(lo)
        else:
            with ThreadPoolExecutor(threads) as pool:
                list(pool.map(work, range(0, len(blocks), step)))
    except zlib.error as exc:
        raise GzipError(f"BGZF: {exc}") from None
    return out


def _inflate_stream(buf):
    """One or more gzip members back to back (RFC 1952), CRC and length checked by zlib."""
    chunks, src = [], memoryview(buf)
    pos = 0
    try:
        while pos < len(src):
            d = zlib.decompressobj(31)
            while pos < len(src) and not d.eof:
                piece = src[pos:pos + (1 << 24)]
                chunks.append(d.decompress(piece))
                pos += len(piece) - len(d.unused_data)
            if not d.eof:
                raise GzipError("gzip: input ends inside a member")
            rest = bytes(src[pos:pos + 2])
            if rest and rest != MAGIC:
                if bytes(src[pos:]).strip(b"\0") == b"":        # zero padding after the last member (gzip(1) accepts it)
                    break
                raise GzipError("gzip: data after the last member")
    except zlib.error as exc:
        raise GzipError(f"gzip: {exc}") from None
    return np.frombuffer(bytearray().join(chunks), dtype=np.uint8)


def inflate(buf, threads=None):
    """Uncompressed bytes (uint8 array) of a gzip / BGZF image held in ``buf`` (bytes-like)."""
    buf = memoryview(buf).cast("B")
    threads = threads or min(32, os.cpu_count() or 1)
    blocks = _bgzf_blocks(buf)
    if blocks is not None:
        return _inflate_bgzf(buf, blocks, threads)
    return _inflate_stream(buf)


def read_bytes(path, threads=None):
    """The file as a uint8 array; inflated when it starts with the gzip magic."""
    if stat.S_ISREG(os.stat(path).st_mode):
        raw = np.fromfile(path, dtype=np.uint8)
    else:                                            # a pipe (`minigraph ... | filter-alignments.py -a /dev/stdin`): no size to ask for
        with open(path, "rb") as fh:
            raw = np.frombuffer(bytearray(fh.read()), dtype=np.uint8)
    if raw.size >= 2 and is_gzip(raw):
        return inflate(raw, threads)
    return raw


def read_text_lines(path):
    """``open(path).readlines()`` (text mode: universal newlines, locale encoding) for a plain or
    compressed file — what predict-genotype.py:95-96 iterates over."""
    import io
    return io.TextIOWrapper(io.BytesIO(read_bytes(path).tobytes())).readlines()
