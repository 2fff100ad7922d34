"""svjg/gzio.py: compressed inputs give the front-ends the bytes of the
uncompressed file — plain gzip, multi-member gzip and BGZF (block-parallel) — and damaged files are
refused."""
import gzip
import struct
import zlib

import numpy as np
import pytest

from conftest import read_golden
from svjg import gzio


def bgzf(data, block=65280, level=6):
    """bgzip's container: gzip members with a 'BC' extra field holding the block size, then the
    28-byte empty end-of-file block."""
    out = bytearray()
    for lo in list(range(0, len(data), block)) + [None]:
        chunk = b"" if lo is None else data[lo:lo + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        payload = c.compress(chunk) + c.flush()
        bsize = 12 + 6 + len(payload) + 8
        out += b"\x1f\x8b\x08\x04" + b"\0\0\0\0" + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += payload + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    return bytes(out)


@pytest.fixture(scope="module")
def gaf():
    return read_golden("s3.gaf.gz").encode()


def test_plain_file_is_read_as_is(tmp_path, gaf):
    p = tmp_path / "a.gaf"
    p.write_bytes(gaf[:5000])
    assert gzio.read_bytes(str(p)).tobytes() == gaf[:5000]
    p.write_bytes(b"")
    assert gzio.read_bytes(str(p)).size == 0
    p.write_bytes(b"\x1f")
    assert gzio.read_bytes(str(p)).tobytes() == b"\x1f"


def test_gzip_and_multi_member(tmp_path, gaf):
    p = tmp_path / "a.gaf.gz"
    p.write_bytes(gzip.compress(gaf))
    got = gzio.read_bytes(str(p))
    assert got.tobytes() == gaf and got.flags.writeable
    cut = len(gaf) // 3
    p.write_bytes(gzip.compress(gaf[:cut]) + gzip.compress(b"") + gzip.compress(gaf[cut:]) + b"\0" * 512)
    assert gzio.read_bytes(str(p)).tobytes() == gaf
    p.write_bytes(gzip.compress(b""))
    assert gzio.read_bytes(str(p)).size == 0


@pytest.mark.parametrize("threads", [1, 4])
def test_bgzf_blocks_in_parallel(tmp_path, gaf, threads):
    img = bgzf(gaf, block=4096)
    assert gzio._bgzf_blocks(memoryview(img)) is not None and len(gzio._bgzf_blocks(memoryview(img))) > 100
    p = tmp_path / "a.gaf.gz"
    p.write_bytes(img)
    assert gzio.read_bytes(str(p), threads=threads).tobytes() == gaf
    assert gzip.decompress(img) == gaf                       # the helper above writes what gzip reads
    # BGZF followed by a plain member is not BGZF end to end: the stream reader takes it
    p.write_bytes(img + gzip.compress(b"tail\n"))
    assert gzio.read_bytes(str(p)).tobytes() == gaf + b"tail\n"


def test_damaged_files_are_refused(tmp_path, gaf):
    p = tmp_path / "a.gaf.gz"
    z = gzip.compress(gaf)
    for bad in (z[:len(z) // 2], z[:-4] + b"\0\0\0\0", z + b"garbage", z[:40] + bytes(64) + z[104:]):
        p.write_bytes(bad)
        with pytest.raises(gzio.GzipError):
            gzio.read_bytes(str(p))
    img = bytearray(bgzf(gaf, block=4096))
    img[len(img) // 2] ^= 0x55                                # inside some block's payload
    p.write_bytes(bytes(img))
    with pytest.raises(gzio.GzipError):
        gzio.read_bytes(str(p))
    assert issubclass(gzio.GzipError, OSError)               # the front-ends' handlers turn it into exit status 1


def test_text_lines_like_open(tmp_path):
    text = "##h\r\n#c\tp\nchr1\t5\r\nlast"
    a, b = tmp_path / "v.vcf", tmp_path / "v.vcf.gz"
    a.write_bytes(text.encode())
    b.write_bytes(gzip.compress(text.encode()))
    want = ["##h\n", "#c\tp\n", "chr1\t5\n", "last"]
    assert gzio.read_text_lines(str(a)) == want == gzio.read_text_lines(str(b))


def test_reads_from_a_pipe(tmp_path, gaf):
    """`minigraph ... | filter-alignments.py -a /dev/stdin`: the reader must not ask a pipe for its size."""
    import os
    import threading
    fifo = str(tmp_path / "p.fifo")
    os.mkfifo(fifo)
    for payload in (gaf[:200000], gzip.compress(gaf[:200000])):
        def feed(data=payload):
            with open(fifo, "wb") as fh:
                fh.write(data)
        th = threading.Thread(target=feed)
        th.start()
        got = gzio.read_bytes(fifo)
        th.join()
        assert got.tobytes() == gaf[:200000] and got.flags.writeable
