"""Compressed input on the host (csrc/pgz.cu): BGZF members and single-member gzip streams are inflated in parallel; the bytes must
be exactly what zlib produces.  CPU only (catt_inflate_file needs no GPU); the GPU test runs a .fastq.gz through catt_run_file."""
import functools
import gzip
import os
import random
import zlib

import numpy as np
import pytest

from oracle import bamio


def fastq_like(seed, n):
    rng = np.random.default_rng(seed)
    out = []
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    ql = np.frombuffer(b"IIIIIFFF:,#", dtype=np.uint8)
    for i in range(n):
        ln = int(rng.choice([50, 100, 150, 151]))
        s = lut[rng.integers(0, 4, ln)].tobytes()
        if rng.random() < 0.3:
            s = s[:ln - 40] + s[:40]                       # an internal repeat: matches inside the record
        if rng.random() < 0.1:
            s += b"N"
        q = ql[rng.integers(0, len(ql), len(s))].tobytes()
        out.append(b"@SRR11799691.%d %d/1\n%s\n+\n%s\n" % (i, i, s, q))
    return b"".join(out)


@functools.lru_cache(maxsize=None)
def datasets_cached():
    rng = random.Random(5)
    return {
        "fastq": fastq_like(1, 20000),
        "repetitive": (b"ACGT" * 5 + b"\n") * 200000 + b"x" * 100000 + bytes(range(256)) * 3000,
        "random": rng.randbytes(1500000),                # mostly stored blocks
        "mixed": fastq_like(2, 5000) + rng.randbytes(300000) + fastq_like(3, 5000),
    }


def datasets():
    return datasets_cached()


def gz_bytes(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, name=None):
    co = zlib.compressobj(level, zlib.DEFLATED, 31, 9, strategy)
    return co.compress(data) + co.flush()


def inflate(path, tmp_path, chunk=None, zlib_only=False):
    from catt_b200.host import inflate_file
    out = str(tmp_path / "out.bin")
    for k in ("CATT_PGZ_CHUNK", "CATT_ZLIB_ONLY"):
        os.environ.pop(k, None)
    if chunk:
        os.environ["CATT_PGZ_CHUNK"] = str(chunk)
    if zlib_only:
        os.environ["CATT_ZLIB_ONLY"] = "1"
    try:
        m = inflate_file(path, out)
    finally:
        for k in ("CATT_PGZ_CHUNK", "CATT_ZLIB_ONLY"):
            os.environ.pop(k, None)
    return m, open(out, "rb").read()


@pytest.mark.parametrize("kind", ["fastq", "repetitive", "random", "mixed"])
@pytest.mark.parametrize("level,strategy", [(1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                            (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE)])
def test_parallel_gzip_equals_zlib(tmp_path, kind, level, strategy):
    data = datasets()[kind]
    p = str(tmp_path / "a.gz")
    open(p, "wb").write(gz_bytes(data, level, strategy))
    m, got = inflate(p, tmp_path, chunk=65536)
    assert got == data
    # fixed-Huffman-only streams have no dynamic block to start a chunk at: every chunk but the first is absent, which is still the
    # parallel path (one chunk); everything else must really have been split
    assert m == 2
    m3, got3 = inflate(p, tmp_path, zlib_only=True)
    assert m3 == 3 and got3 == data


def test_chunk_sizes_and_threads(tmp_path):
    from catt_b200.host import set_host_threads
    data = datasets()["fastq"]
    p = str(tmp_path / "a.gz")
    open(p, "wb").write(gz_bytes(data, 6))
    try:
        for threads in (1, 3, 8):
            set_host_threads(threads)
            for chunk in (65536, 100000, 1 << 20, 1 << 26):
                m, got = inflate(p, tmp_path, chunk=chunk)
                assert (m, got == data) == (2, True), (threads, chunk)
    finally:
        set_host_threads(0)


def test_gzip_header_fields_and_python_gzip_module(tmp_path):
    data = datasets()["fastq"]
    p = str(tmp_path / "named.fastq.gz")
    with gzip.GzipFile(p, "wb", compresslevel=6, mtime=12345) as fh:       # FNAME field in the header
        fh.write(data)
    m, got = inflate(p, tmp_path, chunk=65536)
    assert (m, got == data) == (2, True)
    raw = open(p, "rb").read()
    # FEXTRA + FCOMMENT + FHCRC by hand around the same deflate stream
    hdr_end = raw.index(b"\0", 10) + 1
    extra = b"XY\x03\x00abc"
    hdr = b"\x1f\x8b\x08" + bytes([4 | 16 | 2]) + raw[4:10] + len(extra).to_bytes(2, "little") + extra + b"a comment\0"
    hdr += (zlib.crc32(hdr) & 0xFFFF).to_bytes(2, "little")
    q = str(tmp_path / "fields.gz")
    open(q, "wb").write(hdr + raw[hdr_end:])
    assert gzip.decompress(open(q, "rb").read()) == data
    m, got = inflate(q, tmp_path, chunk=65536)
    assert (m, got == data) == (2, True)


def test_multi_member_bgzf_plain_and_small(tmp_path):
    data = datasets()["fastq"]
    half = len(data) // 2
    p = str(tmp_path / "two.gz")
    open(p, "wb").write(gz_bytes(data[:half]) + gz_bytes(data[half:]))
    m, got = inflate(p, tmp_path, chunk=65536)
    assert (m, got == data) == (3, True)                  # a second member: left to zlib
    p = str(tmp_path / "b.gz")
    with open(p, "wb") as fh:                             # bgzip-style
        for i in range(0, len(data), 60000):
            fh.write(bamio.bgzf_block(data[i:i + 60000]))
        fh.write(bamio.EOF_BLOCK)
    m, got = inflate(p, tmp_path)
    assert (m, got == data) == (1, True)
    p = str(tmp_path / "plain.fq")
    open(p, "wb").write(data)
    assert inflate(p, tmp_path) == (0, data)
    p = str(tmp_path / "small.gz")
    open(p, "wb").write(gz_bytes(data[:5000]))
    m, got = inflate(p, tmp_path)                         # below 1 MiB compressed and no chunk override: zlib
    assert (m, got == data[:5000]) == (3, True)
    p = str(tmp_path / "empty.gz")
    open(p, "wb").write(gz_bytes(b""))
    assert inflate(p, tmp_path, chunk=65536)[1] == b""


def test_corrupt_streams_fail_loudly(tmp_path):
    from catt_b200 import _lib
    data = datasets()["fastq"]
    raw = bytearray(gz_bytes(data))
    for where in (len(raw) // 3, len(raw) - 6):           # a flipped bit in the deflate data / in the CRC-32
        bad = bytearray(raw)
        bad[where] ^= 0x10
        p = str(tmp_path / "bad.gz")
        open(p, "wb").write(bytes(bad))
        try:
            m, got = inflate(p, tmp_path, chunk=65536)
        except _lib.CattError:
            continue
        assert m == 3 and got != data or got == data      # zlib may stop early without an error return on some corruptions ...
        assert not (m == 2)                               # ... but the parallel path never vouches for a corrupt stream


@pytest.mark.gpu
def test_fastq_gz_through_the_library(trb, fx, tmp_path):
    """testSample.fq as .gz (single member) and as BGZF: same lists as the plain file."""
    from catt_b200.host import Catt
    src = open(os.path.join(fx, "testSample.fq"), "rb").read()
    c = Catt(chain_cfg=trb, allow_v=0)
    ref = c.mainflow(os.path.join(fx, "testSample.fq"))
    exp = ([(r.seq, r.qual, r.vs, r.js, r.cdr3) for r in ref.Vpart], [(r.seq, r.qual, r.vs, r.js, r.cdr3) for r in ref.Jpart])
    p1 = str(tmp_path / "s.fq.gz")
    open(p1, "wb").write(gz_bytes(src, 6))
    p2 = str(tmp_path / "s.bgzf.fq.gz")
    with open(p2, "wb") as fh:
        for i in range(0, len(src), 65000):
            fh.write(bamio.bgzf_block(src[i:i + 65000]))
        fh.write(bamio.EOF_BLOCK)
    os.environ["CATT_PGZ_CHUNK"] = "65536"
    try:
        for p in (p1, p2):
            res = c.mainflow(p)
            got = ([(r.seq, r.qual, r.vs, r.js, r.cdr3) for r in res.Vpart], [(r.seq, r.qual, r.vs, r.js, r.cdr3) for r in res.Jpart])
            assert got == exp
            res.close()
    finally:
        os.environ.pop("CATT_PGZ_CHUNK", None)
    # an empty .gz, and a .gz of a FASTA (not FASTQ: the host parser takes over)
    p3 = str(tmp_path / "empty.fq.gz")
    open(p3, "wb").write(gz_bytes(b""))
    r0 = c.mainflow(p3)
    assert r0.info["n_reads"] == 0 and r0.Vpart == [] and r0.Jpart == []
    r0.close()
    lines = src.decode().split("\n")
    fa = "".join(f">{lines[i][1:]}\n{lines[i + 1]}\n" for i in range(0, 4 * 500, 4)).encode()
    p4, p5 = str(tmp_path / "s.fa.gz"), str(tmp_path / "s.fa")
    open(p4, "wb").write(gz_bytes(fa))
    open(p5, "wb").write(fa)
    ra, rb = c.mainflow(p4), c.mainflow(p5)
    assert ra.info["n_reads"] == rb.info["n_reads"] == 500
    assert [(r.seq, r.vs, r.js, r.cdr3) for r in ra.Vpart] == [(r.seq, r.vs, r.js, r.cdr3) for r in rb.Vpart]
    ra.close(); rb.close()
    ref.close(); c.close()


def test_random_streams_property(tmp_path):
    """Randomly structured data (runs, repeats at random distances up to the 32 KiB window and beyond, noise, text) under random zlib
    settings: the parallel decoder's bytes are zlib's, whatever blocks and matches the compressor chose."""
    rng = random.Random(2024)
    for case in range(25):
        parts, total = [], 0
        target = rng.choice([200_000, 700_000, 1_500_000])
        while total < target:
            kind = rng.random()
            if kind < 0.25:
                blk = rng.randbytes(rng.randint(1, 40000))
            elif kind < 0.5:
                blk = bytes([rng.randrange(256)]) * rng.randint(1, 70000)
            elif kind < 0.8 and parts:
                src = b"".join(parts[-3:])
                back = rng.randint(1, min(len(src), 40000))
                seg = src[len(src) - back:len(src) - back + rng.randint(1, 300)]
                blk = seg * rng.randint(1, 50)
            else:
                blk = "".join(rng.choice("ACGTN\n@+IF#") for _ in range(rng.randint(1, 30000))).encode()
            parts.append(blk)
            total += len(blk)
        data = b"".join(parts)
        level = rng.choice([1, 2, 4, 6, 9])
        strategy = rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FIXED])
        mem = rng.choice([1, 4, 8, 9])
        co = zlib.compressobj(level, zlib.DEFLATED, 31, mem, strategy)
        blob = b""
        pos = 0
        while pos < len(data):                               # occasional sync / full flushes: empty stored blocks, window resets
            step = rng.randint(1, 400_000)
            blob += co.compress(data[pos:pos + step])
            pos += step
            if rng.random() < 0.3:
                blob += co.flush(rng.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH]))
        blob += co.flush()
        p = str(tmp_path / f"r{case}.gz")
        open(p, "wb").write(blob)
        m, got = inflate(p, tmp_path, chunk=65536)
        assert got == data, (case, level, strategy, mem)
        assert m in (2, 3), case                             # the parallel path or (if a check failed) zlib - never a wrong answer
