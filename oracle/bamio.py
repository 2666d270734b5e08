"""TEST-ONLY: a BAM / SAM writer and a pure-Python restatement of what `samtools fastq -N` extracts from such a file
(catt.jl:760, the `--bam` branch).  samtools is third-party and absent here (README.md:141 names "samtools >= 1.8"); the
rules below are its documented behaviour and the SAM/BAM specification - parity unpinned, there is no golden:

  * default filter -F 0x900: SECONDARY and SUPPLEMENTARY records are skipped, everything else is written in file order;
  * -N: "/1" after READ1-only names, "/2" after READ2-only names;
  * REVERSE (0x10) records are reverse-complemented (IUPAC-aware, 4-bit alphabet "=ACMGRSVTWYHKDBN"), quality reversed;
  * missing quality -> 'B' per base.

Never imported by the product; tests use it to make inputs for catt_run_bam / catt_bam_to_fastq and to say what they should read.
"""
from __future__ import annotations

import gzip
import struct
import zlib

NT16 = "=ACMGRSVTWYHKDBN"
CODE = {c: i for i, c in enumerate(NT16)}
CODE.update({c.lower(): i for i, c in enumerate(NT16)})
COMP = {"=": "=", "A": "T", "C": "G", "M": "K", "G": "C", "R": "Y", "S": "S", "V": "B", "T": "A", "W": "W", "Y": "R", "H": "D",
        "K": "M", "D": "H", "B": "V", "N": "N"}
EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def fold(seq: str) -> str:
    """What a base string looks like after a trip through the 4-bit BAM alphabet (unknown characters become N)."""
    return "".join(NT16[CODE.get(c, 15)] for c in seq)


def bam_record(name: str, flag: int, seq: str, qual, cigar=(), ref_id=-1, pos=-1, extra: bytes = b"") -> bytes:
    """seq / qual as STORED (reference orientation for reverse-strand records); qual = list of phred ints, or None = missing."""
    l_seq = len(seq)
    packed = bytearray((l_seq + 1) // 2)
    for i, c in enumerate(seq):
        packed[i >> 1] |= CODE.get(c, 15) << (0 if i & 1 else 4)
    q = bytes([0xFF] * l_seq) if qual is None else bytes(qual)
    cig = b"".join(struct.pack("<I", (ln << 4) | "MIDNSHP=X".index(op)) for ln, op in cigar)
    body = struct.pack("<iiBBHHHiiii", ref_id, pos, len(name) + 1, 0, 4680, len(cigar), flag, l_seq, -1, -1, 0)
    body += name.encode() + b"\0" + cig + bytes(packed) + q + extra
    return struct.pack("<I", len(body)) + body


def bgzf_block(data: bytes, level: int = 6) -> bytes:
    assert len(data) <= 65280
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    cdata = co.compress(data) + co.flush()
    bsize = len(cdata) + 25
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + cdata +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def write_bam(path: str, records, block_bytes: int = 65280, level: int = 6, refs=(("chr1", 1000000),), compressed: bool = True):
    """records: iterable of dicts(name, flag, seq, qual[, cigar, extra]).  Blocks are cut every `block_bytes` of payload regardless of
    record boundaries, so small values make records straddle BGZF blocks."""
    text = b"@HD\tVN:1.6\tSO:unsorted\n" + b"".join(f"@SQ\tSN:{n}\tLN:{l}\n".encode() for n, l in refs)
    payload = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for n, l in refs:
        payload += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    payload += b"".join(bam_record(r["name"], r["flag"], r["seq"], r["qual"], r.get("cigar", ()), extra=r.get("extra", b"")) for r in records)
    with open(path, "wb") as fh:
        if not compressed:
            fh.write(payload)
            return
        for i in range(0, len(payload), block_bytes):
            fh.write(bgzf_block(payload[i:i + block_bytes], level))
        fh.write(EOF_BLOCK)


def write_sam(path: str, records, gz: bool = False, crlf: bool = False):
    nl = "\r\n" if crlf else "\n"
    lines = ["@HD\tVN:1.6\tSO:unsorted", "@SQ\tSN:chr1\tLN:1000000"]
    for r in records:
        q = "*" if r["qual"] is None else "".join(chr(33 + x) for x in r["qual"])
        f = [r["name"], str(r["flag"]), "*", "0", "0", "*", "*", "0", "0", r["seq"] or "*", q if r["seq"] else "*"]
        if r.get("tags"):
            f += r["tags"]
        lines.append("\t".join(f))
    data = (nl.join(lines) + nl).encode()
    with (gzip.open(path, "wb") if gz else open(path, "wb")) as fh:
        fh.write(data)


def read_bam(path: str):
    """Independent reader (gzip module + struct): the stored records back as dicts, to check the writer."""
    raw = open(path, "rb").read()
    d = gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw
    assert d[:4] == b"BAM\1"
    l_text, = struct.unpack_from("<i", d, 4)
    p = 8 + l_text
    n_ref, = struct.unpack_from("<i", d, p)
    p += 4
    for _ in range(n_ref):
        ln, = struct.unpack_from("<i", d, p)
        p += 4 + ln + 4
    out = []
    while p < len(d):
        bs, = struct.unpack_from("<I", d, p)
        ref_id, pos, l_rn, mapq, bin_, n_cig, flag, l_seq, _, _, _ = struct.unpack_from("<iiBBHHHiiii", d, p + 4)
        q = p + 4 + 32
        name = d[q:q + l_rn - 1].decode()
        q += l_rn + 4 * n_cig
        seq = "".join(NT16[(d[q + (i >> 1)] >> (0 if i & 1 else 4)) & 15] for i in range(l_seq))
        q += (l_seq + 1) // 2
        qual = None if l_seq and d[q] == 0xFF else list(d[q:q + l_seq])
        out.append({"name": name, "flag": flag, "seq": seq, "qual": qual})
        p += 4 + bs
    return out


def samtools_fastq_N(records):
    """[(name, seq, qual)] as `samtools fastq -N` would write them for the stored records."""
    out = []
    for r in records:
        flag = r["flag"]
        if flag & 0x900:
            continue
        r1, r2 = bool(flag & 0x40), bool(flag & 0x80)
        name = r["name"] + ("/1" if r1 and not r2 else "/2" if r2 and not r1 else "")
        seq = fold(r["seq"])
        qual = "B" * len(seq) if r["qual"] is None else "".join(chr(33 + x) for x in r["qual"])
        if flag & 0x10:
            seq = "".join(COMP[c] for c in reversed(seq))
            qual = qual[::-1]
        out.append((name, seq, qual))
    return out


def fastq_text(reads) -> bytes:
    return "".join(f"@{n}\n{s}\n+\n{q}\n" for n, s, q in reads).encode()
