"""`--bam` input (catt.jl:756-764): catt_bam_to_fastq / catt_run_bam against the test-side restatement of `samtools fastq -N`
(oracle/bamio.py).  No samtools in the image, so these are consistency tests of two independent implementations of the documented
rules - parity unpinned."""
import os
import random

import pytest

from oracle import bamio


def make_records(seed, n, sample=None):
    """Stored BAM records with every flag combination that matters; sequences from the reference's sample reads when given."""
    rng = random.Random(seed)
    recs = []
    for i in range(n):
        if sample is not None:
            seq, qual = sample[1][i % len(sample[1])], [ord(c) - 33 for c in sample[2][i % len(sample[2])]]
            name = sample[0][i % len(sample[0])]
        else:
            ln = rng.choice([0, 1, 2, 17, 50, 101, 150])
            seq = "".join(rng.choice("ACGTNacgtnRYKM=") for _ in range(ln))
            qual = [rng.randint(0, 60) for _ in range(ln)]
            name = f"read{i}"
        flag = 0
        r = rng.random()
        if r < 0.30:
            flag |= 0x10
        if rng.random() < 0.5:
            flag |= 0x1 | rng.choice([0x40, 0x80, 0xC0, 0x0])
        if rng.random() < 0.08:
            flag |= 0x100
        if rng.random() < 0.05:
            flag |= 0x800
        if rng.random() < 0.2:
            flag |= 0x4
        if rng.random() < 0.1:
            qual = None
        if flag & 0x10:                              # aligners store reverse-strand reads reverse-complemented
            seq = "".join(bamio.COMP[c] for c in reversed(bamio.fold(seq)))
            qual = qual[::-1] if qual is not None else None
        recs.append({"name": name, "flag": flag, "seq": seq, "qual": qual, "cigar": [(len(seq), "M")] if seq and not flag & 4 else [],
                     "extra": b"NMC\x00" if rng.random() < 0.5 else b""})
    return recs


def convert(path, tmp_path):
    from catt_b200.host import bam_to_fastq
    out = str(tmp_path / "out.fq")
    bam_to_fastq(path, out)
    return open(out, "rb").read()


def test_writer_and_independent_reader_agree(tmp_path):
    recs = make_records(1, 300)
    p = str(tmp_path / "a.bam")
    bamio.write_bam(p, recs, block_bytes=777)
    back = bamio.read_bam(p)
    assert [(r["name"], r["flag"], bamio.fold(r["seq"]), r["qual"] if r["seq"] else None) for r in recs] == \
           [(r["name"], r["flag"], r["seq"], r["qual"] if r["seq"] else None) for r in back]


@pytest.mark.parametrize("block_bytes,level", [(65280, 6), (4096, 1), (333, 9), (65280, 0)])
def test_bam_to_fastq_matches_the_restatement(tmp_path, block_bytes, level):
    recs = make_records(block_bytes, 2000)
    p = str(tmp_path / "a.bam")
    bamio.write_bam(p, recs, block_bytes=block_bytes, level=level)
    exp = bamio.samtools_fastq_N(recs)
    assert 0 < len(exp) < len(recs)                      # secondary / supplementary records were dropped
    assert any(n.endswith("/1") for n, _, _ in exp) and any(n.endswith("/2") for n, _, _ in exp)
    assert convert(p, tmp_path) == bamio.fastq_text(exp)


def test_uncompressed_bam_sam_and_gzipped_sam(tmp_path):
    recs = make_records(7, 500)
    exp = bamio.fastq_text(bamio.samtools_fastq_N(recs))
    p = str(tmp_path / "raw.bam")
    bamio.write_bam(p, recs, compressed=False)
    assert convert(p, tmp_path) == exp
    sam_recs = [dict(r, tags=["NM:i:0", "XX:Z:a b"]) if i % 3 == 0 else r for i, r in enumerate(recs)]
    for gz, crlf in ((False, False), (True, False), (False, True)):
        p = str(tmp_path / f"a{int(gz)}{int(crlf)}.sam")
        bamio.write_sam(p, sam_recs, gz=gz, crlf=crlf)
        assert convert(p, tmp_path) == exp


def test_empty_inputs(tmp_path):
    p = str(tmp_path / "empty.bam")
    bamio.write_bam(p, [])
    assert convert(p, tmp_path) == b""
    p = str(tmp_path / "only_secondary.bam")
    bamio.write_bam(p, [{"name": "x", "flag": 0x100, "seq": "ACGT", "qual": [30] * 4}])
    assert convert(p, tmp_path) == b""
    p = str(tmp_path / "header_only.sam")
    open(p, "w").write("@HD\tVN:1.6\n")
    assert convert(p, tmp_path) == b""


def test_malformed_inputs_fail_loudly(tmp_path):
    from catt_b200 import _lib
    from catt_b200.host import bam_to_fastq
    recs = make_records(3, 200)
    p = str(tmp_path / "a.bam")
    bamio.write_bam(p, recs, block_bytes=2048)
    raw = open(p, "rb").read()
    out = str(tmp_path / "o.fq")
    cases = {"truncated": raw[:len(raw) // 2], "flipped": raw[:300] + bytes([raw[300] ^ 0x55]) + raw[301:]}
    for name, data in cases.items():
        q = str(tmp_path / f"{name}.bam")
        open(q, "wb").write(data)
        with pytest.raises(_lib.CattError):
            bam_to_fastq(q, out)
    q = str(tmp_path / "bad.sam")
    open(q, "w").write("r1\t0\t*\t0\t0\t*\t*\t0\t0\tACGT\tIII\n")          # SEQ / QUAL lengths differ
    with pytest.raises(_lib.CattError):
        bam_to_fastq(q, out)
    open(q, "w").write("r1\t0\t*\t0\t0\n")
    with pytest.raises(_lib.CattError):
        bam_to_fastq(q, out)
    with pytest.raises(_lib.CattError):
        bam_to_fastq(str(tmp_path / "missing.bam"), out)


# ---------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def sample(fx):
    names, seqs, quals = [], [], []
    with open(os.path.join(fx, "testSample.fq")) as fh:
        lines = fh.read().split("\n")
    for i in range(0, len(lines) - 3, 4):
        names.append(lines[i][1:].split()[0]); seqs.append(lines[i + 1]); quals.append(lines[i + 3])
    return names, seqs, quals


def read_tuple(r):
    return (r.seq, r.qual, r.vs, r.js, r.cdr3, bool(r.able))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["bam", "sam"])
def test_run_bam_equals_the_fastq_path(trb, fx, sample, tmp_path, kind):
    """catt_run_bam on a BAM / SAM of the reference's sample reads (random strands, mates, secondary records) == catt_run_reads on
    the reads `samtools fastq -N` would have written == the oracle on the same reads."""
    from catt_b200.host import Catt
    from oracle import oracle as O
    recs = make_records(11, len(sample[0]), sample)
    p = str(tmp_path / f"s.{kind}")
    if kind == "bam":
        bamio.write_bam(p, recs, block_bytes=30000)
    else:
        bamio.write_sam(p, recs)
    reads = bamio.samtools_fastq_N(recs)
    names, seqs, quals = [r[0] for r in reads], [r[1] for r in reads], [r[2] for r in reads]
    c = Catt(chain_cfg=trb, allow_v=0)
    a = c.mainflow_bam(p)
    b = c.mainflow_reads(names, seqs, quals)
    V, J = O.RefSet(trb["vregion"]), O.RefSet(trb["jregion"])
    ref = O.run_mem(V, J, O.make_config(trb, allow_v=0), names, seqs, quals)
    assert [read_tuple(r) for r in a.Vpart] == [read_tuple(r) for r in b.Vpart] == [read_tuple(r) for r in ref.Vpart]
    assert [read_tuple(r) for r in a.Jpart] == [read_tuple(r) for r in b.Jpart] == [read_tuple(r) for r in ref.Jpart]
    assert len(a.Vpart) > 500
    for k in ("kmer", "v_hits", "j_hits", "vpart", "jpart", "direct", "left", "rescued", "the_err"):
        assert a.info[k] == b.info[k] == ref.info[k], k
    a.close(); b.close(); c.close()


def test_mutated_inputs_never_crash(tmp_path):
    """Random byte flips, truncations and insertions in BAM / raw BAM / SAM / .gz inputs: either a clean result or a CattError."""
    import zlib
    from catt_b200 import _lib
    from catt_b200.host import bam_to_fastq, inflate_file
    rng = random.Random(99)
    recs = make_records(5, 300)
    srcs = {}
    for kind, kw in (("bam", {"block_bytes": 3000}), ("rawbam", {"compressed": False})):
        p = str(tmp_path / f"{kind}.bin")
        bamio.write_bam(p, recs, **kw)
        srcs[kind] = open(p, "rb").read()
    p = str(tmp_path / "s.sam")
    bamio.write_sam(p, recs)
    srcs["sam"] = open(p, "rb").read()
    text = bamio.fastq_text(bamio.samtools_fastq_N(recs)) * 8
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    srcs["gz"] = co.compress(text) + co.flush()
    os.environ["CATT_PGZ_CHUNK"] = "65536"
    outcomes = {"ok": 0, "err": 0}
    try:
        for it in range(240):
            kind = ("bam", "rawbam", "sam", "gz")[it % 4]
            data = bytearray(srcs[kind])
            for _ in range(rng.choice([1, 1, 2, 5, 20])):
                op = rng.random()
                if op < 0.6:
                    data[rng.randrange(len(data))] = rng.randrange(256)
                elif op < 0.8:
                    del data[rng.randrange(1, len(data)):]
                else:
                    q = rng.randrange(len(data) + 1)
                    data[q:q] = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 9)))
            f = str(tmp_path / "x.bin")
            open(f, "wb").write(bytes(data))
            try:
                if kind == "gz":
                    inflate_file(f, str(tmp_path / "o"))
                else:
                    bam_to_fastq(f, str(tmp_path / "o"))
                outcomes["ok"] += 1
            except _lib.CattError:
                outcomes["err"] += 1
    finally:
        os.environ.pop("CATT_PGZ_CHUNK", None)
    assert outcomes["err"] > 50 and outcomes["ok"] + outcomes["err"] == 240
