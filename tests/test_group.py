"""One sample over several GPU ranks (catt_config.ngpu) must give bit-identical results to one GPU.

The driver's GPU box for `pytest -m gpu` has ONE GPU: the ranks of a group may share a device (catt_config.devices with
repeated ordinals), in which case they exchange through peer copies instead of NCCL - every step of the sharded Stage B
(gather of the mapped records, global order, ownership split, straddling mates, item flags, k-mer all-to-all + sub-table
gather, sizes, arena reduce) runs exactly as it does on 8 GPUs; only the transport differs.  With >= 2 visible GPUs the NCCL
transport is covered too.
"""
import os

import numpy as np
import pytest

from util import myread_tuple, synth_reads

pytestmark = pytest.mark.gpu

INFO_KEYS = ("kmer", "the_err", "avg_len", "n_reads", "v_hits", "j_hits", "v_final", "j_final", "share", "pair_seq_mismatch", "full_pushed",
             "vpart", "jpart", "direct", "left", "rescued", "cand_v", "cand_j", "cells_v", "cells_j", "gapped_v", "gapped_j")


@pytest.fixture(scope="module")
def sample(fx):
    names, seqs, quals = [], [], []
    with open(os.path.join(fx, "testSample.fq")) as fh:
        lines = fh.read().split("\n")
    for i in range(0, len(lines) - 3, 4):
        names.append(lines[i][1:].split()[0]); seqs.append(lines[i + 1]); quals.append(lines[i + 3])
    return names, seqs, quals


def lists(res):
    return [myread_tuple(r) + (r.ordinal, r.flags) for r in res.Vpart], [myread_tuple(r) + (r.ordinal, r.flags) for r in res.Jpart]


def same(a, b):
    va, ja = lists(a)
    vb, jb = lists(b)
    assert len(va) == len(vb) and len(ja) == len(jb)
    assert va == vb, "Vpart differs"
    assert ja == jb, "Jpart differs"
    for k in INFO_KEYS:
        assert a.info[k] == b.info[k], (k, a.info[k], b.info[k])


@pytest.mark.parametrize("ngpu", [2, 3, 4])
@pytest.mark.parametrize("only_cdr3", [False, True])
def test_group_equals_single_testsample(trb, sample, ngpu, only_cdr3):
    """The reference's own sample: 7,922 reads with /1 /2 mates in one file, so QNAME dedup and the both-mapped pairing cross
    rank boundaries (367 SEQ mismatches, 708 kept pairs)."""
    from catt_b200.host import Catt
    one = Catt(chain_cfg=trb, device=0, only_cdr3=only_cdr3, thread=2)
    grp = Catt(chain_cfg=trb, device=0, only_cdr3=only_cdr3, thread=2, ngpu=ngpu, devices=[0] * ngpu)
    try:
        a = one.mainflow_reads(*sample)
        b = grp.mainflow_reads(*sample)
        same(a, b)
        assert b.info["n_ranks"] == ngpu and a.info["n_ranks"] == 0
        assert b.info["pair_seq_mismatch"] == 367 and b.info["share"] > 0
        assert b.info["bytes_comm"] > 0
        b2 = grp.mainflow_reads(*sample)           # the group is reusable
        same(a, b2)
    finally:
        one.close(); grp.close()


def test_group_far_mates_straddle_ranks(trb, sample):
    """All /1 reads first, then all /2 reads: the two records of most both-mapped QNAMEs sit on different ranks, so the other
    read's packed words travel (CATT_COLL_MATES)."""
    from catt_b200.host import Catt
    names, seqs, quals = sample
    idx = sorted(range(len(names)), key=lambda i: (names[i].endswith("/2"), i))
    re = ([names[i] for i in idx], [seqs[i] for i in idx], [quals[i] for i in idx])
    one = Catt(chain_cfg=trb, device=0)
    grp = Catt(chain_cfg=trb, device=0, ngpu=2, devices=[0, 0])
    try:
        a = one.mainflow_reads(*re)
        b = grp.mainflow_reads(*re)
        same(a, b)
        assert b.info["bytes_coll"][1] > 0, "no mate reads travelled: the case is not exercised"
    finally:
        one.close(); grp.close()


@pytest.mark.parametrize("length,n", [(150, 30000), (50, 40000)])
def test_group_equals_single_synthetic(trb, length, n):
    """Synthetic reads (k-mer table + extension busy; 50 bp: k = 15, most reads partial), 4 ranks, reads with N."""
    from catt_b200.host import Catt
    from oracle import oracle as O
    V, J = O.RefSet(trb["vregion"]), O.RefSet(trb["jregion"])
    vrecs = [V.record(i) for i in range(V.nrec)]
    jrecs = [J.record(i) for i in range(J.nrec)]
    reads = synth_reads(3000, vrecs, jrecs, length, with_n=True)
    one = Catt(chain_cfg=trb, device=0, only_cdr3=True)
    grp = Catt(chain_cfg=trb, device=0, only_cdr3=True, ngpu=4, devices=[0, 0, 0, 0])
    try:
        same(one.mainflow_reads(*reads), grp.mainflow_reads(*reads))
        bufs = one.synth_reads(n, length, first_index=5)
        a = one.mainflow_buffers(n, *bufs)
        b = grp.mainflow_buffers(n, *bufs)
        same(a, b)
        assert b.info["left"] > 0 and b.info["bytes_coll"][3] > 0 and b.info["bytes_coll"][4] > 0
    finally:
        one.close(); grp.close()


def test_group_paired_files_and_tiny_inputs(trb, sample):
    """Two input files (per-file order / dedup / pairing, one k-mer table), blocks that straddle the file boundary, and inputs
    smaller than the group (empty blocks)."""
    from catt_b200.host import Catt
    names, seqs, quals = sample
    f1 = (names[:3000], seqs[:3000], quals[:3000])
    f2 = (names[3000:4500], seqs[3000:4500], quals[3000:4500])
    one = Catt(chain_cfg=trb, device=0)
    grp = Catt(chain_cfg=trb, device=0, ngpu=3, devices=[0, 0, 0])
    try:
        same(one.mainflow_reads_files([f1, f2]), grp.mainflow_reads_files([f1, f2]))
        for k in (0, 1, 2, 5):
            same(one.mainflow_reads(names[:k], seqs[:k], quals[:k]), grp.mainflow_reads(names[:k], seqs[:k], quals[:k]))
    finally:
        one.close(); grp.close()


def test_group_file_input(trb, fx, tmp_path):
    """catt_run_file on a group: every rank ingests its byte range of the FASTQ on the device (record-aligned cuts, global read
    ordinals exchanged before Stage A); other formats take the host parser.  Same lists either way."""
    import gzip
    from catt_b200.host import Catt
    one = Catt(chain_cfg=trb, device=0)
    fq = os.path.join(fx, "testSample.fq")
    txt = open(fq, "rb").read()
    variants = {"plain": txt, "no_final_newline": txt.rstrip(b"\n"), "trailing_blank_lines": txt + b"\n\n \n", "crlf": txt.replace(b"\n", b"\r\n")}
    try:
        want = one.mainflow(fq)
        for ngpu in (2, 3, 5):
            grp = Catt(chain_cfg=trb, device=0, ngpu=ngpu, devices=[0] * ngpu)
            try:
                for name, data in variants.items():
                    p = tmp_path / f"{name}.fq"
                    p.write_bytes(data)
                    same(want, grp.mainflow(str(p)))
                gz = tmp_path / "s.fq.gz"
                with gzip.open(gz, "wb") as fh:
                    fh.write(txt)
                same(want, grp.mainflow(str(gz)))                      # host path (inflate + parse), then split
                tiny = tmp_path / "tiny.fq"
                tiny.write_bytes(b"\n".join(txt.split(b"\n")[:8]) + b"\n")   # 2 reads on up to 5 ranks: empty byte ranges
                same(one.mainflow(str(tiny)), grp.mainflow(str(tiny)))
            finally:
                grp.close()
    finally:
        one.close()


@pytest.mark.parametrize("piece", [4096, 70000, 1 << 20])
def test_group_file_input_many_pieces(trb, fx, tmp_path, monkeypatch, piece):
    """The ranks of a group take their byte range in record-aligned pieces: seeding, SW and the selection up to bwa's tie-break run
    on piece k while piece k+1 is read; the tie-break (which hashes the GLOBAL read ordinal) and the TRACK pass follow when every
    rank has counted its reads.  Small pieces (CATT_GROUP_PIECE) put dozens of them through a 2.7 MB file: same lists as one GPU,
    also for CRLF text and a file whose records carry long '+' lines."""
    from catt_b200.host import Catt
    fq = os.path.join(fx, "testSample.fq")
    txt = open(fq, "rb").read()
    lines = txt.split(b"\n")
    plus = b"\n".join(b"+" + lines[i - 2][1:] if i % 4 == 2 and i + 1 < len(lines) else l for i, l in enumerate(lines))
    variants = {"plain": txt, "crlf": txt.replace(b"\n", b"\r\n"), "plus_names": plus, "no_final_newline": txt.rstrip(b"\n")}
    one = Catt(chain_cfg=trb, device=0)
    try:
        want = one.mainflow(fq)
        monkeypatch.setenv("CATT_GROUP_PIECE", str(piece))
        for ngpu in (2, 3):
            grp = Catt(chain_cfg=trb, device=0, ngpu=ngpu, devices=[0] * ngpu)
            try:
                for name, data in variants.items():
                    p = tmp_path / f"{name}_{ngpu}.fq"
                    p.write_bytes(data)
                    same(want, grp.mainflow(str(p)))
            finally:
                grp.close()
    finally:
        one.close()


def test_group_records_fixed_kmer_and_unmapped_samples(trb, sample):
    """keep_records on a group (every rank's records, in input order, on the root), a fixed args["kmer"], and samples in which
    nothing maps (random reads, reads of N): the exchanges must cope with empty gathers."""
    from catt_b200.host import Catt
    from util import random_reads
    names, seqs, quals = [x[:3000] for x in sample]
    one = Catt(chain_cfg=trb, device=0, keep_records=True, kmer=19)
    grp = Catt(chain_cfg=trb, device=0, keep_records=True, kmer=19, ngpu=3, devices=[0, 0, 0])
    try:
        a, b = one.mainflow_reads(names, seqs, quals), grp.mainflow_reads(names, seqs, quals)
        same(a, b)
        assert a.info["kmer"] == 19
        for which in (0, 1):
            ra, rb = a.records(which), b.records(which)
            assert len(ra) == len(rb) == 3000 and ra.tobytes() == rb.tobytes()
        rnd = random_reads(4000, 60, 11)
        rn = [f"x{i}" for i in range(len(rnd))]
        rq = ["I" * 60] * len(rnd)
        a, b = one.mainflow_reads(rn, rnd, rq), grp.mainflow_reads(rn, rnd, rq)
        same(a, b)
        alln = ["N" * 80] * 500
        a, b = one.mainflow_reads(rn[:500], alln, ["#" * 80] * 500), grp.mainflow_reads(rn[:500], alln, ["#" * 80] * 500)
        same(a, b)
        assert b.info["v_hits"] == 0 and b.info["vpart"] == 0
    finally:
        one.close(); grp.close()


def test_group_nccl_when_two_gpus(trb, sample):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (NCCL transport)")
    from catt_b200.host import Catt
    one = Catt(chain_cfg=trb, device=0, only_cdr3=True)
    grp = Catt(chain_cfg=trb, device=0, only_cdr3=True, ngpu=2)
    try:
        same(one.mainflow_reads(*sample), grp.mainflow_reads(*sample))
        bufs = one.synth_reads(200000, 150)
        same(one.mainflow_buffers(200000, *bufs), grp.mainflow_buffers(200000, *bufs))
    finally:
        one.close(); grp.close()


def test_kmer_out_of_range_is_rejected(trb):
    """args["kmer"] is free in the reference's CLI (catt:18); the table holds one 64-bit word per k-mer."""
    from catt_b200 import _lib
    from catt_b200.host import Catt
    with pytest.raises(_lib.CattError):
        Catt(chain_cfg=trb, device=0, kmer=32)
    with pytest.raises(_lib.CattError):
        Catt(chain_cfg=trb, device=0, kmer=0)
    Catt(chain_cfg=trb, device=0, kmer=31).close()
