"""Multi-GPU decomposition (catt_b200/shard.py).  CPU: world_size-2 gloo run of the sharding logic with the oracle as the
stand-in aligner (block bounds, global first_ordinal, padded all-gather).  GPU: the same class over the real library."""
import os
import socket

import numpy as np
import pytest

from util import synth_reads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_balance():
    from catt_b200.shard import shard_bounds
    for n in (0, 1, 7, 8, 1000, 1001, 12345):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fx, n_reads, outdir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from catt_b200 import _lib
    from catt_b200.refg import chain_config
    from catt_b200.shard import ShardedCatt
    from oracle import oracle as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    cc = chain_config("hs", "TRB", "CDR3", os.path.join(fx, "resource"))
    V, J = O.RefSet(cc["vregion"]), O.RefSet(cc["jregion"])
    vrecs = [V.record(i) for i in range(V.nrec)]
    jrecs = [J.record(i) for i in range(J.nrec)]
    names, seqs, quals = synth_reads(n_reads, vrecs, jrecs, 100, seed=4242)
    nb, no = _lib.concat(names)
    sb, so = _lib.concat(seqs)
    qb, _ = _lib.concat(quals)
    cols = {k: i for i, k in enumerate(O.REC_FIELDS)}
    calls = []

    def to_aln(mat, first):
        out = np.zeros(mat.shape[0], dtype=_lib.ALN_DTYPE)
        out["ordinal"] = np.arange(first, first + mat.shape[0])
        for gk, ok in (("rid", "rid"), ("strand", "strand"), ("mapped", "mapped"), ("as", "AS"), ("qb", "qb"), ("m_end", "m_end"),
                       ("rb", "rb"), ("qe", "qe"), ("re", "re"), ("nm", "nm"), ("mi", "mi"), ("ncand", "ncand"), ("ntied", "ntied")):
            out[gk] = mat[:, cols[ok]]
        return out

    def align_fn(n, seq_buf, seq_off, first_ordinal):               # stand-in for catt_align_shard: the CPU oracle
        sub = [bytes(seq_buf[seq_off[i]:seq_off[i + 1]]).decode() for i in range(n)]
        calls.append((n, first_ordinal))
        return to_aln(V.align(sub, first_ordinal=first_ordinal), first_ordinal), to_aln(J.align(sub, first_ordinal=first_ordinal), first_ordinal)

    def assemble_fn(n, name_buf, name_off, seq_buf, seq_off, qual_buf, v, j):
        return v, j

    sc = ShardedCatt(align_fn=align_fn, assemble_fn=assemble_fn)
    res = sc.mainflow_buffers(n_reads, nb, no, sb, so, qb)
    from catt_b200.shard import shard_bounds
    lo, hi = shard_bounds(n_reads, world, rank)
    assert calls == [(hi - lo, lo)]
    if rank == 0:
        v, j = res
        whole_v, whole_j = to_aln(V.align(seqs), 0), to_aln(J.align(seqs), 0)
        np.save(os.path.join(outdir, "ok.npy"), np.array([int(np.array_equal(v, whole_v)), int(np.array_equal(j, whole_j)), int(v["mapped"].sum())]))
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_alignment_equals_whole_gloo(fx, tmp_path):
    """Two ranks, gloo: per-block alignment with the global first_ordinal + padded all-gather == aligning the whole sample."""
    import torch.multiprocessing as mp
    port = _free_port()
    n = 301                                                          # odd: the blocks differ in size, the gather pads
    mp.spawn(_worker, args=(2, port, fx, n, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(os.path.join(str(tmp_path), "ok.npy"))
    assert ok[0] == 1 and ok[1] == 1 and ok[2] > 100


@pytest.mark.gpu
def test_sharded_class_on_gpu_world1(fx, trb):
    """The real library behind ShardedCatt (world size 1, gloo group): identical to the unsharded call."""
    import torch.distributed as dist
    from catt_b200 import _lib
    from catt_b200.host import Catt
    from catt_b200.shard import ShardedCatt
    from oracle import oracle as O
    from util import myread_tuple
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1)
    try:
        V, J = O.RefSet(trb["vregion"]), O.RefSet(trb["jregion"])
        names, seqs, quals = synth_reads(4000, [V.record(i) for i in range(V.nrec)], [J.record(i) for i in range(J.nrec)], 150, seed=5)
        nb, no = _lib.concat(names)
        sb, so = _lib.concat(seqs)
        qb, _ = _lib.concat(quals)
        c = Catt(chain_cfg=trb)
        whole = c.mainflow_buffers(len(seqs), nb, no, sb, so, qb)
        res = ShardedCatt(c).mainflow_buffers(len(seqs), nb, no, sb, so, qb)
        assert [myread_tuple(r) for r in res.Vpart] == [myread_tuple(r) for r in whole.Vpart]
        assert [myread_tuple(r) for r in res.Jpart] == [myread_tuple(r) for r in whole.Jpart]
        c.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_device_gather_world1(fx, trb):
    """on_device=True: records stay in HBM (catt_records_device + NCCL all-gather + catt_assemble_device); NCCL group of one rank.
    The N-rank run of the same code is profiles/shard_check.py under torchrun (profiles/r01e_shard_2gpu.txt)."""
    import torch
    import torch.distributed as dist
    from catt_b200 import _lib
    from catt_b200.host import Catt
    from catt_b200.shard import ShardedCatt
    from oracle import oracle as O
    from util import myread_tuple
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        V, J = O.RefSet(trb["vregion"]), O.RefSet(trb["jregion"])
        names, seqs, quals = synth_reads(5000, [V.record(i) for i in range(V.nrec)], [J.record(i) for i in range(J.nrec)], 150, seed=6)
        nb, no = _lib.concat(names)
        sb, so = _lib.concat(seqs)
        qb, _ = _lib.concat(quals)
        c = Catt(chain_cfg=trb)
        whole = c.mainflow_buffers(len(seqs), nb, no, sb, so, qb)
        want = ([myread_tuple(r) for r in whole.Vpart], [myread_tuple(r) for r in whole.Jpart])
        res = ShardedCatt(c, device=torch.device("cuda", 0), on_device=True).mainflow_buffers(len(seqs), nb, no, sb, so, qb)
        assert ([myread_tuple(r) for r in res.Vpart], [myread_tuple(r) for r in res.Jpart]) == want
        ref = O.run_mem(V, J, O.make_config(trb), names, seqs, quals)
        assert want == ([myread_tuple(r) for r in ref.Vpart], [myread_tuple(r) for r in ref.Jpart])
        c.close()
    finally:
        dist.destroy_process_group()
