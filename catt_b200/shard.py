"""One sample over N GPUs of one node: reads shard naturally, so every rank runs Stage A
(`catt_align_shard`: seeding + SW + hit selection, > 70 % of the device time) on a contiguous block of the input with
the GLOBAL `first_ordinal` (bwa's hash tie-break uses the global read ordinal), the 32-byte alignment records are
all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests), and Stage B (`catt_assemble`: the SAM order / QNAME
dedup / V-J join are global, the k-mer table is a global last-writer-wins map) runs once, on rank `root`.

With `on_device=True` neither the records nor the reads leave HBM: the ranks all-gather the records between the
contexts' record arrays (`catt_records_device`) and the 2-bit packed reads between the contexts' read buffers
(`catt_reads_device`), NCCL over NVLink, and the root runs `catt_assemble_resident` - it re-reads only the QNAMEs (and,
in full-list mode, the quality strings) from the host.

One process per GPU, launched with torchrun; `torch.distributed` is plumbing only - every computation is in
libcatt_b200.so.  Independent samples (the weak-scaling mode of bench.py) need no exchange at all.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`: sizes differ by at most one read, earlier ranks take the extra ones."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_records(local: np.ndarray, n: int, group=None, device=None) -> np.ndarray:
    """All-gather the per-read alignment records (ALN_DTYPE) of every rank's block into the full array of n records.

    Blocks are padded to the largest block so one all_gather_into_tensor moves everything; `device` is the CUDA device
    to stage through for the NCCL backend (None = CPU tensors, gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rec = _lib.ALN_DTYPE.itemsize
    cap = shard_bounds(n, world, 0)[1]                      # rank 0 owns the largest block
    buf = np.zeros(max(cap, 1) * rec, dtype=np.uint8)
    raw = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    buf[:raw.size] = raw
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device, non_blocking=False)
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    flat = out.cpu().numpy().reshape(world, -1)
    full = np.zeros(n, dtype=_lib.ALN_DTYPE)
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        full[lo:hi] = flat[r, :(hi - lo) * rec].view(_lib.ALN_DTYPE)
    return full


class _DevMem:
    """A device allocation owned by libcatt_b200, exposed to torch through the CUDA array interface (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def gather_records_device(ctx, n_local: int, n: int, group=None, device=None):
    """NCCL all-gather of the alignment records WITHOUT leaving HBM: every rank's block sits at the start of its context's
    record arrays (catt_align_shard with NULL outputs); the gathered, padded blocks are copied device-to-device into the
    arrays of this rank's context at their global positions (NVLink all-gather + one D2D pass, no PCIe)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rec = _lib.ALN_DTYPE.itemsize
    cap = max(shard_bounds(n, world, 0)[1], 1)
    for which in (0, 1):
        ptr = ctx.records_device(which, max(n, cap))
        whole = torch.as_tensor(_DevMem(ptr, max(n, cap) * rec), device=device)
        send = torch.zeros(cap * rec, dtype=torch.uint8, device=device)
        send[:n_local * rec].copy_(whole[:n_local * rec])
        out = torch.empty(world * cap * rec, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(out, send, group=group)
        for r in range(world):
            lo, hi = shard_bounds(n, world, r)
            whole[lo * rec:hi * rec].copy_(out[r * cap * rec:r * cap * rec + (hi - lo) * rec])
    torch.cuda.synchronize(device)


def gather_reads_device(ctx, n_local: int, n: int, group=None, device=None) -> int:
    """NCCL all-gather of the 2-bit PACKED reads Stage A left in every context (words + lengths; ~4x smaller than the
    ASCII bases), so the root's Stage B finds all n reads resident without re-reading, re-uploading or re-packing them
    from the host.  Word offsets are rebuilt from the gathered lengths (same layout rule as the packer).  Returns the
    longest read of the sample."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nl, wl, ml = ctx.shard_counts()
    if nl != n_local:
        raise RuntimeError(f"context holds {nl} packed reads, expected this rank's block of {n_local}")
    meta = torch.tensor([wl, ml], dtype=torch.int64, device=device)
    metas = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(metas, meta, group=group)
    metas = metas.cpu().view(world, 2)
    words = [int(x) for x in metas[:, 0]]
    max_len = int(metas[:, 1].max())
    total, cap_w, cap_n = sum(words), max(max(words), 1), max(shard_bounds(n, world, 0)[1], 1)
    wp, op, lp = ctx.reads_device(max(n, cap_n), max(total, cap_w))
    W = torch.as_tensor(_DevMem(wp, max(total, cap_w) * 4), device=device).view(torch.int32)
    L = torch.as_tensor(_DevMem(lp, max(n, cap_n) * 4), device=device).view(torch.int32)
    O = torch.as_tensor(_DevMem(op, (n + 1) * 4), device=device).view(torch.int32)
    send_w = torch.zeros(cap_w, dtype=torch.int32, device=device)
    send_w[:wl].copy_(W[:wl])
    send_l = torch.zeros(cap_n, dtype=torch.int32, device=device)
    send_l[:n_local].copy_(L[:n_local])
    out_w = torch.empty(world * cap_w, dtype=torch.int32, device=device)
    out_l = torch.empty(world * cap_n, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(out_w, send_w, group=group)
    dist.all_gather_into_tensor(out_l, send_l, group=group)
    wbase = 0
    for r in range(world):
        lo, hi = shard_bounds(n, world, r)
        W[wbase:wbase + words[r]].copy_(out_w[r * cap_w:r * cap_w + words[r]])
        L[lo:hi].copy_(out_l[r * cap_n:r * cap_n + (hi - lo)])
        wbase += words[r]
    ln = L[:n].to(torch.int64)
    off = torch.zeros(n + 1, dtype=torch.int64, device=device)
    torch.cumsum(((ln + 15) >> 4) + ((ln + 31) >> 5), 0, out=off[1:])
    if int(off[n]) != total:
        raise RuntimeError("gathered read lengths do not add up to the gathered packed words")
    O.copy_(off.view(torch.int32)[::2])          # low 32 bits (offsets are uint32 words)
    torch.cuda.synchronize(device)
    return max_len


class ShardedCatt:
    """`Catt.mainflow_buffers` for one sample spread over the ranks of a process group.

    `align_fn(n, seq_view, seq_off_view, first_ordinal) -> (v_recs, j_recs)` and
    `assemble_fn(n, names, name_off, seqs, seq_off, quals, v_recs, j_recs) -> result` default to the context's
    `catt_align_shard` / `catt_assemble`; the CPU tests inject stand-ins so the sharding logic runs under gloo."""

    def __init__(self, ctx=None, group=None, device=None, align_fn=None, assemble_fn=None, root=0, on_device=False):
        self.ctx, self.group, self.device, self.root, self.on_device = ctx, group, device, root, on_device
        self.laps_ms = {}                         # host-clock laps of the last on_device call (this rank)
        self.align_fn = align_fn or self._align
        self.assemble_fn = assemble_fn or self._assemble

    def _align(self, n, seq_buf, seq_off, first_ordinal):
        return self.ctx.align_shard_buffers(n, seq_buf, seq_off, first_ordinal=first_ordinal)

    def _assemble(self, n, name_buf, name_off, seq_buf, seq_off, qual_buf, v, j):
        return self.ctx.assemble_buffers(n, name_buf, name_off, seq_buf, seq_off, qual_buf, v, j)

    def mainflow_buffers(self, n, name_buf, name_off, seq_buf, seq_off, qual_buf):
        """Every rank passes the same whole-sample buffers; returns the Result on `root`, None elsewhere."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        lo, hi = shard_bounds(n, world, rank)
        seq_off = np.asarray(seq_off, dtype=np.int64)
        base = int(seq_off[lo])
        sub_off = seq_off[lo:hi + 1] - base
        sub_seq = np.frombuffer(seq_buf, dtype=np.uint8)[base:int(seq_off[hi])] if not isinstance(seq_buf, np.ndarray) else \
            seq_buf[base:int(seq_off[hi])]
        if self.on_device:                       # records never leave HBM: NCCL all-gather between the contexts' record arrays
            import time
            t = [time.perf_counter()]
            self.ctx.align_shard_device(hi - lo, np.ascontiguousarray(sub_seq), np.ascontiguousarray(sub_off), lo)
            t.append(time.perf_counter())
            max_len = gather_reads_device(self.ctx, hi - lo, n, self.group, self.device)
            t.append(time.perf_counter())
            gather_records_device(self.ctx, hi - lo, n, self.group, self.device)
            t.append(time.perf_counter())
            res = None
            if rank == self.root:
                res = self.ctx.assemble_resident(n, max_len, name_buf, np.asarray(name_off, dtype=np.int64), seq_off, qual_buf)
            t.append(time.perf_counter())
            self.laps_ms = dict(zip(("align", "gather_reads", "gather_records", "assemble"), [(b - a) * 1e3 for a, b in zip(t, t[1:])]))
            return res
        v, j = self.align_fn(hi - lo, np.ascontiguousarray(sub_seq), np.ascontiguousarray(sub_off), lo)
        v_all = gather_records(v, n, self.group, self.device)
        j_all = gather_records(j, n, self.group, self.device)
        if rank != self.root:
            return None
        return self.assemble_fn(n, name_buf, name_off, seq_buf, seq_off, qual_buf, v_all, j_all)
