"""Host-side mirror of the reference's interface for the CDR3 hot path, on top of the C ABI.

The reference is Julia (absent from this image), so this layer is Python; INTEGRATION.md shows the equivalent
`ccall` stubs for catt.jl.  Names and argument meaning follow the reference:

  reference (catt.jl / Jtool.jl)                         here
  ------------------------------------------------------------------------------------------
  parsed_args + refg[species][chain][region]          -> Catt(species, chain, region, resource_dir, ...)
  input_convert + read_alignrs + pool/extension part
      of catt()   (catt.jl:134-151,186-339,544-577)   -> Catt.mainflow(path) / Catt.mainflow_reads(...)
  Myread (Jtool.jl:12-19)                              -> Myread
  selBestMatchD loop (catt.jl:676-683)                -> Catt.selBestMatchD(seqs)
  the_ERR / args["kmer"] (catt.jl:26,162-171)         -> result.info["the_err"], result.info["kmer"]

Everything computational happens in libcatt_b200.so on the GPU; this module only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .refg import chain_config


@dataclass
class Myread:                      # Jtool.jl:12-19
    seq: str
    qual: str
    vs: str
    js: str
    cdr3: str
    able: bool
    ordinal: int = -1
    flags: int = 0


class Result:
    """Vpart / Jpart exactly as catt.jl:579 expects them, plus the log counters and per-read alignment records."""

    def __init__(self, owner: "Catt", handle):
        self._owner, self._h = owner, handle
        L = _lib.load()
        info = _lib.Info()
        _lib.check(L.catt_result_info(handle, C.byref(info)))
        self.info = info.as_dict()

    def _part(self, which):
        L = _lib.load()
        p, n = C.POINTER(_lib.Read)(), C.c_int64()
        _lib.check(L.catt_result_reads(self._h, which, C.byref(p), C.byref(n)))
        vn, jn = self._owner.v_names, self._owner.j_names
        out = []
        for i in range(n.value):
            r = p[i]
            seq = C.string_at(r.seq, r.seq_len).decode()
            qual = C.string_at(r.qual, r.qual_len).decode()
            cdr3 = seq[r.cdr3_off:r.cdr3_off + r.cdr3_len] if r.cdr3_off >= 0 else "None"
            out.append(Myread(seq, qual, vn[r.v_id] if r.v_id >= 0 else "None", jn[r.j_id] if r.j_id >= 0 else "None", cdr3,
                              bool(r.able), r.ordinal, r.flags))
        return out

    @property
    def Vpart(self):
        return self._part(0)

    @property
    def Jpart(self):
        return self._part(1)

    def records(self, which) -> np.ndarray:
        L = _lib.load()
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(L.catt_align_records(self._h, which, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=_lib.ALN_DTYPE)
        buf = (C.c_char * (n.value * _lib.ALN_DTYPE.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=_lib.ALN_DTYPE).copy()

    def clonotypes(self):
        """The clonotype table (Catt(clonotypes=True)): [(cdr3, merged qual, V name, J name, count)], ordered by (len, cdr3, V id, J id)."""
        L = _lib.load()
        p, n = C.POINTER(_lib.Clonotype)(), C.c_int64()
        _lib.check(L.catt_result_clonotypes(self._h, C.byref(p), C.byref(n)))
        vn, jn = self._owner.v_names, self._owner.j_names
        return [(p[i].cdr3.decode(), p[i].qual.decode(), vn[p[i].v_id] if p[i].v_id >= 0 else "None", jn[p[i].j_id] if p[i].j_id >= 0 else "None",
                 p[i].count) for i in range(n.value)]

    def digest(self):
        """(Vpart digest, Jpart digest): order-sensitive 64-bit digests of every field and string of the two lists."""
        out = (C.c_uint64 * 2)()
        _lib.check(_lib.load().catt_result_digest(self._h, out))
        return int(out[0]), int(out[1])

    def text_digest(self):
        """(Vpart, Jpart) digests over the lists as text (seq, qual, V name, J name, cdr3, able): comparable across implementations."""
        out = (C.c_uint64 * 2)()
        _lib.check(_lib.load().catt_result_text_digest(self._owner._h, self._h, out))
        return int(out[0]), int(out[1])

    def close(self):
        if self._h:
            _lib.load().catt_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Batch:
    """A shard of reads resident in HBM (bench.py / multi-GPU sharding)."""

    def __init__(self, owner, handle, n):
        self._owner, self._h, self.n = owner, handle, n

    def align(self):
        _lib.check(_lib.load().catt_batch_align(self._owner._h, self._h))

    def sync(self) -> dict:
        info = _lib.Info()
        _lib.check(_lib.load().catt_batch_sync(self._owner._h, self._h, C.byref(info)))
        return info.as_dict()

    def download(self):
        v = np.zeros(self.n, dtype=_lib.ALN_DTYPE)
        j = np.zeros(self.n, dtype=_lib.ALN_DTYPE)
        _lib.check(_lib.load().catt_batch_download(self._owner._h, self._h, v.ctypes.data, j.ctypes.data))
        return v, j

    def assemble(self, name_buf, name_off, seq_buf, seq_off, qual_buf) -> "Result":
        h = C.c_void_p()
        _lib.check(_lib.load().catt_batch_assemble(self._owner._h, self._h, _addr(name_buf), name_off.ctypes.data, _addr(seq_buf),
                                                   seq_off.ctypes.data, _addr(qual_buf), C.byref(h)))
        return Result(self._owner, h)

    def close(self):
        if self._h:
            _lib.load().catt_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def bam_to_fastq(bam_path: str, fastq_path: str) -> None:
    """The `samtools fastq -N in.bam > out.fastq` step of catt.jl:760 (BAM or SAM in, FASTQ out); needs no GPU."""
    _lib.check(_lib.load().catt_bam_to_fastq(bam_path.encode(), fastq_path.encode()))


def inflate_file(path: str, out_path: str) -> int:
    """Write the bytes the library reads from a plain / BGZF / gzip input file; returns the method (0 plain, 1 BGZF in parallel,
    2 parallel gzip, 3 one zlib stream).  Needs no GPU."""
    m = C.c_int32(-1)
    _lib.check(_lib.load().catt_inflate_file(path.encode(), out_path.encode(), C.byref(m)))
    return m.value


def set_host_threads(n: int = 0):
    """Host threads for read packing / result materialisation (0 = all hardware threads); one process per GPU -> cores // ranks."""
    _lib.check(_lib.load().catt_set_host_threads(n))


def comm_id() -> bytes:
    """128-byte id of a new multi-process GPU group (rank 0 creates it, the launcher's channel carries it to the other ranks)."""
    buf = (C.c_uint8 * 128)()
    _lib.check(_lib.load().catt_comm_id(buf))
    return bytes(buf)


class Catt:
    """One CATT configuration bound to one GPU (catt_ctx) - or, with `ngpu` > 1, to a group of GPUs of this process that
    share every sample (catt_config.ngpu / devices; same calls, same results)."""

    def __init__(self, species="hs", chain="TRB", region="CDR3", resource_dir=None, bowsc=12, kmer=10000, thread=1,
                 allow_v=3, allow_j=9, device=0, chain_cfg=None, keep_records=False, chunk_reads=0, only_cdr3=False, ngpu=1, devices=None,
                 clonotypes=False):
        L = _lib.load()
        cc = chain_cfg or chain_config(species, chain, region, resource_dir)
        self.chain_cfg = cc
        d = cc.get("Dregion", [])
        self._keep = [(C.c_char_p * max(len(d), 1))(*[x[0].encode() for x in d]),
                      (C.c_char_p * max(len(d), 1))(*[x[1].encode() for x in d])]
        cfg = _lib.Config(cc["vregion"].encode(), cc["jregion"].encode(), cc["cmotif"].encode(), cc["fmotif"].encode(),
                          cc["innerC"].encode(), cc["innerF"].encode(), cc["coffset"], cc["foffset"], bowsc, kmer, thread,
                          allow_v, allow_j, device, len(d), self._keep[0], self._keep[1], int(keep_records), int(chunk_reads), int(only_cdr3),
                          int(ngpu), (C.c_int32 * len(devices))(*devices) if devices else None, int(clonotypes), 0)
        if devices and len(devices) != ngpu:
            raise ValueError("one device ordinal per rank")
        h = C.c_void_p()
        _lib.check(L.catt_ctx_create(C.byref(h), C.byref(cfg)))
        self._h = h
        self.d_list = d
        self.v_names = [L.catt_ref_name(h, 0, i).decode() for i in range(L.catt_ref_count(h, 0))]
        self.j_names = [L.catt_ref_name(h, 1, i).decode() for i in range(L.catt_ref_count(h, 1))]

    # -- reference records (alignment bases) -------------------------------------------------------------------
    def ref_seq(self, which, i) -> str:
        buf = C.create_string_buffer(1024)
        n = _lib.load().catt_ref_seq(self._h, which, i, buf, 1024)
        if n < 0:
            _lib.check(n)
        return buf.raw[:n].decode()

    # -- the hot path ------------------------------------------------------------------------------------------
    def mainflow(self, fastx_path: str) -> Result:
        """input_convert + read_alignrs + pool/extension for one single-end file (catt.jl:767-768 up to :577)."""
        h = C.c_void_p()
        _lib.check(_lib.load().catt_run_file(self._h, fastx_path.encode(), C.byref(h)))
        return Result(self, h)

    def mainflow_bam(self, bam_path: str) -> Result:
        """--bam input (catt.jl:756-768): what `samtools fastq -N` would extract from a BAM / SAM file, through the single-end path."""
        h = C.c_void_p()
        _lib.check(_lib.load().catt_run_bam(self._h, bam_path.encode(), C.byref(h)))
        return Result(self, h)

    def mainflow_paired(self, *fastx_paths: str) -> Result:
        """Paired-end sample (--f1/--f2, catt.jl:738-745): each mate file goes through input_convert + the read_alignrs loop on
        its own, then one pool/extension pass over the joined lists."""
        arr = (C.c_char_p * len(fastx_paths))(*[p.encode() for p in fastx_paths])
        h = C.c_void_p()
        _lib.check(_lib.load().catt_run_files(self._h, len(fastx_paths), arr, C.byref(h)))
        return Result(self, h)

    def mainflow_reads_files(self, files) -> Result:
        """In-memory form of mainflow_paired: files = [(names, seqs, quals), ...]."""
        nb, no = _lib.concat([x for f in files for x in f[0]])
        sb, so = _lib.concat([x for f in files for x in f[1]])
        qb = _lib.concat([x for f in files for x in f[2]])[0]
        fs = np.zeros(len(files) + 1, dtype=np.int64)
        np.cumsum([len(f[1]) for f in files], out=fs[1:])
        h = C.c_void_p()
        _lib.check(_lib.load().catt_run_reads_files(self._h, int(fs[-1]), nb, no.ctypes.data, sb, so.ctypes.data, qb, len(files),
                                                    fs.ctypes.data, C.byref(h)))
        return Result(self, h)

    def mainflow_reads(self, names, seqs, quals=None) -> Result:
        nb, no = _lib.concat(names)
        sb, so = _lib.concat(seqs)
        qb = _lib.concat(quals)[0] if quals is not None else None
        h = C.c_void_p()
        _lib.check(_lib.load().catt_run_reads(self._h, len(seqs), nb, no.ctypes.data, sb, so.ctypes.data, qb, C.byref(h)))
        return Result(self, h)

    def mainflow_buffers(self, n, name_buf, name_off, seq_buf, seq_off, qual_buf) -> Result:
        """Zero-copy variant for large inputs: bytes-like buffers + int64 offset arrays."""
        h = C.c_void_p()
        _lib.check(_lib.load().catt_run_reads(self._h, n, _addr(name_buf), name_off.ctypes.data, _addr(seq_buf), seq_off.ctypes.data,
                                              _addr(qual_buf) if qual_buf is not None else None, C.byref(h)))
        return Result(self, h)

    # -- one sample over several GPUs, one process per GPU ------------------------------------------------------
    def join(self, nranks: int, rank: int, id128: bytes):
        """Collective: this context becomes rank `rank` of a multi-process GPU group (catt_ctx_join, ncclCommInitRank)."""
        buf = (C.c_uint8 * 128)(*id128)
        _lib.check(_lib.load().catt_ctx_join(self._h, nranks, rank, buf))

    def mainflow_shard(self, n_total, lo, n_local, name_buf, name_off, seq_buf, seq_off, qual_buf, file_start=None) -> Result:
        """Collective: every rank passes ITS block [lo, lo + n_local) of the sample (offsets local to its buffers); the lists
        come back on rank 0, the counters on every rank."""
        fs = np.ascontiguousarray(file_start if file_start is not None else [0, n_total], dtype=np.int64)
        h = C.c_void_p()
        _lib.check(_lib.load().catt_run_reads_shard(self._h, n_total, lo, n_local, _addr(name_buf), name_off.ctypes.data, _addr(seq_buf),
                                                    seq_off.ctypes.data, _addr(qual_buf) if qual_buf is not None else None, len(fs) - 1,
                                                    fs.ctypes.data, C.byref(h)))
        return Result(self, h)

    def mainflow_file_shard(self, fastq_path: str) -> Result:
        """Collective: every rank ingests its byte range of the same plain FASTQ file on its own GPU (catt_run_file_shard)."""
        h = C.c_void_p()
        _lib.check(_lib.load().catt_run_file_shard(self._h, fastq_path.encode(), C.byref(h)))
        return Result(self, h)

    def shard_upload(self, n_total, lo, n_local, name_buf, name_off, seq_buf, seq_off, qual_buf, file_start=None):
        """This rank's block to HBM (packed reads, QNAMEs, quality strings), for repeated `shard_run` calls."""
        fs = np.ascontiguousarray(file_start if file_start is not None else [0, n_total], dtype=np.int64)
        _lib.check(_lib.load().catt_shard_upload(self._h, n_total, lo, n_local, _addr(name_buf), name_off.ctypes.data, _addr(seq_buf),
                                                 seq_off.ctypes.data, _addr(qual_buf) if qual_buf is not None else None, len(fs) - 1, fs.ctypes.data))

    def shard_run(self) -> Result:
        """Collective: Stage A on the resident block + the global Stage B (lists on rank 0)."""
        h = C.c_void_p()
        _lib.check(_lib.load().catt_shard_run(self._h, C.byref(h)))
        return Result(self, h)

    def selBestMatchD(self, seqs):
        """Jtool.jl:40-52 for a list of CDR3 nucleotide strings -> D segment names ("None" when nothing passes)."""
        if not seqs:
            return []
        arr = (C.c_char_p * len(seqs))(*[s.encode() for s in seqs])
        out = np.zeros(len(seqs), dtype=np.int32)
        _lib.check(_lib.load().catt_best_d(self._h, arr, len(seqs), out.ctypes.data))
        return ["None" if i < 0 else self.d_list[i][0] for i in out]

    # -- stage level -------------------------------------------------------------------------------------------
    def align_shard(self, seqs, first_ordinal=0):
        sb, so = _lib.concat(seqs)
        n = len(seqs)
        v = np.zeros(n, dtype=_lib.ALN_DTYPE)
        j = np.zeros(n, dtype=_lib.ALN_DTYPE)
        _lib.check(_lib.load().catt_align_shard(self._h, n, sb, so.ctypes.data, first_ordinal, v.ctypes.data, j.ctypes.data))
        return v, j

    def align_shard_buffers(self, n, seq_buf, seq_off, first_ordinal=0):
        v = np.zeros(n, dtype=_lib.ALN_DTYPE)
        j = np.zeros(n, dtype=_lib.ALN_DTYPE)
        _lib.check(_lib.load().catt_align_shard(self._h, n, _addr(seq_buf), seq_off.ctypes.data, first_ordinal, v.ctypes.data,
                                                j.ctypes.data))
        return v, j

    def align_shard_device(self, n, seq_buf, seq_off, first_ordinal=0):
        """Stage A of a shard; the records stay in HBM (at the start of the context's record arrays)."""
        _lib.check(_lib.load().catt_align_shard(self._h, n, _addr(seq_buf), seq_off.ctypes.data, first_ordinal, None, None))

    def records_device(self, which, capacity) -> int:
        """Device address of the context's record array for `which` (0 = V, 1 = J), grown to `capacity` records."""
        p = C.c_uint64()
        _lib.check(_lib.load().catt_records_device(self._h, which, capacity, C.byref(p)))
        return p.value

    def assemble_device(self, n, name_buf, name_off, seq_buf, seq_off, qual_buf) -> Result:
        """Stage B on the n records already sitting in the context's record arrays (after an NCCL all-gather)."""
        h = C.c_void_p()
        _lib.check(_lib.load().catt_assemble_device(self._h, n, _addr(name_buf), name_off.ctypes.data, _addr(seq_buf), seq_off.ctypes.data,
                                                    _addr(qual_buf) if qual_buf is not None else None, C.byref(h)))
        return Result(self, h)

    def shard_counts(self):
        """(reads, packed words, longest read) of what the last align_shard* call left in the context's read buffers."""
        n, w, m = C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(_lib.load().catt_shard_counts(self._h, C.byref(n), C.byref(w), C.byref(m)))
        return n.value, w.value, m.value

    def reads_device(self, cap_reads, cap_words):
        """Device addresses (words, offsets, lengths) of the context's packed-read buffers, grown to the given capacities."""
        a, b, c_ = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _lib.check(_lib.load().catt_reads_device(self._h, cap_reads, cap_words, C.byref(a), C.byref(b), C.byref(c_)))
        return a.value, b.value, c_.value

    def assemble_resident(self, n, max_len, name_buf, name_off, seq_off, qual_buf) -> Result:
        """Stage B on n reads + records that are all resident in the context's buffers (after the NCCL gathers)."""
        h = C.c_void_p()
        _lib.check(_lib.load().catt_assemble_resident(self._h, n, max_len, _addr(name_buf), name_off.ctypes.data, seq_off.ctypes.data,
                                                      _addr(qual_buf) if qual_buf is not None else None, C.byref(h)))
        return Result(self, h)

    @staticmethod
    def _flat(seqs, lgt):
        """n x lgt bytes, back to back: a list of equal-length strings or an (n, lgt) uint8 array."""
        if isinstance(seqs, np.ndarray):
            a = np.ascontiguousarray(seqs, dtype=np.uint8)
            if a.ndim != 2 or a.shape[1] != lgt:
                raise ValueError("sequence array must be (n, lgt) uint8")
            return a, a.shape[0]
        if any(len(s) != lgt for s in seqs):
            raise ValueError("all sequences of one nbsorb group have the same length")
        return np.frombuffer("".join(seqs).encode(), dtype=np.uint8), len(seqs)

    def link_leaves(self, lgt, roots, uplimit, leaves, leaf_count):
        """linkRLG for every leaf (Jtool.jl:396-421, 524-531): (statu, root_index) int32 arrays."""
        r, nr = self._flat(roots, lgt)
        l, nl = self._flat(leaves, lgt)
        up = np.ascontiguousarray(np.asarray(uplimit, dtype=np.float64).reshape(nr, 7))
        cnt = np.ascontiguousarray(leaf_count, dtype=np.int64)
        if cnt.shape != (nl,):
            raise ValueError("one count per leaf")
        statu, hold = np.zeros(nl, dtype=np.int32), np.zeros(nl, dtype=np.int32)
        _lib.check(_lib.load().catt_link_leaves(self._h, lgt, r.ctypes.data, nr, up.ctypes.data, l.ctypes.data, cnt.ctypes.data, nl,
                                                statu.ctypes.data, hold.ctypes.data))
        return statu, hold

    def near_any(self, lgt, queries, targets, upper=3, skip=3):
        """Jtool.jl:553-558: per query, is any target within `upper` mismatches of it (first / last `skip` positions ignored)?"""
        q, nq = self._flat(queries, lgt)
        t, nt = self._flat(targets, lgt)
        flag = np.zeros(nq, dtype=np.uint8)
        _lib.check(_lib.load().catt_near_any(self._h, lgt, q.ctypes.data, nq, t.ctypes.data, nt, upper, skip, flag.ctypes.data))
        return flag

    def upload(self, n, seq_buf, seq_off, first_ordinal=0) -> Batch:
        h = C.c_void_p()
        _lib.check(_lib.load().catt_batch_upload(self._h, n, _addr(seq_buf), seq_off.ctypes.data, first_ordinal, C.byref(h)))
        return Batch(self, h, n)

    def assemble(self, names, seqs, quals, vrec, jrec) -> Result:
        nb, no = _lib.concat(names)
        sb, so = _lib.concat(seqs)
        qb = _lib.concat(quals)[0] if quals is not None else None
        h = C.c_void_p()
        vrec = np.ascontiguousarray(vrec)
        jrec = np.ascontiguousarray(jrec)
        _lib.check(_lib.load().catt_assemble(self._h, len(seqs), nb, no.ctypes.data, sb, so.ctypes.data, qb, vrec.ctypes.data,
                                             jrec.ctypes.data, C.byref(h)))
        return Result(self, h)

    def assemble_buffers(self, n, name_buf, name_off, seq_buf, seq_off, qual_buf, vrec, jrec) -> Result:
        """catt_assemble on caller buffers (bytes-like + int64 offsets) and the gathered alignment records."""
        h = C.c_void_p()
        vrec = np.ascontiguousarray(vrec)
        jrec = np.ascontiguousarray(jrec)
        _lib.check(_lib.load().catt_assemble(self._h, n, _addr(name_buf), name_off.ctypes.data, _addr(seq_buf), seq_off.ctypes.data,
                                             _addr(qual_buf) if qual_buf is not None else None, vrec.ctypes.data, jrec.ctypes.data,
                                             C.byref(h)))
        return Result(self, h)

    def synth_reads(self, n, length=150, first_index=0, seed=20251019, p_err=0.003, with_n=False):
        """SURVEY §8d config-3 generator.  Returns (name_buf, name_off, seq_buf, seq_off, qual_buf) as numpy arrays."""
        seqs = np.empty(n * length, dtype=np.uint8)
        quals = np.empty(n * length, dtype=np.uint8)
        names = np.zeros(n * 24 + 24, dtype=np.uint8)
        name_off = np.zeros(n + 1, dtype=np.int64)
        _lib.check(_lib.load().catt_synth_reads(self._h, n, first_index, length, seed, p_err, int(with_n), seqs.ctypes.data,
                                                quals.ctypes.data, names.ctypes.data, name_off.ctypes.data))
        seq_off = np.arange(n + 1, dtype=np.int64) * length
        return names, name_off, seqs, seq_off, quals

    def stream(self) -> int:
        return _lib.load().catt_ctx_stream(self._h)

    # -- parity hooks -------------------------------------------------------------------------------------------
    def seed_candidates(self, which, seqs) -> np.ndarray:
        sb, so = _lib.concat(seqs)
        nrec = len(self.j_names if which else self.v_names)
        out = np.zeros((len(seqs), nrec * 2), dtype=np.uint8)
        _lib.check(_lib.load().catt_seed_candidates(self._h, which, len(seqs), sb, so.ctypes.data, out.ctypes.data))
        return out

    def sw_pairs(self, which, seqs, rid, strand) -> np.ndarray:
        sb, so = _lib.concat(seqs)
        rid = np.ascontiguousarray(rid, dtype=np.int32)
        strand = np.ascontiguousarray(strand, dtype=np.int32)
        out = np.zeros((len(seqs), 3), dtype=np.int32)
        _lib.check(_lib.load().catt_sw_pairs(self._h, which, len(seqs), sb, so.ctypes.data, rid.ctypes.data, strand.ctypes.data,
                                             out.ctypes.data))
        return out

    def finder(self, seqs, array_version=False) -> np.ndarray:
        sb, so = _lib.concat(seqs)
        out = np.zeros((len(seqs), 3), dtype=np.int32)
        _lib.check(_lib.load().catt_finder(self._h, int(array_version), len(seqs), sb, so.ctypes.data, out.ctypes.data))
        return out

    def real_score(self, which, seqs, rid, strand, qb, m_end, rb, allowance) -> np.ndarray:
        """real_score (catt.jl:58-84) alone, one SAM record per read."""
        sb, so = _lib.concat(seqs)
        a = [np.ascontiguousarray(x, dtype=np.int32) for x in (rid, strand, qb, m_end, rb)]
        out = np.zeros(len(seqs), dtype=np.uint8)
        _lib.check(_lib.load().catt_real_score(self._h, which, len(seqs), sb, so.ctypes.data, *[x.ctypes.data for x in a], allowance, out.ctypes.data))
        return out

    def order_records(self, names, seq_lens, vrec, jrec):
        """The ordering stage alone: (items [t, 3] = kind, read, read2; counts = final V, final J, V-only, pairs, J-only)."""
        nb, no = _lib.concat(names)
        n = len(names)
        vrec, jrec = np.ascontiguousarray(vrec), np.ascontiguousarray(jrec)
        lens = np.ascontiguousarray(seq_lens, dtype=np.int32)
        items = np.zeros((2 * n + 1, 3), dtype=np.int32)
        cnt = np.zeros(5, dtype=np.int64)
        t = C.c_int64()
        _lib.check(_lib.load().catt_order_records(self._h, n, nb, no.ctypes.data, lens.ctypes.data, vrec.ctypes.data, jrec.ctypes.data,
                                                  items.ctypes.data, C.byref(t), cnt.ctypes.data))
        return items[:t.value].copy(), cnt

    def alu_peak(self):
        ops, ms = C.c_double(), C.c_double()
        _lib.check(_lib.load().catt_alu_peak(self._h, C.byref(ops), C.byref(ms)))
        return ops.value, ms.value

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().catt_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _addr(buf):
    if buf is None:
        return None
    if isinstance(buf, np.ndarray):
        return buf.ctypes.data
    if isinstance(buf, bytearray):
        return C.addressof((C.c_char * len(buf)).from_buffer(buf))
    return buf      # bytes: ctypes passes the address of the object's buffer for a c_void_p parameter
