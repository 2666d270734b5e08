"""ctypes binding of the CPU oracle (oracle/catt_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product (catt_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcatt_oracle.so")

REC_FIELDS = ["ordinal", "rid", "strand", "mapped", "AS", "qb", "m_end", "rb", "qe", "re", "nm", "mi", "ncand",
              "ntied", "cells_lo", "cells_hi"]
INFO_FIELDS = ["kmer", "the_err", "avg_len", "n_reads", "v_hits", "j_hits", "v_gapped", "j_gapped", "v_tied", "j_tied",
               "v_final", "j_final", "share", "pair_seq_mismatch", "full_pushed", "vpart", "jpart", "direct", "left",
               "rescued", "cand_v", "cand_j", "cells_v", "cells_j", "qual_oob", "t_align", "t_extract", "t_extend"]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "catt_oracle.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


class CoConfig(C.Structure):
    _fields_ = [("cmotif", C.c_char_p), ("fmotif", C.c_char_p), ("innerC", C.c_char_p), ("innerF", C.c_char_p),
                ("coffset", C.c_int), ("foffset", C.c_int), ("min_score", C.c_int), ("kmer", C.c_int),
                ("nthreads", C.c_int), ("allow_v", C.c_int), ("allow_j", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.co_refset_load.restype = C.c_void_p
        L.co_refset_load.argtypes = [C.c_char_p, C.c_int]
        L.co_refset_free.argtypes = [C.c_void_p]
        L.co_refset_nrec.argtypes = [C.c_void_p]
        L.co_refset_len.argtypes = [C.c_void_p]
        L.co_refset_name.restype = C.c_char_p
        L.co_refset_name.argtypes = [C.c_void_p, C.c_int]
        L.co_refset_reclen.argtypes = [C.c_void_p, C.c_int]
        L.co_refset_recoff.argtypes = [C.c_void_p, C.c_int]
        L.co_refset_pac_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.co_refset_bases.argtypes = [C.c_void_p, C.c_void_p]
        L.co_align.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_void_p,
                               C.c_void_p, C.c_int, C.c_int]
        L.co_seed.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_char_p), C.c_void_p]
        L.co_sw.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.co_sw_facts.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p]
        L.co_hash64.restype = C.c_uint64
        L.co_hash64.argtypes = [C.c_uint64]
        L.co_finder.argtypes = [C.POINTER(CoConfig), C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.co_motif_expand.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.co_global_affine.argtypes = [C.c_char_p, C.c_char_p]
        L.co_best_d.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p)]
        L.co_run.restype = C.c_void_p
        L.co_run.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CoConfig), C.c_char_p, C.c_int64, C.c_int]
        L.co_run_mem.restype = C.c_void_p
        L.co_run_mem.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CoConfig), C.c_int64, C.POINTER(C.c_char_p),
                                 C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int]
        L.co_result_free.argtypes = [C.c_void_p]
        L.co_result_info.argtypes = [C.c_void_p, C.c_void_p]
        L.co_result_count.restype = C.c_int64
        L.co_result_count.argtypes = [C.c_void_p, C.c_int]
        L.co_result_dump.restype = C.c_char_p
        L.co_result_dump.argtypes = [C.c_void_p, C.c_int]
        L.co_result_records.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.co_link_leaves.restype = None
        L.co_link_leaves.argtypes = [C.c_int, C.c_char_p, C.c_int64, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.co_near_any.restype = None
        L.co_near_any.argtypes = [C.c_int, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _strarr(items):
    arr = (C.c_char_p * len(items))()
    arr[:] = [s.encode() if isinstance(s, str) else s for s in items]
    return arr


class RefSet:
    def __init__(self, fasta: str, use_pac: bool = True):
        self.h = lib().co_refset_load(fasta.encode(), int(use_pac))
        if not self.h:
            raise FileNotFoundError(fasta)
        L = lib()
        self.nrec = L.co_refset_nrec(self.h)
        self.L = L.co_refset_len(self.h)
        self.names = [L.co_refset_name(self.h, i).decode() for i in range(self.nrec)]
        self.lens = [L.co_refset_reclen(self.h, i) for i in range(self.nrec)]
        self.offs = [L.co_refset_recoff(self.h, i) for i in range(self.nrec)]

    def pac_stats(self):
        a, b = C.c_int(), C.c_int()
        lib().co_refset_pac_stats(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def bases(self) -> np.ndarray:
        out = np.zeros(self.L, dtype=np.uint8)
        lib().co_refset_bases(self.h, out.ctypes.data)
        return out

    def record(self, i: int) -> str:
        b = self.bases()[self.offs[i]:self.offs[i] + self.lens[i]]
        return "".join("ACGT"[x] for x in b)

    def align(self, seqs, first_ordinal=0, min_score=12, max_cand=0, threads=0):
        n = len(seqs)
        out = np.zeros((n, 16), dtype=np.int32)
        cand = np.full((n, max_cand, 3), -1, dtype=np.int32) if max_cand else None
        lib().co_align(self.h, n, _strarr(seqs), first_ordinal, min_score, out.ctypes.data,
                       cand.ctypes.data if cand is not None else None, max_cand, threads)
        return (out, cand) if max_cand else out

    def seed(self, seqs) -> np.ndarray:
        n = len(seqs)
        out = np.zeros((n, self.nrec * 2), dtype=np.uint8)
        lib().co_seed(self.h, n, _strarr(seqs), out.ctypes.data)
        return out

    def __del__(self):
        try:
            if self.h:
                lib().co_refset_free(self.h)
        except Exception:
            pass


@dataclass
class Myread:
    seq: str
    qual: str
    vs: str
    js: str
    cdr3: str
    able: bool
    ordinal: int = -1
    flags: int = 0


def make_config(chain_cfg: dict, min_score=12, kmer=10000, nthreads=1, allow_v=3, allow_j=9) -> CoConfig:
    return CoConfig(chain_cfg["cmotif"].encode(), chain_cfg["fmotif"].encode(), chain_cfg["innerC"].encode(),
                    chain_cfg["innerF"].encode(), chain_cfg["coffset"], chain_cfg["foffset"], min_score, kmer, nthreads,
                    allow_v, allow_j)


class RunResult:
    def __init__(self, h):
        if not h:
            raise RuntimeError("oracle run failed (bad input file?)")
        self.h = h
        buf = np.zeros(64, dtype=np.float64)
        n = lib().co_result_info(h, buf.ctypes.data)
        self.info = {k: (float(buf[i]) if k in ("the_err", "avg_len", "t_align", "t_extract", "t_extend") else int(buf[i]))
                     for i, k in enumerate(INFO_FIELDS[:n])}

    def part(self, which: int):
        txt = lib().co_result_dump(self.h, which).decode()
        out = []
        for line in txt.split("\n"):
            if not line:
                continue
            f = line.split("\t")
            out.append(Myread(f[0], f[1], f[2], f[3], f[4], f[5] == "1", int(f[6]), int(f[7])))
        return out

    @property
    def Vpart(self):
        return self.part(0)

    @property
    def Jpart(self):
        return self.part(1)

    def digest(self, only_cdr3=False):
        """(Vpart, Jpart) canonical digests - the twin of catt_b200's Result.text_digest()."""
        L = lib()
        L.co_result_digest.restype = C.c_uint64
        L.co_result_digest.argtypes = [C.c_void_p, C.c_int, C.c_int]
        return int(L.co_result_digest(self.h, 0, int(only_cdr3))), int(L.co_result_digest(self.h, 1, int(only_cdr3)))

    def records(self, which: int) -> np.ndarray:
        n = self.info["n_reads"]
        out = np.zeros((n, 16), dtype=np.int32)
        lib().co_result_records(self.h, which, out.ctypes.data)
        return out

    def __del__(self):
        try:
            lib().co_result_free(self.h)
        except Exception:
            pass


def run_file(vref: RefSet, jref: RefSet, cfg: CoConfig, fastq: str, max_reads=0, threads=0) -> RunResult:
    return RunResult(lib().co_run(vref.h, jref.h, C.byref(cfg), fastq.encode(), max_reads, threads))


def run_mem(vref: RefSet, jref: RefSet, cfg: CoConfig, names, seqs, quals, threads=0) -> RunResult:
    n = len(seqs)
    return RunResult(lib().co_run_mem(vref.h, jref.h, C.byref(cfg), n, _strarr(names) if names is not None else None,
                                      _strarr(seqs), _strarr(quals) if quals is not None else None, threads))


def run_mem_files(vref: RefSet, jref: RefSet, cfg: CoConfig, files, threads=0) -> RunResult:
    """Paired-end / multi-file sample: `files` = [(names, seqs, quals), ...] in --f1, --f2 order (catt.jl:738-745)."""
    names = [x for f in files for x in f[0]]
    seqs = [x for f in files for x in f[1]]
    quals = [x for f in files for x in f[2]]
    fs = np.zeros(len(files) + 1, dtype=np.int64)
    np.cumsum([len(f[1]) for f in files], out=fs[1:])
    L = lib()
    L.co_run_mem_files.restype = C.c_void_p
    L.co_run_mem_files.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CoConfig), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_int]
    return RunResult(L.co_run_mem_files(vref.h, jref.h, C.byref(cfg), len(seqs), _strarr(names), _strarr(seqs), _strarr(quals),
                                        len(files), fs.ctypes.data, threads))


def real_score(refset: RefSet, seqs, rid, strand, qb, m_end, rb, allowance) -> np.ndarray:
    """real_score (catt.jl:58-84) per SAM record; seqs are the raw reads."""
    L = lib()
    L.co_real_score.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_char_p)] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p]
    a = [np.ascontiguousarray(x, dtype=np.int32) for x in (rid, strand, qb, m_end, rb)]
    out = np.zeros(len(seqs), dtype=np.uint8)
    L.co_real_score(refset.h, len(seqs), _strarr(seqs), *[x.ctypes.data for x in a], allowance, out.ctypes.data)
    return out


def order_items(names, vrec16: np.ndarray, jrec16: np.ndarray, nthreads=1):
    """The ordering stage alone: (items [t, 3] = kind, read, read2; counts = final V, final J, V-only, pairs, J-only)."""
    L = lib()
    L.co_order.restype = C.c_int64
    L.co_order.argtypes = [C.c_int64, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    n = len(names)
    v = np.ascontiguousarray(vrec16, dtype=np.int32)
    j = np.ascontiguousarray(jrec16, dtype=np.int32)
    items = np.zeros((2 * n + 1, 3), dtype=np.int32)
    counts = np.zeros(5, dtype=np.int64)
    t = L.co_order(n, _strarr(names), v.ctypes.data, j.ctypes.data, nthreads, items.ctypes.data, counts.ctypes.data)
    return items[:t].copy(), counts


def sw(q: str, t: str):
    a, b = C.c_int(), C.c_int()
    s = lib().co_sw(q.encode(), t.encode(), C.byref(a), C.byref(b))
    return s, a.value, b.value


def sw_facts(q: str, t: str):
    out = np.zeros(8, dtype=np.int32)
    s = lib().co_sw_facts(q.encode(), t.encode(), out.ctypes.data)
    return s, dict(zip(["qb", "qe", "rb", "re", "first_m", "nm", "mi", "n_gap_ops"], out.tolist()))


def hash64(k: int) -> int:
    return lib().co_hash64(k)


def finder(cfg: CoConfig, seq: str, array_version=False):
    lo, hi = C.c_int(), C.c_int()
    r = lib().co_finder(C.byref(cfg), seq.encode(), int(array_version), C.byref(lo), C.byref(hi))
    return bool(r & 1), (lo.value, hi.value) if r & 2 else None


def motif_expand(re_: str):
    ln = C.c_int()
    n = lib().co_motif_expand(re_.encode(), C.byref(ln))
    return n, ln.value


def global_affine(a: str, b: str) -> int:
    return lib().co_global_affine(a.encode(), b.encode())


def best_d(cdr3: str, dlist) -> str:
    i = lib().co_best_d(cdr3.encode(), len(dlist), _strarr([d[1] for d in dlist]))
    return "None" if i < 0 else dlist[i][0]


def synth_reads(vref: RefSet, jref: RefSet, n, length=150, first_index=0, seed=20251019, p_err=0.003, with_n=False):
    """SURVEY §8d config-3 generator (C++ twin of tests/util.py::synth_read).  Returns (seqs, quals) uint8 arrays [n, length]."""
    L = lib()
    L.co_synth_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_uint64, C.c_double, C.c_int,
                                 C.c_void_p, C.c_void_p]
    seqs = np.empty((n, length), dtype=np.uint8)
    quals = np.empty((n, length), dtype=np.uint8)
    L.co_synth_reads(vref.h, jref.h, n, first_index, length, seed, p_err, int(with_n), seqs.ctypes.data, quals.ctypes.data)
    return seqs, quals


def run_flat(vref: RefSet, jref: RefSet, cfg: CoConfig, seqs: np.ndarray, quals: np.ndarray, first_index=0, threads=0) -> RunResult:
    L = lib()
    L.co_run_flat.restype = C.c_void_p
    L.co_run_flat.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CoConfig), C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    n, length = seqs.shape
    return RunResult(L.co_run_flat(vref.h, jref.h, C.byref(cfg), n, first_index, length, seqs.ctypes.data, quals.ctypes.data, threads))


def link_leaves(lgt: int, roots, uplimit, leaves, leaf_count):
    """Jtool.jl:396-421 over every leaf: (statu, holder) int32 arrays.  roots / leaves: lists of equal-length strings."""
    up = np.ascontiguousarray(np.asarray(uplimit, dtype=np.float64).reshape(len(roots), 7))
    cnt = np.ascontiguousarray(leaf_count, dtype=np.int64)
    statu, hold = np.zeros(len(leaves), dtype=np.int32), np.zeros(len(leaves), dtype=np.int32)
    lib().co_link_leaves(lgt, "".join(roots).encode(), len(roots), up.ctypes.data, "".join(leaves).encode(), cnt.ctypes.data, len(leaves),
                         statu.ctypes.data, hold.ctypes.data)
    return statu, hold


def near_any(lgt: int, queries, targets, upper=3, skip=3):
    """Jtool.jl:553-558: per query, is any target within `upper` mismatches (ends skipped)?"""
    flag = np.zeros(len(queries), dtype=np.uint8)
    lib().co_near_any(lgt, "".join(queries).encode(), len(queries), "".join(targets).encode(), len(targets), upper, skip, flag.ctypes.data)
    return flag
