"""ctypes binding of libcatt_b200.so (include/catt_b200.h).  Fails loudly when the library is missing: there is
no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcatt_b200.so")


class CattError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libcatt_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("v_fasta", C.c_char_p), ("j_fasta", C.c_char_p), ("cmotif", C.c_char_p), ("fmotif", C.c_char_p),
                ("inner_c", C.c_char_p), ("inner_f", C.c_char_p), ("coffset", C.c_int32), ("foffset", C.c_int32),
                ("min_score", C.c_int32), ("kmer", C.c_int32), ("julia_nthreads", C.c_int32), ("allow_v", C.c_int32),
                ("allow_j", C.c_int32), ("device", C.c_int32), ("n_d", C.c_int32), ("d_names", C.POINTER(C.c_char_p)),
                ("d_seqs", C.POINTER(C.c_char_p)), ("keep_records", C.c_int32), ("chunk_reads", C.c_int32), ("only_cdr3", C.c_int32), ("ngpu", C.c_int32),
                ("devices", C.POINTER(C.c_int32)), ("clonotypes", C.c_int32), ("reserved_", C.c_int32)]


ALN_DTYPE = np.dtype([("ordinal", "<i4"), ("rid", "<i4"), ("strand", "<i2"), ("mapped", "<i2"), ("as", "<i2"), ("qb", "<i2"),
                      ("m_end", "<i2"), ("rb", "<i2"), ("qe", "<i2"), ("re", "<i2"), ("nm", "<i2"), ("mi", "<i2"),
                      ("ncand", "<i2"), ("ntied", "<i2")])
assert ALN_DTYPE.itemsize == 32


class Read(C.Structure):
    _fields_ = [("seq", C.c_char_p), ("qual", C.c_char_p), ("seq_len", C.c_int32), ("qual_len", C.c_int32), ("v_id", C.c_int32),
                ("j_id", C.c_int32), ("cdr3_off", C.c_int32), ("cdr3_len", C.c_int32), ("ordinal", C.c_int32),
                ("able", C.c_uint8), ("flags", C.c_uint8), ("pad_", C.c_uint8 * 2)]


class Clonotype(C.Structure):
    _fields_ = [("cdr3", C.c_char_p), ("qual", C.c_char_p), ("cdr3_len", C.c_int32), ("v_id", C.c_int32), ("j_id", C.c_int32),
                ("reserved_", C.c_int32), ("count", C.c_int64)]


class Info(C.Structure):
    _fields_ = [("kmer", C.c_int32), ("reserved_", C.c_int32), ("the_err", C.c_double), ("avg_len", C.c_double)] + \
               [(k, C.c_int64) for k in ("n_reads", "v_hits", "j_hits", "v_final", "j_final", "share", "pair_seq_mismatch",
                                         "full_pushed", "vpart", "jpart", "direct", "left", "rescued", "cand_v", "cand_j",
                                         "cells_v", "cells_j", "gapped_v", "gapped_j", "kernel_launches")] + \
               [(k, C.c_double) for k in ("ms_seed", "ms_sw", "ms_select", "ms_extract", "ms_pool", "ms_h2d", "ms_d2h", "ms_host")] + \
               [(k, C.c_int64) for k in ("cells_computed_v", "cells_computed_j")] + \
               [(k, C.c_double) for k in ("ms_order", "ms_pack", "ms_seed_kernel", "ms_pool_kernel", "ms_strings")] + \
               [(k, C.c_int64) for k in ("bytes_h2d", "bytes_d2h")] + [("ms_j_stream", C.c_double)] + \
               [("n_ranks", C.c_int32), ("reserved2_", C.c_int32), ("ms_comm", C.c_double), ("bytes_comm", C.c_int64),
                ("ms_coll", C.c_double * 8), ("bytes_coll", C.c_int64 * 8), ("ms_stage_a_max", C.c_double), ("ms_stage_a_min", C.c_double)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}
        d["ms_coll"] = list(self.ms_coll)
        d["bytes_coll"] = list(self.bytes_coll)
        return d


COLL_NAMES = ("records_allgather", "mate_reads_allreduce", "item_flags_allreduce", "kmer_alltoall", "kmer_table_allgather",
              "string_sizes_allreduce", "result_reduce")      # CATT_COLL_* of include/catt_b200.h


EXPORTS = ["catt_ctx_create", "catt_ctx_destroy", "catt_last_error", "catt_version", "catt_ref_count", "catt_ref_name",
           "catt_ref_seq", "catt_run_file", "catt_run_reads", "catt_result_info", "catt_result_reads", "catt_align_records",
           "catt_result_free", "catt_best_d", "catt_align_shard", "catt_batch_upload", "catt_batch_align", "catt_batch_sync",
           "catt_batch_download", "catt_batch_free", "catt_ctx_stream", "catt_assemble", "catt_seed_candidates", "catt_sw_pairs",
           "catt_finder", "catt_alu_peak", "catt_batch_assemble", "catt_synth_reads", "catt_run_files", "catt_run_reads_files", "catt_set_host_threads", "catt_records_device", "catt_assemble_device", "catt_shard_counts", "catt_reads_device",
           "catt_assemble_resident", "catt_link_leaves", "catt_near_any", "catt_run_bam", "catt_bam_to_fastq", "catt_inflate_file",
           "catt_comm_id", "catt_ctx_join", "catt_run_reads_shard", "catt_shard_upload", "catt_shard_run", "catt_result_digest", "catt_result_text_digest", "catt_run_file_shard", "catt_result_clonotypes", "catt_real_score", "catt_order_records"]

_lib = None


def load():
    """dlopen the in-tree library.  Raises if it has not been built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} is missing: build it with `make -C catt_b200/csrc` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
    L.catt_last_error.restype = C.c_char_p
    L.catt_version.restype = C.c_char_p
    L.catt_ctx_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
    L.catt_ctx_destroy.argtypes = [vp]
    L.catt_ctx_destroy.restype = None
    L.catt_ref_count.argtypes = [vp, C.c_int]
    L.catt_ref_name.argtypes = [vp, C.c_int, C.c_int32]
    L.catt_ref_name.restype = C.c_char_p
    L.catt_ref_seq.argtypes = [vp, C.c_int, C.c_int32, C.c_char_p, C.c_int32]
    L.catt_run_file.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.catt_run_reads.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.catt_set_host_threads.argtypes = [C.c_int32]
    L.catt_records_device.argtypes = [vp, C.c_int, C.c_int64, C.POINTER(C.c_uint64)]
    L.catt_assemble_device.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.catt_shard_counts.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    L.catt_reads_device.argtypes = [vp, C.c_int64, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.catt_run_bam.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.catt_bam_to_fastq.argtypes = [C.c_char_p, C.c_char_p]
    L.catt_inflate_file.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(C.c_int32)]
    L.catt_link_leaves.argtypes = [vp, C.c_int32, vp, C.c_int64, vp, vp, vp, C.c_int64, vp, vp]
    L.catt_near_any.argtypes = [vp, C.c_int32, vp, C.c_int64, vp, C.c_int64, C.c_int32, C.c_int32, vp]
    L.catt_assemble_resident.argtypes = [vp, C.c_int64, C.c_int32, vp, vp, vp, vp, C.POINTER(vp)]
    L.catt_run_files.argtypes = [vp, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(vp)]
    L.catt_run_reads_files.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, C.c_int32, vp, C.POINTER(vp)]
    L.catt_result_info.argtypes = [vp, C.POINTER(Info)]
    L.catt_result_reads.argtypes = [vp, C.c_int, C.POINTER(C.POINTER(Read)), i64p]
    L.catt_align_records.argtypes = [vp, C.c_int, C.POINTER(vp), i64p]
    L.catt_result_free.argtypes = [vp]
    L.catt_result_free.restype = None
    L.catt_best_d.argtypes = [vp, C.POINTER(C.c_char_p), C.c_int64, vp]
    L.catt_align_shard.argtypes = [vp, C.c_int64, vp, vp, C.c_int64, vp, vp]
    L.catt_batch_upload.argtypes = [vp, C.c_int64, vp, vp, C.c_int64, C.POINTER(vp)]
    L.catt_batch_align.argtypes = [vp, vp]
    L.catt_batch_sync.argtypes = [vp, vp, C.POINTER(Info)]
    L.catt_batch_download.argtypes = [vp, vp, vp, vp]
    L.catt_batch_free.argtypes = [vp]
    L.catt_batch_free.restype = None
    L.catt_ctx_stream.argtypes = [vp]
    L.catt_ctx_stream.restype = C.c_uint64
    L.catt_assemble.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.catt_seed_candidates.argtypes = [vp, C.c_int, C.c_int64, vp, vp, vp]
    L.catt_sw_pairs.argtypes = [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp]
    L.catt_finder.argtypes = [vp, C.c_int, C.c_int64, vp, vp, vp]
    L.catt_batch_assemble.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.catt_synth_reads.argtypes = [vp, C.c_int64, C.c_int64, C.c_int32, C.c_uint64, C.c_double, C.c_int32, vp, vp, vp, vp]
    L.catt_alu_peak.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.catt_shard_upload.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, C.c_int32, vp]
    L.catt_shard_run.argtypes = [vp, C.POINTER(vp)]
    L.catt_run_file_shard.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.catt_result_digest.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.catt_result_text_digest.argtypes = [vp, vp, C.POINTER(C.c_uint64)]
    L.catt_result_clonotypes.argtypes = [vp, C.POINTER(C.POINTER(Clonotype)), i64p]
    L.catt_real_score.argtypes = [vp, C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp, vp, C.c_int32, vp]
    L.catt_order_records.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp, i64p, vp]
    L.catt_comm_id.argtypes = [vp]
    L.catt_ctx_join.argtypes = [vp, C.c_int32, C.c_int32, vp]
    L.catt_run_reads_shard.argtypes = [vp, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp, vp, vp, C.c_int32, vp, C.POINTER(vp)]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise CattError(rc, load().catt_last_error().decode(errors="replace"))


def concat(strings):
    """list[str|bytes] -> (bytes buffer, int64 offsets[n+1]) in the layout catt_run_reads expects."""
    bs = [s.encode() if isinstance(s, str) else s for s in strings]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        np.cumsum([len(b) for b in bs], out=off[1:])
    return b"".join(bs), off
