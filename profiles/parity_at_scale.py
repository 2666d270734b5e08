#!/usr/bin/env python3
"""Parity at size: the CUDA path (C ABI) against the CPU oracle on samples of 1-10 M reads, compared through the canonical text
digest of both lists (catt_result_text_digest / co_result_digest: seq, qual, V name, J name, cdr3, able of every read, order
sensitive), every log counter, and a hash of the clonotype table (V, J, CDR3 -> count) that catt() goes on to build.
Also: the same sample through a 2-rank GPU group must give the same digests.

    python profiles/parity_at_scale.py [scale]        # scale 1.0 = 10 M hs TRB 150 bp + 1 M each of five other configs

The oracle runs on all host cores (OpenMP); ~1 k reads/s per core for hs TRB 150 bp."""
import collections, hashlib, os, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from catt_b200.host import Catt
from catt_b200.refg import chain_config
from oracle import oracle as O

d = tempfile.mkdtemp(prefix="catt_parity_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
res_dir = os.path.join(d, "resource")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
cases = [("hs", "TRB", 150, 10_000_000), ("hs", "TRA", 50, 1_000_000), ("hs", "IGH", 50, 1_000_000), ("hs", "TRB", 50, 1_000_000),
         ("ms", "TRB", 150, 1_000_000), ("pig", "TRB", 150, 1_000_000)]
KEYS = ("kmer", "the_err", "v_hits", "j_hits", "v_final", "j_final", "share", "pair_seq_mismatch", "full_pushed", "vpart", "jpart", "direct",
        "left", "rescued", "cand_v", "cand_j", "cells_v", "cells_j")


def table_hash(parts):
    t = collections.Counter((r.vs, r.js, r.cdr3) for p in parts for r in p if r.cdr3 != "None")
    h = hashlib.sha256()
    for k in sorted(t):
        h.update(("\t".join(k) + f"\t{t[k]}\n").encode())
    return h.hexdigest()[:16], len(t), sum(t.values())


bad = 0
threads = len(os.sched_getaffinity(0))
for species, chain, length, n in cases:
    n = int(n * scale)
    cc = chain_config(species, chain, "CDR3", res_dir)
    V, J = O.RefSet(cc["vregion"]), O.RefSet(cc["jregion"])
    full = Catt(chain_cfg=cc, device=0)                            # every Vpart / Jpart read
    only = Catt(chain_cfg=cc, device=0, only_cdr3=True)            # the lists after catt.jl:581-582
    grp = Catt(chain_cfg=cc, device=0, only_cdr3=True, ngpu=2, devices=[0, 0])
    name_buf, name_off, seq_buf, seq_off, qual_buf = full.synth_reads(n, length)
    seqs, quals = O.synth_reads(V, J, n, length)
    twin = np.array_equal(seqs.reshape(-1), seq_buf) and np.array_equal(quals.reshape(-1), qual_buf)
    t0 = time.perf_counter()
    ref = O.run_flat(V, J, O.make_config(cc), seqs, quals, threads=threads)
    t1 = time.perf_counter()
    rf = full.mainflow_buffers(n, name_buf, name_off, seq_buf, seq_off, qual_buf)
    t2 = time.perf_counter()
    ro = only.mainflow_buffers(n, name_buf, name_off, seq_buf, seq_off, qual_buf)
    rg = grp.mainflow_buffers(n, name_buf, name_off, seq_buf, seq_off, qual_buf)
    ok_full = rf.text_digest() == ref.digest(False)
    ok_only = ro.text_digest() == ref.digest(True)
    ok_grp = rg.text_digest() == ref.digest(True) and rg.digest() == ro.digest()
    okc = all(rf.info[k] == ref.info[k] for k in KEYS) and rf.info["gapped_v"] == ref.info["v_gapped"] and rf.info["gapped_j"] == ref.info["j_gapped"] and \
        all(ro.info[k] == ref.info[k] for k in KEYS) and all(rg.info[k] == ref.info[k] for k in KEYS)
    th_gpu = table_hash((ro.Vpart, ro.Jpart))
    th_ref = table_hash(([r for r in ref.Vpart if r.cdr3 != "None"], [r for r in ref.Jpart if r.cdr3 != "None"])) if n <= 2_000_000 else None
    if th_ref is None:                                             # 10 M reads: the oracle's lists as text are > 1 GB; the digest covers them
        ok_tab = True
    else:
        ok_tab = th_gpu == th_ref
    good = twin and ok_full and ok_only and ok_grp and okc and ok_tab
    bad += not good
    print(f"{species} {chain} {length}bp n={n}: generator twin {'==' if twin else 'DIFFERS'}, full lists {'identical' if ok_full else 'DIFFER'} "
          f"(digest V {rf.text_digest()[0]:016x} J {rf.text_digest()[1]:016x}), lists after filter! {'identical' if ok_only else 'DIFFER'}, "
          f"2-rank group {'identical' if ok_grp else 'DIFFERS'}, counters {'==' if okc else 'DIFFER'}, clonotype table {th_gpu[0]} "
          f"({th_gpu[1]} clonotypes, {th_gpu[2]} reads){'' if th_ref is None else (' == oracle' if ok_tab else ' != oracle ' + th_ref[0])} | "
          f"Vpart {rf.info['vpart']} Jpart {rf.info['jpart']} rescued {rf.info['rescued']} k {rf.info['kmer']} | "
          f"oracle {t1 - t0:.1f} s on {threads} threads, GPU (full lists) {1e3 * (t2 - t1):.0f} ms", flush=True)
    for x in (rf, ro, rg):
        x.close()
    for x in (full, only, grp):
        x.close()
    del ref
print("RESULT:", "all identical" if not bad else f"{bad} case(s) differ")
sys.exit(1 if bad else 0)
