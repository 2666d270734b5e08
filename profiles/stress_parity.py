#!/usr/bin/env python3
"""Large randomized parity run: CUDA path (C ABI) vs the CPU oracle on synthetic samples big enough to hit the rare paths
(seeding run-list overflow -> global scratch pass, SW task-buffer overflow -> re-run, gapped winners -> traceback kernel,
QNAME hash collisions, multi-chunk Stage A, paired files).  Prints one line per case; exits non-zero on any difference."""
import os, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from catt_b200 import _lib
from catt_b200.host import Catt
from catt_b200.refg import chain_config
from oracle import oracle as O

d = tempfile.mkdtemp(prefix="catt_stress_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
res_dir = os.path.join(d, "resource")
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
cases = [("hs", "TRB", 150, 300000, 1, True, 0), ("hs", "TRB", 100, 200000, 4, True, 0), ("hs", "TRB", 50, 200000, 3, False, 0),
         ("hs", "IGH", 150, 120000, 1, True, 0), ("hs", "TRA", 100, 120000, 2, True, 0), ("ms", "TRB", 150, 100000, 1, False, 0),
         ("pig", "TRB", 150, 100000, 1, False, 0), ("hs", "TRG", 150, 60000, 1, True, 0), ("hs", "TRB", 150, 150000, 2, True, 65536)]
bad = 0
for species, chain, length, n, threads, with_n, chunk in cases:
    n = int(n * scale)
    cc = chain_config(species, chain, "CDR3", res_dir)
    V, J = O.RefSet(cc["vregion"]), O.RefSet(cc["jregion"])
    seqs, quals = O.synth_reads(V, J, n, length, seed=90210 + length, with_n=with_n)
    # mutate a slice of reads with indels so the traceback kernel gets work
    rng = np.random.default_rng(5)
    for i in rng.integers(0, n, size=n // 50):
        p = int(rng.integers(10, length - 10))
        seqs[i, p:-1] = seqs[i, p + 1:].copy()
    names = [f"q{i}" if i % 7 else f"q{i // 2}/{1 + i % 2}" for i in range(n)]        # some mate-style names that collapse
    sl = ["".join(map(chr, row)) for row in seqs]
    ql = ["".join(map(chr, row)) for row in quals]
    t0 = time.perf_counter()
    ref = O.run_mem(V, J, O.make_config(cc, nthreads=threads), names, sl, ql)
    t1 = time.perf_counter()
    c = Catt(chain_cfg=cc, thread=threads, chunk_reads=chunk)
    nb, no = _lib.concat(names); sb, so = _lib.concat(sl); qb, _ = _lib.concat(ql)
    res = c.mainflow_buffers(n, nb, no, sb, so, qb)
    t2 = time.perf_counter()
    tup = lambda r: (r.seq, r.qual, r.vs, r.js, r.cdr3)
    ok = [tup(r) for r in res.Vpart] == [tup(r) for r in ref.Vpart] and [tup(r) for r in res.Jpart] == [tup(r) for r in ref.Jpart]
    keys = ("kmer", "the_err", "v_hits", "j_hits", "v_final", "j_final", "share", "pair_seq_mismatch", "full_pushed", "vpart", "jpart", "direct",
            "left", "rescued", "cand_v", "cand_j", "cells_v", "cells_j")
    okc = all(res.info[k] == ref.info[k] for k in keys) and res.info["gapped_v"] == ref.info["v_gapped"] and res.info["gapped_j"] == ref.info["j_gapped"]
    bad += not (ok and okc)
    print(f"{species} {chain} {length}bp n={n} threads={threads} chunk={chunk or 'default'}: lists {'==' if ok else 'DIFFER'}, counters {'==' if okc else 'DIFFER'} | "
          f"Vpart {res.info['vpart']} Jpart {res.info['jpart']} gapped_v {res.info['gapped_v']} rescued {res.info['rescued']} | oracle {t1 - t0:.1f} s, GPU {1e3 * (t2 - t1):.0f} ms", flush=True)
    c.close()
sys.exit(1 if bad else 0)
