#!/usr/bin/env python3
"""catt_run_file on a synthetic FASTQ written to local disk: where does the time go (CATT_TRACE=1)."""
import os, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from catt_b200.host import Catt
from catt_b200.refg import chain_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
d = tempfile.mkdtemp(prefix="catt_file_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
ctx = Catt(chain_cfg=chain_config("hs", "TRB", "CDR3", os.path.join(d, "resource")), only_cdr3="cdr3" in sys.argv)   # "cdr3": the production result mode
names, name_off, seqs, seq_off, quals = ctx.synth_reads(n, 150)
path = os.path.join(d, "synth.fq")
L = 150
rec = np.empty((n, 1 + 10 + 1 + L + 3 + L + 1), dtype=np.uint8)
rec[:, 0] = ord("@")
nm = np.char.zfill(np.arange(n).astype(str), 9)
rec[:, 1] = ord("s")
rec[:, 2:11] = np.frombuffer("".join(nm.tolist()).encode(), dtype=np.uint8).reshape(n, 9)
rec[:, 11] = 10
rec[:, 12:12 + L] = seqs.reshape(n, L)
rec[:, 12 + L] = 10; rec[:, 13 + L] = ord("+"); rec[:, 14 + L] = 10
rec[:, 15 + L:15 + 2 * L] = quals.reshape(n, L)
rec[:, 15 + 2 * L] = 10
rec.tofile(path)
print("file bytes", os.path.getsize(path))
for rep in range(int(os.environ.get('FILE_RUN_REPS', '3'))):
    t0 = time.perf_counter()
    res = ctx.mainflow(path)
    dt = time.perf_counter() - t0
    print(f"catt_run_file: {dt * 1e3:.1f} ms  ({n / dt / 1e6:.2f} M reads/s)", {k: res.info[k] for k in ("vpart", "jpart")}, flush=True)
    res.close()
