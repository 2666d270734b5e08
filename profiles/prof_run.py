#!/usr/bin/env python3
"""Small driver for ncu captures: N synthetic hs-TRB reads through catt_run_reads, `reps` times (profiles/README.md)."""
import os, sys, tarfile, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from catt_b200.host import Catt
from catt_b200.refg import chain_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
length = int(sys.argv[3]) if len(sys.argv) > 3 else 150
d = tempfile.mkdtemp(prefix="catt_prof_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
ctx = Catt(chain_cfg=chain_config("hs", "TRB", "CDR3", os.path.join(d, "resource")))
bufs = ctx.synth_reads(n, length)
for _ in range(reps):
    res = ctx.mainflow_buffers(n, *bufs)
    info = res.info
    res.close()
print({k: info[k] for k in ("n_reads", "v_hits", "j_hits", "vpart", "jpart", "ms_seed", "ms_sw", "ms_select", "ms_order", "ms_extract", "ms_pool", "kernel_launches")})
