#!/usr/bin/env python3
"""Where the end-to-end call (catt_run_reads, host buffers) loses time against the resident path (catt_shard_run) on ONE GPU.

    python profiles/e2e_gap.py [reads=20000000]

Prints wall time and the library's stage times of both paths (environment switches such as CATT_CHUNK_READS apply)."""
import os, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from catt_b200.host import Catt
from catt_b200.refg import chain_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
d = tempfile.mkdtemp(prefix="catt_gap_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
cc = chain_config("hs", "TRB", "CDR3", os.path.join(d, "resource"))
ctx = Catt(chain_cfg=cc, device=0, only_cdr3=True)
bufs = ctx.synth_reads(n, 150)
KEYS = ("ms_seed", "ms_sw", "ms_select", "ms_j_stream", "ms_pack", "ms_order", "ms_extract", "ms_pool", "ms_strings", "ms_d2h", "ms_host")


def run(label, f, reps=3):
    f().close()
    best, info = 1e30, None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = f()
        dt = time.perf_counter() - t0
        if dt < best:
            best, info = dt, r.info
        r.close()
    print(f"{label}: {best * 1e3:8.1f} ms ({n / best / 1e6:.2f} M reads/s)  " + "  ".join(f"{k[3:]} {info[k]:.0f}" for k in KEYS), flush=True)
    return best


ctx.shard_upload(n, 0, n, *bufs)
a = run("resident", ctx.shard_run)
b = run("e2e     ", lambda: ctx.mainflow_buffers(n, *bufs))
print(f"gap {1e3 * (b - a):.1f} ms = {100 * (b - a) / a:.1f} % (chunk_reads {os.environ.get('CATT_CHUNK_READS', 'default')})")
