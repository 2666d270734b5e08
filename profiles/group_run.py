#!/usr/bin/env python3
"""ONE process, N GPUs: what the Julia driver gets from `catt_config.ngpu` - the in-process GPU group behind the
unchanged calls catt_run_reads and catt_run_file, against the same calls on one GPU.

    python profiles/group_run.py [reads=100000000] [file_reads=16000000] [ngpu=all]

Prints the digests (must agree), wall times of 3 calls each, and the collective breakdown of the last group call."""
import os, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from catt_b200 import _lib
from catt_b200.host import Catt
from catt_b200.refg import chain_config

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 16_000_000
ngpu = int(sys.argv[3]) if len(sys.argv) > 3 else torch.cuda.device_count()
file_only = len(sys.argv) > 4 and sys.argv[4] == "file_only"
d = tempfile.mkdtemp(prefix="catt_group_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
cc = chain_config("hs", "TRB", "CDR3", os.path.join(d, "resource"))
one = Catt(chain_cfg=cc, device=0, only_cdr3=True, clonotypes=True)
ndev = torch.cuda.device_count()
# more ranks than GPUs: ranks share devices and exchange through peer copies (how the group path is exercised on a one-GPU box)
grp = Catt(chain_cfg=cc, device=0, only_cdr3=True, clonotypes=True, ngpu=ngpu, devices=[r % ndev for r in range(ngpu)] if ngpu > ndev else None)
length = 150


def timed(f, reps=3):
    f().close()
    ts, last = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = f()
        ts.append(time.perf_counter() - t0)
        if last is not None:
            last.close()
        last = r
    return sorted(ts)[len(ts) // 2], last


t0 = time.perf_counter()
bufs = one.synth_reads(n, length)
print(f"generated {n} reads in {time.perf_counter() - t0:.1f} s; host threads {len(os.sched_getaffinity(0))}, GPUs in the group {ngpu}", flush=True)
ok = True
if not file_only:
    t1, r1 = timed(lambda: one.mainflow_buffers(n, *bufs))
    tg, rg = timed(lambda: grp.mainflow_buffers(n, *bufs))
    same = r1.digest() == rg.digest() and r1.text_digest() == rg.text_digest() and r1.clonotypes()[:2000] == rg.clonotypes()[:2000] and len(r1.clonotypes()) == len(rg.clonotypes()) if n <= 20_000_000 else r1.digest() == rg.digest()
    i = rg.info
    print(f"catt_run_reads, {n} reads from host buffers: 1 GPU {t1 * 1e3:.0f} ms ({n / t1 / 1e6:.2f} M reads/s), {ngpu} GPUs in one process {tg * 1e3:.0f} ms "
          f"({n / tg / 1e6:.2f} M reads/s, {t1 / tg:.2f}x); results {'identical' if same else 'DIFFER'} (digest {rg.digest()[0]:016x})")
    print("  Stage A slowest / fastest rank %.0f / %.0f ms; order %.0f, extract %.0f, k-mer table + extension %.0f, collectives %.1f ms: %s" % (
        i["ms_stage_a_max"], i["ms_stage_a_min"], i["ms_order"], i["ms_extract"], i["ms_pool"], i["ms_comm"],
        ", ".join(f"{nm} {i['ms_coll'][k]:.1f} ms / {i['bytes_coll'][k] / 1e9:.2f} GB" for k, nm in enumerate(_lib.COLL_NAMES))), flush=True)
    r1.close(); rg.close()
    ok = same
# fresh contexts for the file part: the 100 M-read calls left ~100 GB of scratch on device 0 (both contexts live there)
one.close(); grp.close()
one = Catt(chain_cfg=cc, device=0, only_cdr3=True, clonotypes=True)
grp = Catt(chain_cfg=cc, device=0, only_cdr3=True, clonotypes=True, ngpu=ngpu, devices=[r % ndev for r in range(ngpu)] if ngpu > ndev else None)

if nf > 0:
    nf = min(nf, n)
    path = os.path.join(d, "sample.fq")
    w = 10
    t0 = time.perf_counter()
    with open(path, "wb") as fh:
        step = 2_000_000
        for a in range(0, nf, step):
            b = min(nf, a + step)
            m = b - a
            rec = np.empty((m, 1 + 1 + w + 1 + length + 3 + length + 1), dtype=np.uint8)
            rec[:, 0] = ord("@"); rec[:, 1] = ord("s")
            rec[:, 2:2 + w] = np.frombuffer("".join(np.char.zfill(np.arange(a, b).astype(str), w).tolist()).encode(), dtype=np.uint8).reshape(m, w)
            o = 2 + w
            rec[:, o] = 10
            rec[:, o + 1:o + 1 + length] = bufs[2][a * length:b * length].reshape(m, length)
            o += 1 + length
            rec[:, o] = 10; rec[:, o + 1] = ord("+"); rec[:, o + 2] = 10
            rec[:, o + 3:o + 3 + length] = bufs[4][a * length:b * length].reshape(m, length)
            rec[:, o + 3 + length] = 10
            rec.tofile(fh)
    print(f"wrote {os.path.getsize(path) / 1e9:.2f} GB FASTQ ({nf} reads) in {time.perf_counter() - t0:.1f} s", flush=True)
    t1, r1 = timed(lambda: one.mainflow(path))
    tg, rg = timed(lambda: grp.mainflow(path))
    same = r1.digest() == rg.digest() and r1.text_digest() == rg.text_digest()
    print(f"catt_run_file, {nf} reads in the page cache: 1 GPU {t1 * 1e3:.0f} ms ({nf / t1 / 1e6:.2f} M reads/s), {ngpu} GPUs in one process {tg * 1e3:.0f} ms "
          f"({nf / tg / 1e6:.2f} M reads/s, {t1 / tg:.2f}x); results {'identical' if same else 'DIFFER'}; ingest (read + H2D + index + pack) on the slowest rank "
          f"{rg.info['ms_pack']:.0f} ms", flush=True)
    ok = ok and same
    os.remove(path)
print("RESULT:", "identical" if ok else "DIFFERENT")
sys.exit(0 if ok else 1)
