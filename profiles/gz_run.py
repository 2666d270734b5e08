#!/usr/bin/env python3
"""`.fastq.gz` input at scale: N synthetic hs-TRB reads as plain FASTQ, single-member gzip and BGZF through
catt_run_file, and the inflate alone (catt_inflate_file) against one zlib stream.

    python profiles/gz_run.py [reads] [gzip level]
"""
import os, sys, tarfile, tempfile, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from catt_b200.host import Catt, inflate_file
from catt_b200.refg import chain_config
from oracle import bamio

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
L = 150
d = tempfile.mkdtemp(prefix="catt_gz_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
ctx = Catt(chain_cfg=chain_config("hs", "TRB", "CDR3", os.path.join(d, "resource")), device=0, only_cdr3=True)
name_buf, name_off, seq_buf, seq_off, qual_buf = ctx.synth_reads(n, L)
w = 10
rec = np.empty((n, 1 + 1 + w + 1 + L + 3 + L + 1), dtype=np.uint8)
rec[:, 0] = ord("@"); rec[:, 1] = ord("s")
rec[:, 2:2 + w] = np.frombuffer("".join(np.char.zfill(np.arange(n).astype(str), w).tolist()).encode(), dtype=np.uint8).reshape(n, w)
o = 2 + w
rec[:, o] = 10
rec[:, o + 1:o + 1 + L] = np.frombuffer(seq_buf, dtype=np.uint8).reshape(n, L)
o += 1 + L
rec[:, o] = 10; rec[:, o + 1] = ord("+"); rec[:, o + 2] = 10
rec[:, o + 3:o + 3 + L] = np.frombuffer(qual_buf, dtype=np.uint8).reshape(n, L)
rec[:, o + 3 + L] = 10
text = rec.tobytes()
plain, gz, bg = (os.path.join(d, x) for x in ("s.fq", "s.fq.gz", "s.bgzf.fq.gz"))
open(plain, "wb").write(text)
t0 = time.perf_counter()
co = zlib.compressobj(level, zlib.DEFLATED, 31)
with open(gz, "wb") as fh:
    for i in range(0, len(text), 1 << 24):
        fh.write(co.compress(text[i:i + (1 << 24)]))
    fh.write(co.flush())
with open(bg, "wb") as fh:
    for i in range(0, len(text), 65280):
        fh.write(bamio.bgzf_block(text[i:i + 65280], level))
    fh.write(bamio.EOF_BLOCK)
print(f"{n} reads: {len(text) / 1e6:.0f} MB FASTQ, {os.path.getsize(gz) / 1e6:.0f} MB .gz (level {level}), {os.path.getsize(bg) / 1e6:.0f} MB BGZF "
      f"({time.perf_counter() - t0:.0f} s to compress)", flush=True)
out = os.path.join(d, "out.bin")
for label, env in (("parallel gzip (pgz)", {}), ("one zlib stream", {"CATT_ZLIB_ONLY": "1"})):
    os.environ.update(env)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); m = inflate_file(gz, out); ts.append(time.perf_counter() - t0)
    for k in env:
        os.environ.pop(k)
    dt = min(ts)
    print(f"inflate .gz, {label}: method {m}, {dt * 1e3:.0f} ms = {len(text) / dt / 1e9:.2f} GB/s of text (includes writing {len(text) / 1e6:.0f} MB to the page cache)", flush=True)
ts = []
for _ in range(3):
    t0 = time.perf_counter(); m = inflate_file(bg, out); ts.append(time.perf_counter() - t0)
print(f"inflate BGZF: method {m}, {min(ts) * 1e3:.0f} ms = {len(text) / min(ts) / 1e9:.2f} GB/s of text", flush=True)

def timed(path):
    ctx.mainflow(path).close()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r = ctx.mainflow(path); ts.append(time.perf_counter() - t0); info = r.info; r.close()
    return sorted(ts)[1], info

base = None
for label, path, env in (("plain FASTQ", plain, {}), (".fastq.gz, parallel inflate", gz, {}), (".fastq.gz, one zlib stream", gz, {"CATT_ZLIB_ONLY": "1"}),
                         ("BGZF .fastq.gz", bg, {})):
    os.environ.update(env)
    dt, info = timed(path)
    for k in env:
        os.environ.pop(k)
    key = {k: info[k] for k in ("vpart", "jpart", "direct", "left", "rescued", "v_hits", "j_hits", "kmer")}
    base = base or key
    assert key == base, (key, base)
    print(f"catt_run_file, {label:30s}: {dt * 1e3:7.0f} ms = {n / dt / 1e6:6.2f} M reads/s", flush=True)
ctx.close()
