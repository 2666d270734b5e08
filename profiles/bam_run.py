#!/usr/bin/env python3
"""`--bam` input at scale: a BAM of N synthetic hs-TRB reads (SURVEY 8d config 3 generator; half of the records stored reverse-strand)
through catt_run_bam, beside catt_run_file on the FASTQ that catt_bam_to_fastq writes from it.

    python profiles/bam_run.py [reads]
"""
import os, struct, sys, tarfile, tempfile, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from catt_b200.host import Catt, bam_to_fastq
from catt_b200.refg import chain_config
from oracle import bamio

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
L = 150
d = tempfile.mkdtemp(prefix="catt_bam_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
ctx = Catt(chain_cfg=chain_config("hs", "TRB", "CDR3", os.path.join(d, "resource")), device=0, only_cdr3=True)
name_buf, name_off, seq_buf, seq_off, qual_buf = ctx.synth_reads(n, L)
seq = np.frombuffer(seq_buf, dtype=np.uint8).reshape(n, L)
qual = np.frombuffer(qual_buf, dtype=np.uint8).reshape(n, L)
# vectorised BAM records: fixed-width names "s%09d", flag 0 / 16 alternating (reverse-strand records are stored reverse-complemented)
code = np.full(256, 15, dtype=np.uint8)
for i, c in enumerate(bamio.NT16):
    code[ord(c)] = i
comp = np.array([0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15], dtype=np.uint8)
c4 = code[seq]
rev = (np.arange(n) & 1).astype(bool)
c4[rev] = comp[c4[rev, ::-1]]
q = (qual - 33).astype(np.uint8)
q[rev] = q[rev, ::-1]
names = np.frombuffer("".join(f"s{i:09d}\0" for i in range(n)).encode(), dtype=np.uint8).reshape(n, 11)
body = 32 + 11 + L // 2 + L
rec = np.zeros((n, 4 + body), dtype=np.uint8)
rec[:, 0:4] = np.frombuffer(struct.pack("<I", body), dtype=np.uint8)
fixed = bytearray(struct.pack("<iiBBHHHiiii", -1, -1, 11, 0, 4680, 0, 4, L, -1, -1, 0))
rec[:, 4:36] = np.frombuffer(bytes(fixed), dtype=np.uint8)
rec[rev, 4 + 14] = 4 | 16
rec[:, 36:47] = names
rec[:, 47:47 + L // 2] = (c4[:, 0::2] << 4) | c4[:, 1::2]
rec[:, 47 + L // 2:] = q
text = b"@HD\tVN:1.6\tSO:unsorted\n"
payload = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", 0) + rec.tobytes()
path = os.path.join(d, "synth.bam")
t0 = time.perf_counter()
with open(path, "wb") as fh:
    for i in range(0, len(payload), 65280):
        fh.write(bamio.bgzf_block(payload[i:i + 65280], 1))
    fh.write(bamio.EOF_BLOCK)
print(f"BAM of {n} reads: {os.path.getsize(path) / 1e6:.0f} MB compressed, {len(payload) / 1e6:.0f} MB of records ({time.perf_counter() - t0:.1f} s to write)", flush=True)
fq = os.path.join(d, "synth.fq")
os.environ["CATT_TRACE"] = "1"
t0 = time.perf_counter(); bam_to_fastq(path, fq); dt = time.perf_counter() - t0
os.environ.pop("CATT_TRACE")
print(f"catt_bam_to_fastq: {dt * 1e3:.0f} ms = {n / dt / 1e6:.2f} M reads/s (host only, {os.cpu_count()} cores)", flush=True)

def timed(fn):
    fn().close()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t0); info = r.info; r.close()
    return sorted(ts)[1], info

tb, ib = timed(lambda: ctx.mainflow_bam(path))
tf, i_f = timed(lambda: ctx.mainflow(fq))
assert all(ib[k] == i_f[k] for k in ("vpart", "jpart", "direct", "left", "rescued", "v_hits", "j_hits", "kmer")), (ib, i_f)
print(f"catt_run_bam : {tb * 1e3:.0f} ms = {n / tb / 1e6:.2f} M reads/s", flush=True)
print(f"catt_run_file: {tf * 1e3:.0f} ms = {n / tf / 1e6:.2f} M reads/s on the FASTQ written from the BAM (same counters: Vpart {ib['vpart']}, Jpart {ib['jpart']})", flush=True)
ctx.close()
