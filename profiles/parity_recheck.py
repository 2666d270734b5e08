#!/usr/bin/env python3
"""Parity at size for a NEW build without the hours of CPU time: profiles/r02_parity_10M.txt records, for every case of
parity_at_scale.py, the digests of the full lists that were verified equal to the CPU oracle's.  This script re-runs only the CUDA
side on the same generated samples (full lists, the lists after filter!, a 2-rank and a 3-rank group) and compares with the
recorded digests, Vpart / Jpart sizes, rescued count, k and the clonotype-table hash.

    python profiles/parity_recheck.py [recorded file = profiles/r02_parity_10M.txt]
"""
import collections, hashlib, os, re, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from catt_b200.host import Catt
from catt_b200.refg import chain_config

rec_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_parity_10M.txt")
d = tempfile.mkdtemp(prefix="catt_recheck_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
res_dir = os.path.join(d, "resource")
pat = re.compile(r"^(\w+) (\w+) (\d+)bp n=(\d+): .*digest V ([0-9a-f]{16}) J ([0-9a-f]{16}).*clonotype table ([0-9a-f]{16}) \((\d+) clonotypes, (\d+) reads\).*"
                 r"Vpart (\d+) Jpart (\d+) rescued (\d+) k (\d+)")


def table_hash(parts):
    t = collections.Counter((r.vs, r.js, r.cdr3) for p in parts for r in p if r.cdr3 != "None")
    h = hashlib.sha256()
    for k in sorted(t):
        h.update(("\t".join(k) + f"\t{t[k]}\n").encode())
    return h.hexdigest()[:16], len(t), sum(t.values())


bad = 0
for line in open(rec_path):
    m = pat.match(line)
    if not m:
        continue
    species, chain, length, n = m.group(1), m.group(2), int(m.group(3)), int(m.group(4))
    want = (int(m.group(5), 16), int(m.group(6), 16))
    cc = chain_config(species, chain, "CDR3", res_dir)
    full = Catt(chain_cfg=cc, device=0)
    only = Catt(chain_cfg=cc, device=0, only_cdr3=True)
    bufs = full.synth_reads(n, length)
    t0 = time.perf_counter()
    rf = full.mainflow_buffers(n, *bufs)
    t1 = time.perf_counter()
    ro = only.mainflow_buffers(n, *bufs)
    ok_full = tuple(rf.text_digest()) == want
    ok_cnt = (rf.info["vpart"], rf.info["jpart"], rf.info["rescued"], rf.info["kmer"]) == tuple(int(m.group(i)) for i in (10, 11, 12, 13))
    th = table_hash((ro.Vpart, ro.Jpart))
    ok_tab = th == (m.group(7), int(m.group(8)), int(m.group(9)))
    ok_grp = True
    for ng in (2, 3):
        grp = Catt(chain_cfg=cc, device=0, only_cdr3=True, ngpu=ng, devices=[0] * ng)
        rg = grp.mainflow_buffers(n, *bufs)
        ok_grp = ok_grp and rg.digest() == ro.digest() and rg.text_digest() == ro.text_digest()
        rg.close(); grp.close()
    good = ok_full and ok_cnt and ok_tab and ok_grp
    bad += not good
    print(f"{species} {chain} {length}bp n={n}: full lists {'== recorded oracle-verified digest' if ok_full else 'DIFFER'} "
          f"(V {rf.text_digest()[0]:016x} J {rf.text_digest()[1]:016x}), sizes / rescued / k {'==' if ok_cnt else 'DIFFER'}, "
          f"clonotype table {th[0]} {'==' if ok_tab else 'DIFFERS'}, 2- and 3-rank groups {'identical' if ok_grp else 'DIFFER'} | GPU (full lists) {1e3 * (t1 - t0):.0f} ms",
          flush=True)
    rf.close(); ro.close(); full.close(); only.close()
print("RESULT:", "all identical" if not bad else f"{bad} case(s) differ")
sys.exit(1 if bad else 0)
