#!/usr/bin/env python3
"""Times the two nbsorb scans through the C ABI with host buffers, beside the CPU oracle on a bounded sample.

    python profiles/nbsorb_bench.py [--quick]

near_any: 1 M low-count sequences x 10 k high-confidence ones, lgt 45 (random: no early exit -> every pair is compared).
link    : 200 k leaves x 2 k roots, lgt 45.
"""
import os, sys, tarfile, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from catt_b200.host import Catt
from catt_b200.refg import chain_config
from oracle import oracle as O

quick = "--quick" in sys.argv
d = tempfile.mkdtemp(prefix="catt_nb_")
with tarfile.open(os.path.join(ROOT, "tests", "golden", "catt_fixture.tar.gz")) as tar:
    tar.extractall(d)
ctx = Catt(chain_cfg=chain_config("hs", "TRB", "CDR3", os.path.join(d, "resource")), device=0)
rng = np.random.default_rng(1)
lut = np.frombuffer(b"ACGT", dtype=np.uint8)
for lgt in (45, 105):
    nq, nt = (100_000, 2_000) if quick else (1_000_000, 10_000)
    q = lut[rng.integers(0, 4, size=(nq, lgt))]
    t = lut[rng.integers(0, 4, size=(nt, lgt))]
    for _ in range(2):
        ctx.near_any(lgt, q, t)
    t0 = time.perf_counter(); reps = 3
    for _ in range(reps):
        flag = ctx.near_any(lgt, q, t)
    dt = (time.perf_counter() - t0) / reps
    print(f"near_any lgt {lgt}: {nq} x {nt} = {nq * nt / 1e9:.1f} G pairs in {dt * 1e3:.1f} ms (host buffers in/out) = {nq * nt / dt / 1e12:.2f} T pairs/s, "
          f"{nq * nt * (lgt - 6) / dt / 1e12:.1f} T base compares/s; flagged {int(flag.sum())}", flush=True)
    sq = 2000
    qs = ["".join(map(chr, r)) for r in q[:sq]]; ts = ["".join(map(chr, r)) for r in t]
    t0 = time.perf_counter(); ref = O.near_any(lgt, qs, ts); dc = time.perf_counter() - t0
    assert np.array_equal(ref, flag[:sq])
    print(f"   CPU oracle ({os.cpu_count()} threads, {sq} x {nt} sample): {sq * nt / dc / 1e9:.3f} G pairs/s -> GPU/CPU {nq * nt / dt / (sq * nt / dc):.0f}x", flush=True)

    nl, nr = (20_000, 500) if quick else (200_000, 2_000)
    leaves = lut[rng.integers(0, 4, size=(nl, lgt))]
    roots = lut[rng.integers(0, 4, size=(nr, lgt))]
    src = rng.integers(0, nr, size=nl)
    rel = rng.random(nl) < 0.5
    leaves[rel] = roots[src[rel]]
    pos = rng.integers(3, lgt - 3, size=nl); leaves[np.arange(nl), pos] = lut[rng.integers(0, 4, size=nl)]
    up = np.tile(np.array([40.0, 12.0, 4.0, 1.5, 1.0, 1.0, 1.0]), (nr, 1))
    cnt = rng.integers(1, 50, size=nl)
    for _ in range(2):
        ctx.link_leaves(lgt, roots, up, leaves, cnt)
    t0 = time.perf_counter()
    for _ in range(reps):
        st, hd = ctx.link_leaves(lgt, roots, up, leaves, cnt)
    dt = (time.perf_counter() - t0) / reps
    print(f"link_leaves lgt {lgt}: {nl} x {nr} in {dt * 1e3:.1f} ms; statu histogram {np.bincount(st, minlength=3).tolist()}", flush=True)
    sl = 2000
    est, ehd = O.link_leaves(lgt, ["".join(map(chr, r)) for r in roots], up, ["".join(map(chr, r)) for r in leaves[:sl]], cnt[:sl])
    assert np.array_equal(est, st[:sl]) and np.array_equal(ehd, hd[:sl])
ctx.close()
