"""The two Hamming scans of `nbsorb` (Jtool.jl:396-421 / 524-531 / 545-560).

CPU part: the C restatement in oracle/ against the line-by-line Python port of nbsorb that reproduces the reference's
golden CSV.  GPU part: `catt_link_leaves` / `catt_near_any` against the C restatement, bit-exact, and the golden CSV /
a deep synthetic repertoire through nbsorb with the library's scans injected.
"""
import math
import os
import random

import numpy as np
import pytest

from oracle import downstream as D
from oracle import oracle as O


def rand_seq(rng, lgt, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(lgt))


def mutate(rng, s, k, alphabet="ACGT"):
    s = list(s)
    for i in rng.sample(range(len(s)), k):
        s[i] = rng.choice([c for c in alphabet if c != s[i]])
    return "".join(s)


def make_group(seed, lgt, n_roots, n_leaves, max_mut=8, with_n=False):
    """Roots, their uplimit rows, leaves (mutated roots, a few unrelated, some near two roots) and leaf counts."""
    rng = random.Random(seed)
    alpha = "ACGTN" if with_n else "ACGT"
    roots = [rand_seq(rng, lgt, alpha) for _ in range(n_roots)]
    for i in range(1, n_roots, 5):                       # siblings: leaves near both -> statu 2
        roots[i] = mutate(rng, roots[i - 1], min(2, lgt), alpha)
    leaves = []
    for _ in range(n_leaves):
        r = rng.random()
        if r < 0.15:
            leaves.append(rand_seq(rng, lgt, alpha))
        else:
            leaves.append(mutate(rng, rng.choice(roots), rng.randint(0, min(max_mut, lgt)), alpha))
    counts = [rng.randint(1, 60) for _ in leaves]
    p, L = 0.004, max(lgt - 6, 1)
    up = []
    for _ in roots:                                         # root_up as nbsorb builds it (Jtool.jl:515-519) for a random root count
        rc = rng.choice([3, 20, 400, 5000, 100000])
        nt = math.floor(D.max_value(p, L, rc))
        up.append([D.junfen(D.error_number(p, L, nt, x, 3.291), L * 3 ** x, 3.291) for x in range(1, 8)])
    return roots, up, leaves, counts


def py_link(roots, up, leaves, counts):
    rr = [((r, ""), 1) for r in roots]
    return [D.link_rlg(((l, ""), c), rr, up) for l, c in zip(leaves, counts)]


@pytest.mark.parametrize("lgt,nr,nl", [(21, 7, 120), (45, 40, 300), (64, 3, 50), (105, 25, 200)])
def test_oracle_scans_follow_the_python_port(lgt, nr, nl):
    roots, up, leaves, counts = make_group(lgt * 1000 + nr, lgt, nr, nl, with_n=(lgt == 45))
    st, hd = O.link_leaves(lgt, roots, up, leaves, counts)
    assert [(int(a), int(b)) for a, b in zip(st, hd)] == py_link(roots, up, leaves, counts)
    assert set(np.unique(st)) <= {0, 1, 2} and len(set(np.unique(st))) >= 2
    fl = O.near_any(lgt, leaves, roots)
    assert [bool(x) for x in fl] == [any(D.hamming_g3(l, r, 3, 3) for r in roots) for l in leaves]
    fl = O.near_any(lgt, leaves, roots, upper=1, skip=0)
    assert [bool(x) for x in fl] == [any(D.hamming_g3(l, r, 1, 0) for r in roots) for l in leaves]


def deep_group(seed, lgt=45, n_clones=12, top=6000, err=0.004):
    """A deep repertoire of one CDR3 length: a few abundant clonotypes, each read with i.i.d. substitutions."""
    rng = random.Random(seed)
    val = []
    for c in range(n_clones):
        s = rand_seq(rng, lgt)
        for _ in range(max(3, top >> c)):
            t = list(s)
            q = ["I"] * lgt
            for i in range(lgt):
                if rng.random() < err:
                    t[i] = rng.choice([x for x in "ACGT" if x != t[i]])
                    q[i] = "#"
            val.append(("".join(t), "".join(q)))
    rng.shuffle(val)
    return val


def test_oracle_scans_inside_nbsorb():
    """nbsorb on a deep group really runs linkRLG (leaves, trees) and gives the same answer with the C scans."""
    val = deep_group(5)
    calls = {"link": 0, "near": 0, "leaves": 0}

    def link_fn(lgt, roots, up, leaves, counts):
        calls["link"] += 1; calls["leaves"] += len(leaves)
        return O.link_leaves(lgt, roots, up, leaves, counts)

    def near_fn(lgt, q, t):
        calls["near"] += 1
        return O.near_any(lgt, q, t)

    a = D.nbsorb(45, list(val), 0.004)
    b = D.nbsorb(45, list(val), 0.004, link_fn=link_fn, near_fn=near_fn)
    assert a == b and calls["link"] == 1 and calls["near"] == 1 and calls["leaves"] > 0
    assert len(b[1]) > 0                                    # some leaves were merged into their root


# ---------------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx(trb):
    from catt_b200.host import Catt
    c = Catt(chain_cfg=trb, device=0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lgt,nr,nl,with_n", [(21, 7, 500, False), (32, 130, 700, True), (33, 129, 2000, False), (64, 1, 300, False),
                                              (65, 300, 70000, False), (96, 1000, 300, True), (105, 257, 5000, False),
                                              (128, 64, 1000, False), (7, 5, 40, False), (6, 5, 40, False)])
def test_gpu_link_leaves_bit_exact(ctx, lgt, nr, nl, with_n):
    roots, up, leaves, counts = make_group(lgt * 77 + nr, lgt, nr, nl, with_n=with_n)
    st, hd = ctx.link_leaves(lgt, roots, up, leaves, counts)
    est, ehd = O.link_leaves(lgt, roots, up, leaves, counts)
    assert np.array_equal(st, est) and np.array_equal(hd, ehd)
    if lgt > 7:
        assert len(set(np.unique(st))) >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("lgt,nq,nt,upper,skip", [(21, 300, 10, 3, 3), (45, 100000, 700, 3, 3), (33, 50, 20000, 3, 3), (105, 3000, 3000, 3, 3),
                                                  (64, 1000, 100, 0, 0), (90, 1000, 129, 5, 10), (40, 500, 30, -1, 3), (30, 100, 10, 3, 15)])
def test_gpu_near_any_bit_exact(ctx, lgt, nq, nt, upper, skip):
    rng = random.Random(lgt * 31 + nq)
    targets = [rand_seq(rng, lgt, "ACGTN" if lgt == 45 else "ACGT") for _ in range(nt)]
    # half the queries are unrelated, half are targets with 0..7 substitutions (targets taken from the END of the list as
    # often as from the start, so the early exit and the split over the targets both matter)
    queries = [rand_seq(rng, lgt) if rng.random() < 0.5 else mutate(rng, targets[-1 - rng.randrange(nt)] if rng.random() < 0.5
               else targets[rng.randrange(nt)], rng.randint(0, 7), "ACGTN" if lgt == 45 else "ACGT") for _ in range(nq)]
    got = ctx.near_any(lgt, queries, targets, upper, skip)
    exp = O.near_any(lgt, queries, targets, upper, skip)
    assert np.array_equal(got, exp)
    if 2 * skip < lgt:
        assert 0 < int(got.sum()) < nq
    else:
        assert int(got.sum()) == nq                         # empty window: HammingDistanceG3 is vacuously true


@pytest.mark.gpu
def test_gpu_scan_edge_cases(ctx):
    from catt_b200 import _lib
    st, hd = ctx.link_leaves(30, [], np.zeros((0, 7)), ["A" * 30, "C" * 30], [3, 4])
    assert st.tolist() == [0, 0] and hd.tolist() == [1, 1]
    st, hd = ctx.link_leaves(30, ["A" * 30], [[5, 4, 3, 2, 1, 1, 1]], [], [])
    assert st.size == 0
    assert ctx.near_any(30, ["A" * 30], []).tolist() == [0]
    assert ctx.near_any(30, [], ["A" * 30]).size == 0
    with pytest.raises(_lib.CattError):
        ctx.near_any(30, ["A" * 29 + "x"], ["A" * 30])
    with pytest.raises(_lib.CattError):
        ctx.near_any(129, ["A" * 129], ["A" * 129])
    # ties of |uplimit - count|: the FIRST minimum wins (argmin), so reach 1 here although reach 7 is as close
    up = [[10.0, 0.0, 0.0, 0.0, 0.0, 0.0, 20.0]]
    root = "ACGT" * 10
    leaf = root[:10] + "TA" + root[12:]                     # 2 mismatches inside the window
    assert ctx.link_leaves(40, [root], up, [leaf], [15])[0].tolist() == O.link_leaves(40, [root], up, [leaf], [15])[0].tolist() == [0]
    assert ctx.link_leaves(40, [root], up, [leaf], [16])[0].tolist() == O.link_leaves(40, [root], up, [leaf], [16])[0].tolist() == [1]


@pytest.mark.gpu
def test_gpu_scans_inside_nbsorb_deep_and_golden(ctx, fx, trb):
    """nbsorb with the library's scans: a deep synthetic group (leaves, trees, merges) and the reference's golden CSV."""
    used = {"leaves": 0, "maybe": 0}

    def link_fn(lgt, roots, up, leaves, counts):
        used["leaves"] += len(leaves)
        return ctx.link_leaves(lgt, roots, up, leaves, counts)

    def near_fn(lgt, q, t):
        used["maybe"] += len(q)
        return ctx.near_any(lgt, q, t)

    for seed, lgt in ((5, 45), (6, 60), (7, 33)):
        val = deep_group(seed, lgt)
        assert D.nbsorb(lgt, list(val), 0.004) == D.nbsorb(lgt, list(val), 0.004, link_fn=link_fn, near_fn=near_fn)
    assert used["leaves"] > 0 and used["maybe"] > 0
    refs = O.RefSet(trb["vregion"]), O.RefSet(trb["jregion"])
    R = O.run_file(refs[0], refs[1], O.make_config(trb, allow_v=0), os.path.join(fx, "testSample.fq"))
    gold, _ = D.load_golden_csv(os.path.join(fx, "aha.TRB.CDR3.CATT.csv"))
    tab = D.clonotype_table(R.Vpart, R.Jpart, R.info["the_err"], trb.get("Dregion"), O.best_d, link_fn=link_fn, near_fn=near_fn)
    assert tab == gold


@pytest.mark.gpu
def test_gpu_near_any_full_size_property(ctx):
    """1 M low-count sequences against 10 k high-confidence ones (no oracle at this size): every planted neighbour
    (<= 3 substitutions inside the window, anything at the skipped ends) is flagged, and a sample agrees with the oracle."""
    rng = np.random.default_rng(3)
    lgt, nq, nt = 45, 1_000_000, 10_000
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    targets = lut[rng.integers(0, 4, size=(nt, lgt))]
    queries = lut[rng.integers(0, 4, size=(nq, lgt))]
    planted = np.arange(0, nq, 7)
    src = rng.integers(0, nt, size=planted.size)
    queries[planted] = targets[src]
    for k in range(3):                                       # up to 3 substitutions inside the window
        pos = rng.integers(3, lgt - 3, size=planted.size)
        queries[planted, pos] = lut[rng.integers(0, 4, size=planted.size)]
    queries[planted, 0] = ord("N"); queries[planted, lgt - 1] = ord("N")       # the skipped ends do not count
    got = ctx.near_any(lgt, queries, targets)
    assert got[planted].all()
    sample = np.arange(0, nq, 997)
    exp = O.near_any(lgt, ["".join(map(chr, queries[i])) for i in sample], ["".join(map(chr, t)) for t in targets])
    assert np.array_equal(got[sample], exp)
    assert int(got.sum()) < nq // 6                          # random 39-mers are essentially never within 3 of 10 k others
