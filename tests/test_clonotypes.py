"""The clonotype table (catt_config.clonotypes, clonotype.cu): distinct (CDR3, V, J) with counts and merged quality strings, built on
the device from the result lists - what nbsorb's grouping (Jtool.jl:475-490) and the output counter (catt.jl:644) compute from the
reads.  Checked against the same table computed from the returned lists in Python, against the reference's golden CSV (the table
alone, through the downstream port, must reproduce all 721 rows), on a deep repertoire with heavy duplication, and on a GPU group."""
import collections
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def table_from_lists(res):
    """(cdr3, V, J) -> [count, merged qual]; rows in the library's order: (len, cdr3, V id, J id) with "None" (-1) first."""
    t = {}
    for part in (res.Vpart, res.Jpart):
        for r in part:
            if r.cdr3 == "None":
                continue
            k = (r.cdr3, r.vs, r.js)
            if k in t:
                t[k][0] += 1
                t[k][1] = "".join(max(a, b) for a, b in zip(t[k][1], r.qual))
            else:
                t[k] = [1, r.qual]
    return t


def check_table(ctx, res):
    rows = res.clonotypes()
    want = table_from_lists(res)
    got = {(c, v, j): [n, q] for c, q, v, j, n in rows}
    assert len(got) == len(rows), "duplicate rows"
    assert got == want
    vid = {n: i for i, n in enumerate(ctx.v_names)}; vid["None"] = -1
    jid = {n: i for i, n in enumerate(ctx.j_names)}; jid["None"] = -1
    keys = [(len(c), c, vid[v], jid[j]) for c, q, v, j, n in rows]
    assert keys == sorted(keys), "rows are not in (len, cdr3, V id, J id) order"
    return rows


def test_table_matches_lists_and_golden_csv(trb, fx):
    from catt_b200.host import Catt
    from oracle import downstream as D
    fq = os.path.join(fx, "testSample.fq")
    gold, _ = D.load_golden_csv(os.path.join(fx, "aha.TRB.CDR3.CATT.csv"))
    for kw in ({}, {"only_cdr3": True}, {"only_cdr3": True, "ngpu": 3, "devices": [0, 0, 0]}):
        c = Catt(chain_cfg=trb, allow_v=0, clonotypes=True, **kw)
        try:
            res = c.mainflow(fq)
            rows = check_table(c, res)
            assert sum(r[4] for r in rows) == sum(1 for p in (res.Vpart, res.Jpart) for r in p if r.cdr3 != "None")
            # the reference's own output from the TABLE alone: no read is looked at downstream
            tab = D.clonotype_table_from_rows(rows, res.info["the_err"], trb["Dregion"], lambda nn, dl: c.selBestMatchD([nn])[0])
            assert tab == gold
            assert tab == D.clonotype_table(res.Vpart, res.Jpart, res.info["the_err"], trb["Dregion"], lambda nn, dl: c.selBestMatchD([nn])[0])
        finally:
            c.close()
    off = Catt(chain_cfg=trb)                         # not asked for: no table
    try:
        assert off.mainflow(fq).clonotypes() == []
    finally:
        off.close()


def test_deep_repertoire_with_duplicates(trb):
    """A few hundred clonotypes at very different depths, qualities varied per read: counts add up, the merged quality is the
    per-position maximum, the same CDR3 under different V/J calls stays in separate (adjacent) rows."""
    from catt_b200.host import Catt
    from oracle import oracle as O
    V, J = O.RefSet(trb["vregion"]), O.RefSet(trb["jregion"])
    seqs, quals = O.synth_reads(V, J, 4000, 150, seed=77)
    rng = np.random.default_rng(3)
    reps = np.minimum(rng.zipf(1.3, size=4000), 400)
    idx = np.repeat(np.arange(4000), reps)
    rng.shuffle(idx)
    S = seqs[idx].copy()
    Q = rng.integers(ord("#"), ord("J"), size=S.shape, dtype=np.uint8)
    names = [f"r{i}" for i in range(len(idx))]
    sl = ["".join(map(chr, r)) for r in S]
    ql = ["".join(map(chr, r)) for r in Q]
    for kw in ({"only_cdr3": True}, {"only_cdr3": False}, {"only_cdr3": True, "ngpu": 2, "devices": [0, 0]}):
        c = Catt(chain_cfg=trb, clonotypes=True, **kw)
        try:
            res = c.mainflow_reads(names, sl, ql)
            rows = check_table(c, res)
            assert max(r[4] for r in rows) >= 50 and len(rows) >= 20
        finally:
            c.close()


def test_table_equals_oracle_lists(trb):
    """Parity, not only self-consistency: the table against one built from the CPU checker's lists."""
    from catt_b200.host import Catt
    from oracle import oracle as O
    V, J = O.RefSet(trb["vregion"]), O.RefSet(trb["jregion"])
    n = 60000
    c = Catt(chain_cfg=trb, only_cdr3=True, clonotypes=True)
    try:
        bufs = c.synth_reads(n, 150)
        seqs, quals = O.synth_reads(V, J, n, 150)
        ref = O.run_flat(V, J, O.make_config(trb), seqs, quals)
        res = c.mainflow_buffers(n, *bufs)
        want = table_from_lists(ref)
        got = {(cd, v, j): [cnt, q] for cd, q, v, j, cnt in res.clonotypes()}
        assert got == want and len(got) > 300
    finally:
        c.close()
