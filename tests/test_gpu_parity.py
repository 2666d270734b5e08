"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.
Bit-exact: everything on this path is integer / byte / index work."""
import os

import numpy as np
import pytest

from util import myread_tuple, random_reads, revcomp, synth_reads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def sample(fx):
    names, seqs, quals = [], [], []
    with open(os.path.join(fx, "testSample.fq")) as fh:
        lines = fh.read().split("\n")
    for i in range(0, len(lines) - 3, 4):
        names.append(lines[i][1:].split()[0]); seqs.append(lines[i + 1]); quals.append(lines[i + 3])
    assert len(seqs) == 7922
    return names, seqs, quals


@pytest.fixture(scope="module")
def ctx(trb):
    from catt_b200.host import Catt
    c = Catt(chain_cfg=trb, device=0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orefs(O, trb):
    return O.RefSet(trb["vregion"]), O.RefSet(trb["jregion"])


REC_KEYS = [("rid", "rid"), ("strand", "strand"), ("mapped", "mapped"), ("as", "AS"), ("qb", "qb"), ("m_end", "m_end"),
            ("rb", "rb"), ("qe", "qe"), ("re", "re"), ("nm", "nm"), ("mi", "mi"), ("ncand", "ncand"), ("ntied", "ntied")]


def assert_records_equal(gpu, ora, O):
    cols = {k: i for i, k in enumerate(O.REC_FIELDS)}
    mapped = ora[:, cols["mapped"]] == 1
    assert np.array_equal(gpu["mapped"] == 1, mapped)
    for gk, ok in REC_KEYS:
        g, o = gpu[gk].astype(np.int64), ora[:, cols[ok]].astype(np.int64)
        sel = mapped if gk not in ("mapped", "as", "ncand") else np.ones_like(mapped)
        bad = np.nonzero((g != o) & sel)[0]
        assert bad.size == 0, f"{gk}: {bad.size} mismatches, first at read {bad[:5]}: gpu {g[bad[:5]]} oracle {o[bad[:5]]}"


# ---------------------------------------------------------------------------------------------------------------------
def test_sw_kernel_random_and_real_pairs(ctx, O, orefs, sample):
    rng = np.random.default_rng(7)
    seqs, rid, strand = [], [], []
    V = orefs[0]
    for s in sample[1][:1500]:                       # real reads against random records, both strands
        seqs.append(s); rid.append(int(rng.integers(0, V.nrec))); strand.append(int(rng.integers(0, 2)))
    recs = [V.record(i) for i in range(V.nrec)]
    for i in range(1500):                            # mutated copies of records with indels
        r = int(rng.integers(0, V.nrec))
        t = list(recs[r])
        for _ in range(int(rng.integers(0, 6))):
            p = int(rng.integers(0, len(t)))
            op = int(rng.integers(0, 3))
            if op == 0: t[p] = "ACGT"[int(rng.integers(0, 4))]
            elif op == 1: del t[p]
            else: t.insert(p, "ACGT"[int(rng.integers(0, 4))])
        s = "".join(random_reads(1, int(rng.integers(0, 40)), int(rng.integers(1 << 30)))[0:1]) + "".join(t) + \
            random_reads(1, int(rng.integers(1, 40)), int(rng.integers(1 << 30)))[0]
        sd = int(rng.integers(0, 2))
        seqs.append(revcomp(s) if sd else s); rid.append(r); strand.append(sd)
    out = ctx.sw_pairs(0, seqs, rid, strand)
    for i, (s, r, sd) in enumerate(zip(seqs, rid, strand)):
        q = revcomp(s) if sd else s
        sc, qi, tj = O.sw(q, recs[r])
        if sc == 0:
            assert out[i, 0] == 0
        else:
            assert tuple(out[i]) == (sc, qi, tj), (i, tuple(out[i]), (sc, qi, tj))


def test_sw_kernel_reads_with_n(ctx, O, orefs):
    V = orefs[0]
    rng = np.random.default_rng(11)
    seqs, rid, strand = [], [], []
    for i in range(300):
        r = int(rng.integers(0, V.nrec))
        t = list(V.record(r))
        for _ in range(3):
            t[int(rng.integers(0, len(t)))] = "N"
        seqs.append("ACGTAC" + "".join(t) + "TTGACA"); rid.append(r); strand.append(0)
    out = ctx.sw_pairs(0, seqs, rid, strand)
    for i in range(300):
        assert tuple(out[i]) == O.sw(seqs[i], V.record(rid[i]))


@pytest.mark.parametrize("which", [0, 1])
def test_seeding_matches_bwa_model(ctx, orefs, sample, which):
    gpu = ctx.seed_candidates(which, sample[1])
    ora = orefs[which].seed(sample[1])
    bad = np.nonzero((gpu != ora).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} reads differ, first {bad[:5]}"
    assert int(ora.sum()) == (4137 if which else 54840)


def test_alignment_records_testsample(ctx, O, orefs, sample):
    v, j = ctx.align_shard(sample[1])
    assert_records_equal(v, orefs[0].align(sample[1]), O)
    assert_records_equal(j, orefs[1].align(sample[1]), O)
    assert int(v["mapped"].sum()) == 2653 and int(j["mapped"].sum()) == 1752


@pytest.mark.parametrize("env", [{"CATT_OLD_TRACK": "1"}, {"CATT_NO_DIRECT": "1"}, {"CATT_OLD_TRACK": "1", "CATT_NO_DIRECT": "1"}, {"CATT_TILE32": "2"}])
def test_alternative_kernels_give_the_same_records(trb, O, orefs, sample, monkeypatch, env):
    """The A/B switches select the previous kernels (TRACK pass by the PRMT kernel on unsorted winners; J sets through the bulk
    pass + TRACK pass; another lane tiling for short records): every path must produce the oracle's records, i.e. the grouped
    profile TRACK kernel, the direct J pass and the old kernels all agree cell for cell."""
    from catt_b200.host import Catt
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    c = Catt(chain_cfg=trb, device=0)
    try:
        v, j = c.align_shard(sample[1])
        assert_records_equal(v, orefs[0].align(sample[1]), O)
        assert_records_equal(j, orefs[1].align(sample[1]), O)
    finally:
        c.close()


def test_first_ordinal_shifts_the_tie_break(ctx, O, orefs, sample):
    """bwa's hash tie-break uses the GLOBAL read ordinal: a shard must reproduce it through first_ordinal."""
    sub = sample[1][3000:3800]
    v, _ = ctx.align_shard(sub, first_ordinal=3000)
    assert_records_equal(v, orefs[0].align(sub, first_ordinal=3000), O)


@pytest.mark.parametrize("allow_v,threads", [(0, 1), (3, 1), (3, 4)])
def test_end_to_end_testsample(trb, fx, O, orefs, sample, allow_v, threads):
    from catt_b200.host import Catt
    from oracle import downstream as D
    c = Catt(chain_cfg=trb, allow_v=allow_v, thread=threads)
    res = c.mainflow(os.path.join(fx, "testSample.fq"))
    ref = O.run_file(orefs[0], orefs[1], O.make_config(trb, allow_v=allow_v, nthreads=threads), os.path.join(fx, "testSample.fq"))
    assert [myread_tuple(r) for r in res.Vpart] == [myread_tuple(r) for r in ref.Vpart]
    assert [myread_tuple(r) for r in res.Jpart] == [myread_tuple(r) for r in ref.Jpart]
    for k in ("kmer", "v_hits", "j_hits", "v_final", "j_final", "share", "pair_seq_mismatch", "full_pushed", "vpart", "jpart",
              "direct", "left", "rescued", "cand_v", "cand_j", "cells_v", "cells_j"):
        assert res.info[k] == ref.info[k], (k, res.info[k], ref.info[k])
    assert res.info["the_err"] == ref.info["the_err"]
    assert res.info["gapped_v"] == ref.info["v_gapped"] and res.info["gapped_j"] == ref.info["j_gapped"]
    if allow_v == 0:      # the reference's own golden output, through the unchanged-Julia stand-in
        gold, _ = D.load_golden_csv(os.path.join(fx, "aha.TRB.CDR3.CATT.csv"))
        dnames = c.selBestMatchD
        tab = D.clonotype_table(res.Vpart, res.Jpart, res.info["the_err"], trb["Dregion"], lambda nn, dl: dnames([nn])[0])
        assert tab == gold
    c.close()


def test_selbestmatchd_golden(ctx, fx):
    from oracle import downstream as D
    gold, _ = D.load_golden_csv(os.path.join(fx, "aha.TRB.CDR3.CATT.csv"))
    keys = list(gold)
    got = ctx.selBestMatchD([k[1] for k in keys])
    assert got == [k[4] for k in keys]


@pytest.mark.parametrize("length,n", [(150, 6000), (100, 3000), (50, 3000)])
def test_synthetic_trb_end_to_end(ctx, O, orefs, trb, length, n):
    V, J = orefs
    vrecs = [V.record(i) for i in range(V.nrec)]
    jrecs = [J.record(i) for i in range(J.nrec)]
    names, seqs, quals = synth_reads(n, vrecs, jrecs, length, with_n=True)
    v, j = ctx.align_shard(seqs)
    assert_records_equal(v, V.align(seqs), O)
    assert_records_equal(j, J.align(seqs), O)
    res = ctx.mainflow_reads(names, seqs, quals)
    ref = O.run_mem(V, J, O.make_config(trb), names, seqs, quals)
    assert [myread_tuple(r) for r in res.Vpart] == [myread_tuple(r) for r in ref.Vpart]
    assert [myread_tuple(r) for r in res.Jpart] == [myread_tuple(r) for r in ref.Jpart]
    for k in ("kmer", "vpart", "jpart", "direct", "left", "rescued"):
        assert res.info[k] == ref.info[k], (k, res.info[k], ref.info[k])
    assert res.info["the_err"] == ref.info["the_err"]


@pytest.mark.parametrize("species,chain", [("hs", "TRA"), ("hs", "IGH"), ("hs", "TRG"), ("hs", "TRD"), ("ms", "TRB"), ("pig", "TRB")])
def test_other_chains(fx, O, species, chain):
    from catt_b200.host import Catt
    from catt_b200.refg import chain_config
    cc = chain_config(species, chain, "CDR3", os.path.join(fx, "resource"))
    V, J = O.RefSet(cc["vregion"]), O.RefSet(cc["jregion"])
    vrecs = [V.record(i) for i in range(V.nrec)]
    jrecs = [J.record(i) for i in range(J.nrec)]
    names, seqs, quals = synth_reads(1500, vrecs, jrecs, 150, seed=99)
    c = Catt(chain_cfg=cc)
    v, j = c.align_shard(seqs)
    assert_records_equal(v, V.align(seqs), O)
    assert_records_equal(j, J.align(seqs), O)
    res = c.mainflow_reads(names, seqs, quals)
    ref = O.run_mem(V, J, O.make_config(cc), names, seqs, quals)
    assert [myread_tuple(r) for r in res.Vpart] == [myread_tuple(r) for r in ref.Vpart]
    assert [myread_tuple(r) for r in res.Jpart] == [myread_tuple(r) for r in ref.Jpart]
    if "Dregion" in cc:
        cd = [r.cdr3 for r in res.Vpart if r.cdr3 != "None"][:200]
        assert c.selBestMatchD(cd) == [O.best_d(s, cc["Dregion"]) for s in cd]
    c.close()


def test_finder_hook(ctx, O, trb, sample):
    cfg = O.make_config(trb)
    seqs = [s for s in sample[1][:1200] if "N" not in s]
    for arr in (False, True):
        out = ctx.finder(seqs, array_version=arr)
        for i, s in enumerate(seqs):
            able, rng_ = O.finder(cfg, s, arr)
            assert bool(out[i, 0]) == able, (i, arr)
            if rng_ is None:
                assert out[i, 1] == -1
            else:
                assert (out[i, 1] + 1, out[i, 1] + out[i, 2]) == rng_, (i, arr, out[i], rng_)


def test_edge_cases(ctx, O, orefs, trb):
    from catt_b200 import _lib
    # empty input
    res = ctx.mainflow_reads([], [], [])
    assert res.info["n_reads"] == 0 and res.Vpart == [] and res.Jpart == []
    # reads shorter than a seed, all-N reads, ragged lengths, the maximum length
    V = orefs[0]
    rec = V.record(5)
    seqs = ["ACGT", "N" * 30, rec[:9], rec[:10], rec, rec + "ACGT" * 100, ("ACGT" * 128)[:512], revcomp(rec) + "GG"]
    names = [f"e{i}" for i in range(len(seqs))]
    quals = ["I" * len(s) for s in seqs]
    v, j = ctx.align_shard(seqs)
    assert_records_equal(v, V.align(seqs), O)
    assert_records_equal(j, orefs[1].align(seqs), O)
    res = ctx.mainflow_reads(names, seqs, quals)
    ref = O.run_mem(orefs[0], orefs[1], O.make_config(trb), names, seqs, quals)
    assert [myread_tuple(r) for r in res.Vpart] == [myread_tuple(r) for r in ref.Vpart]
    # longer than the compiled-in limit: loud error, not truncation
    with pytest.raises(_lib.CattError) as ei:
        ctx.align_shard(["A" * 513])
    assert ei.value.code == -4


def test_mate_names_collapse(ctx, O, orefs, trb, sample):
    """/1 and /2 mates share a trimmed QNAME: the awk dedup keeps one record per name and SEQ-mismatching V/J pairs drop."""
    names, seqs, quals = sample
    idx = list(range(0, 600)) + list(range(4200, 4800))
    n, s, q = [names[i] for i in idx], [seqs[i] for i in idx], [quals[i] for i in idx]
    res = ctx.mainflow_reads(n, s, q)
    ref = O.run_mem(orefs[0], orefs[1], O.make_config(trb), n, s, q)
    assert [myread_tuple(r) for r in res.Vpart] == [myread_tuple(r) for r in ref.Vpart]
    assert [myread_tuple(r) for r in res.Jpart] == [myread_tuple(r) for r in ref.Jpart]
    assert res.info["pair_seq_mismatch"] == ref.info["pair_seq_mismatch"]


def test_stage_split_equals_whole(ctx, trb, sample):
    """catt_align_shard over two shards + catt_assemble == catt_run_reads (the multi-GPU decomposition)."""
    names, seqs, quals = [x[:3000] for x in sample]
    whole = ctx.mainflow_reads(names, seqs, quals)
    v1, j1 = ctx.align_shard(seqs[:1700], first_ordinal=0)
    v2, j2 = ctx.align_shard(seqs[1700:], first_ordinal=1700)
    res = ctx.assemble(names, seqs, quals, np.concatenate([v1, v2]), np.concatenate([j1, j2]))
    assert [myread_tuple(r) for r in res.Vpart] == [myread_tuple(r) for r in whole.Vpart]
    assert [myread_tuple(r) for r in res.Jpart] == [myread_tuple(r) for r in whole.Jpart]


def test_paired_end_files(trb, fx, O, orefs, sample, tmp_path):
    """--f1/--f2 (catt.jl:738-745): each mate file is aligned and de-duplicated on its own, kmer comes from the first file, the_ERR
    from the last, and one k-mer table / extension pass runs over the joined lists.  Also reads .gz input."""
    import gzip
    from catt_b200.host import Catt
    names, seqs, quals = sample
    f1 = [(n, s, q) for n, s, q in zip(names, seqs, quals) if n.endswith("/1")][:1500]
    f2 = [(n, s, q) for n, s, q in zip(names, seqs, quals) if n.endswith("/2")][:1400]
    assert len(f1) > 1000 and len(f2) > 1000
    files = [tuple(map(list, zip(*f1))), tuple(map(list, zip(*f2)))]
    for threads in (1, 3):
        c = Catt(chain_cfg=trb, thread=threads)
        ref = O.run_mem_files(orefs[0], orefs[1], O.make_config(trb, nthreads=threads), files)
        res = c.mainflow_reads_files(files)
        assert [myread_tuple(r) for r in res.Vpart] == [myread_tuple(r) for r in ref.Vpart]
        assert [myread_tuple(r) for r in res.Jpart] == [myread_tuple(r) for r in ref.Jpart]
        for k in ("kmer", "v_hits", "j_hits", "v_final", "j_final", "share", "pair_seq_mismatch", "full_pushed", "vpart", "jpart",
                  "direct", "left", "rescued"):
            assert res.info[k] == ref.info[k], (k, res.info[k], ref.info[k])
        assert res.info["the_err"] == ref.info["the_err"]
        if threads == 1:                      # the same sample from two files on disk, the second one gzip-compressed
            p1, p2 = str(tmp_path / "m_1.fq"), str(tmp_path / "m_2.fq.gz")
            with open(p1, "w") as fh:
                fh.write("".join(f"@{n}\n{s}\n+\n{q}\n" for n, s, q in f1))
            with gzip.open(p2, "wt") as fh:
                fh.write("".join(f"@{n}\n{s}\n+\n{q}\n" for n, s, q in f2))
            res2 = c.mainflow_paired(p1, p2)
            assert [myread_tuple(r) for r in res2.Vpart] == [myread_tuple(r) for r in ref.Vpart]
            assert [myread_tuple(r) for r in res2.Jpart] == [myread_tuple(r) for r in ref.Jpart]
        # joining the mates into one file is NOT the same thing (the dedup and the V/J pairing would cross the files)
        c.close()


def test_fasta_input_multiline(ctx, O, orefs, trb, sample, tmp_path):
    """FASTA input (no qualities -> 'I'), sequences wrapped over several lines."""
    names, seqs, _ = [x[:800] for x in sample]
    p = str(tmp_path / "s.fa")
    with open(p, "w") as fh:
        for n, s in zip(names, seqs):
            fh.write(f">{n} some comment\n" + "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n")
    res = ctx.mainflow(p)
    ref = O.run_mem(orefs[0], orefs[1], O.make_config(trb), names, seqs, ["I" * len(s) for s in seqs])
    assert [myread_tuple(r) for r in res.Vpart] == [myread_tuple(r) for r in ref.Vpart]
    assert [myread_tuple(r) for r in res.Jpart] == [myread_tuple(r) for r in ref.Jpart]


@pytest.mark.parametrize("chain", ["TRA", "IGH"])
def test_short_reads_other_chains(fx, O, chain):
    """BASELINE configs[3]: 50-bp reads (k = 15 path, most reads are partial -> k-mer table + extension) on hs TRA / IGH."""
    from catt_b200.host import Catt
    from catt_b200.refg import chain_config
    cc = chain_config("hs", chain, "CDR3", os.path.join(fx, "resource"))
    V, J = O.RefSet(cc["vregion"]), O.RefSet(cc["jregion"])
    names, seqs, quals = synth_reads(2500, [V.record(i) for i in range(V.nrec)], [J.record(i) for i in range(J.nrec)], 50, seed=404, with_n=True)
    c = Catt(chain_cfg=cc)
    res = c.mainflow_reads(names, seqs, quals)
    ref = O.run_mem(V, J, O.make_config(cc), names, seqs, quals)
    assert res.info["kmer"] == ref.info["kmer"] == 15
    assert [myread_tuple(r) for r in res.Vpart] == [myread_tuple(r) for r in ref.Vpart]
    assert [myread_tuple(r) for r in res.Jpart] == [myread_tuple(r) for r in ref.Jpart]
    for k in ("vpart", "jpart", "direct", "left", "rescued", "cand_v", "cand_j"):
        assert res.info[k] == ref.info[k], (k, res.info[k], ref.info[k])
    c.close()


def test_rnaseq_like_mix(ctx, O, orefs, trb):
    """BASELINE configs[1] substitute (the SRR11799691 input is not shipped): 1 % TRB-derived 100-bp reads among
    uniform-random reads, seed 11799691.  Almost every read has no seed; a few random reads hit by chance."""
    V, J = orefs
    vrecs = [V.record(i) for i in range(V.nrec)]
    jrecs = [J.record(i) for i in range(J.nrec)]
    rng = np.random.default_rng(11799691)
    n = 20000
    tn, ts, tq = synth_reads(n // 100, vrecs, jrecs, 100, seed=11799691)
    names, seqs, quals = [], [], []
    t = 0
    for i in range(n):
        if i % 100 == 37:
            names.append(f"t{t}"); seqs.append(ts[t]); quals.append(tq[t]); t += 1
        else:
            names.append(f"r{i}"); seqs.append("".join("ACGT"[x] for x in rng.integers(0, 4, 100))); quals.append("I" * 100)
    v, j = ctx.align_shard(seqs)
    assert_records_equal(v, V.align(seqs), O)
    assert_records_equal(j, J.align(seqs), O)
    res = ctx.mainflow_reads(names, seqs, quals)
    ref = O.run_mem(V, J, O.make_config(trb), names, seqs, quals)
    assert [myread_tuple(r) for r in res.Vpart] == [myread_tuple(r) for r in ref.Vpart]
    assert [myread_tuple(r) for r in res.Jpart] == [myread_tuple(r) for r in ref.Jpart]
    assert res.info["v_hits"] == ref.info["v_hits"] and res.info["v_hits"] < n // 10     # chance hits with AS >= 12 included


def test_fastq_ingest_on_device_equals_host_parser(ctx, O, orefs, trb, sample, tmp_path):
    """catt_run_file indexes, validates and packs strict 4-line FASTQ on the device (ingest.cu); everything else falls back to the
    host parser.  Both must give the oracle's lists: CRLF line ends, header comments, missing / extra trailing newlines."""
    from catt_b200 import _lib
    names, seqs, quals = [x[:1200] for x in sample]
    ref = O.run_mem(orefs[0], orefs[1], O.make_config(trb), names, seqs, quals)
    want = ([myread_tuple(r) for r in ref.Vpart], [myread_tuple(r) for r in ref.Jpart])
    body = lambda nl, tail="": "".join(f"@{n} 1:N:0 some\tcomment{nl}{s}{nl}+{n}{nl}{q}{nl}" for n, s, q in zip(names, seqs, quals)) + tail
    variants = {"plain.fq": body("\n"), "crlf.fq": body("\r\n"), "notail.fq": body("\n")[:-1], "blanktail.fq": body("\n", "\n\n \n"),
                "blankmid.fq": body("\n").replace("\n@", "\n\n@", 3)}          # blank lines inside: host parser (kseq skips them)
    for fn, txt in variants.items():
        p = str(tmp_path / fn)
        with open(p, "w", newline="") as fh:
            fh.write(txt)
        for host in (False, True):
            if host:
                os.environ["CATT_HOST_PARSE"] = "1"
            try:
                res = ctx.mainflow(p)
            finally:
                os.environ.pop("CATT_HOST_PARSE", None)
            assert ([myread_tuple(r) for r in res.Vpart], [myread_tuple(r) for r in res.Jpart]) == want, (fn, host)
            assert res.info["n_reads"] == len(seqs)
    # a read longer than the compiled-in limit: loud error from either path
    p = str(tmp_path / "long.fq")
    with open(p, "w") as fh:
        fh.write(f"@x\n{'A' * 600}\n+\n{'I' * 600}\n")
    with pytest.raises(_lib.CattError) as ei:
        ctx.mainflow(p)
    assert ei.value.code == -4
    # truncated record: the device path declines, the host parser reports it
    p = str(tmp_path / "trunc.fq")
    with open(p, "w") as fh:
        fh.write(body("\n")[:-200])
    with pytest.raises(_lib.CattError) as ei:
        ctx.mainflow(p)
    assert ei.value.code == -2


def test_large_result_ring_copy(ctx, O, orefs, trb, sample):
    """Results too large to pin per call come back through the two pinned ring buffers (forced here with CATT_PIN_MAX): the lists
    must still be the oracle's."""
    seqs, quals = sample[1][:7000] * 8, sample[2][:7000] * 8  # > 1 MiB of strings: above the malloc threshold
    names = [f"c{k}_{nm}" for k in range(8) for nm in sample[0][:7000]]       # eight copies with QNAMEs of their own (mate suffixes kept)
    ref = O.run_mem(orefs[0], orefs[1], O.make_config(trb), names, seqs, quals)
    want = ([myread_tuple(r) for r in ref.Vpart], [myread_tuple(r) for r in ref.Jpart])
    os.environ["CATT_PIN_MAX"] = "1"
    try:
        got = ctx.mainflow_reads(names, seqs, quals)
    finally:
        os.environ.pop("CATT_PIN_MAX", None)
    assert ([myread_tuple(r) for r in got.Vpart], [myread_tuple(r) for r in got.Jpart]) == want
    assert got.text_digest() == ref.digest(False)
    again = ctx.mainflow_reads(names, seqs, quals)            # ... and the pinned-block path gives the same
    assert again.digest() == got.digest()
    assert len(want[0]) > 2000


def test_real_score_hook(ctx, O, orefs, sample):
    """real_score + HammingDistanceG4 (catt.jl:58-84) alone: every mapped record of the sample at the allowances assignV / assignJ
    use (3, 9), the golden-file setting (0) and a generous one, plus records shifted so that the paddings hit both ends."""
    seqs = sample[1]
    for which in (0, 1):
        recs = ctx.align_shard(seqs)[which]
        idx = np.nonzero(recs["mapped"] == 1)[0]
        rs = [seqs[i] for i in idx]
        f = {k: recs[k][idx].astype(np.int32) for k in ("rid", "strand", "qb", "m_end", "rb")}
        n_pass = {}
        for allow in (0, 3, 9, 40):
            g = ctx.real_score(which, rs, f["rid"], f["strand"], f["qb"], f["m_end"], f["rb"], allow)
            o = O.real_score(orefs[which], rs, f["rid"], f["strand"], f["qb"], f["m_end"], f["rb"], allow)
            assert np.array_equal(g, o), (which, allow, int((g != o).sum()))
            n_pass[allow] = int(g.sum())
        assert 0 < n_pass[0] <= n_pass[3] <= n_pass[9] <= n_pass[40] <= len(idx)
        if which == 0:
            assert n_pass[0] < n_pass[40], "the V records of the sample do discriminate the allowance (SURVEY 0.5)"


@pytest.mark.parametrize("threads", [1, 2, 5])
def test_order_records_hook(trb, O, orefs, sample, threads):
    """The ordering stage alone (samtools sort -t AS | view -F 2308 | tac | awk dedup, share_name, the NR % nthreads chunk order, the
    name-sorted pairing): the work items and their order from the sample's records, and from records with forced ties - equal AS /
    record / position on several reads of one QNAME - where only the input order decides."""
    from catt_b200.host import Catt
    c = Catt(chain_cfg=trb, thread=threads)
    try:
        names, seqs, _ = sample
        v, j = c.align_shard(seqs)
        lens = [len(s) for s in seqs]

        def to16(r):
            out = np.zeros((len(r), 16), dtype=np.int32)
            for col, k in enumerate(("ordinal", "rid", "strand", "mapped", "as", "qb", "m_end", "rb", "qe", "re", "nm", "mi", "ncand", "ntied")):
                out[:, col] = r[k]
            return out
        for variant in range(2):
            if variant == 1:                              # collapse QNAMEs 4 to 1 and flatten the scores: ties everywhere
                names = [f"q{i // 4}/{1 + i % 2}" for i in range(len(names))]
                v, j = v.copy(), j.copy()
                v["as"] = np.where(v["mapped"] == 1, 40 + (np.arange(len(v)) % 3), v["as"])
                j["as"] = np.where(j["mapped"] == 1, 25, j["as"])
            items, cnt = c.order_records(names, lens, v, j)
            oitems, ocnt = O.order_items(names, to16(v), to16(j), threads)
            assert np.array_equal(cnt, ocnt), (variant, cnt, ocnt)
            assert np.array_equal(items, oitems), (variant, int((items != oitems).any(axis=1).sum()))
            assert cnt[3] > 100 and cnt[2] > 500
    finally:
        c.close()


def test_only_cdr3_lists(trb, O, orefs):
    """only_cdr3 = 1 returns the lists as they stand after catt.jl:581-582 (filter cdr3 != "None"): same reads, same order, same
    strings as the filtered full lists; the info counters still describe the full lists."""
    from catt_b200.host import Catt
    V, J = orefs
    names, seqs, quals = synth_reads(6000, [V.record(i) for i in range(V.nrec)], [J.record(i) for i in range(J.nrec)], 100, seed=77)
    full_ctx, cdr3_ctx = Catt(chain_cfg=trb), Catt(chain_cfg=trb, only_cdr3=True)
    full, comp = full_ctx.mainflow_reads(names, seqs, quals), cdr3_ctx.mainflow_reads(names, seqs, quals)
    fv = [myread_tuple(r) for r in full.Vpart if r.cdr3 != "None"]
    fj = [myread_tuple(r) for r in full.Jpart if r.cdr3 != "None"]
    assert len(fv) + len(fj) > 20
    assert [myread_tuple(r) for r in comp.Vpart] == fv and [myread_tuple(r) for r in comp.Jpart] == fj
    for k in ("vpart", "jpart", "direct", "left", "rescued", "kmer", "the_err"):
        assert comp.info[k] == full.info[k]
    assert comp.info["bytes_d2h"] < full.info["bytes_d2h"] / 5
    ref = O.run_mem(V, J, O.make_config(trb), names, seqs, quals)
    assert [myread_tuple(r) for r in ref.Vpart if r.cdr3 != "None"] == fv
    full_ctx.close(); cdr3_ctx.close()


def test_file_pipeline_many_pieces(trb, fx, O, orefs, tmp_path):
    """File input with a tiny chunk size: the reader thread cuts the FASTQ into many record-aligned pieces (size ramp, several Stage A
    chunks per piece, growing device buffers); result == oracle.  Also: a one-read file, an empty file, paired files with small pieces."""
    from catt_b200 import _lib
    from catt_b200.host import Catt
    path = os.path.join(fx, "testSample.fq")
    ref = O.run_file(orefs[0], orefs[1], O.make_config(trb), path)
    want = ([myread_tuple(r) for r in ref.Vpart], [myread_tuple(r) for r in ref.Jpart])
    for chunk in (1024, 3000):
        c = Catt(chain_cfg=trb, chunk_reads=chunk)
        res = c.mainflow(path)
        assert ([myread_tuple(r) for r in res.Vpart], [myread_tuple(r) for r in res.Jpart]) == want, chunk
        assert res.info["n_reads"] == 7922 and res.info["the_err"] == ref.info["the_err"]
        # the in-memory path with the same tiny chunks (size ramp, many chunks)
        lines = open(path).read().split("\n")
        names = [lines[i][1:].split()[0] for i in range(0, len(lines) - 3, 4)]
        seqs = [lines[i + 1] for i in range(0, len(lines) - 3, 4)]
        quals = [lines[i + 3] for i in range(0, len(lines) - 3, 4)]
        res2 = c.mainflow_reads(names, seqs, quals)
        assert ([myread_tuple(r) for r in res2.Vpart], [myread_tuple(r) for r in res2.Jpart]) == want, chunk
        c.close()
    c = Catt(chain_cfg=trb, chunk_reads=1024)
    one = str(tmp_path / "one.fq")
    with open(one, "w") as fh:
        fh.write("\n".join(open(path).read().split("\n")[:4]) + "\n")
    r1 = c.mainflow(one)
    assert r1.info["n_reads"] == 1
    empty = str(tmp_path / "empty.fq")
    open(empty, "w").close()
    r0 = c.mainflow(empty)
    assert r0.info["n_reads"] == 0 and r0.Vpart == [] and r0.Jpart == []
    # paired files, small pieces: same as the oracle's per-file restatement
    lines = open(path).read().split("\n")
    recs = [(lines[i][1:].split()[0], lines[i + 1], lines[i + 3]) for i in range(0, len(lines) - 3, 4)]
    f1 = [r for r in recs if r[0].endswith("/1")][:2500]
    f2 = [r for r in recs if r[0].endswith("/2")][:2300]
    p1, p2 = str(tmp_path / "a_1.fq"), str(tmp_path / "a_2.fq")
    for p, f in ((p1, f1), (p2, f2)):
        with open(p, "w") as fh:
            fh.write("".join(f"@{n}\n{s}\n+\n{q}\n" for n, s, q in f))
    refp = O.run_mem_files(orefs[0], orefs[1], O.make_config(trb), [tuple(map(list, zip(*f1))), tuple(map(list, zip(*f2)))])
    rp = c.mainflow_paired(p1, p2)
    assert [myread_tuple(r) for r in rp.Vpart] == [myread_tuple(r) for r in refp.Vpart]
    assert [myread_tuple(r) for r in rp.Jpart] == [myread_tuple(r) for r in refp.Jpart]
    c.close()


def test_file_pipeline_sliding_pinned_window(trb, tmp_path):
    """Inputs larger than the pinned window pass through a sliding window on the host (forced here with CATT_PIN_WINDOW = 4 MiB on
    13 MB files): plain FASTQ, two paired files, and a .gz inflated by zlib - same lists as the whole-input mode."""
    import gzip
    import numpy as np
    from catt_b200.host import Catt
    c = Catt(chain_cfg=trb, chunk_reads=4096)
    n, L = 40000, 150
    name_buf, name_off, seq_buf, seq_off, qual_buf = c.synth_reads(n, L)
    seqs = np.frombuffer(seq_buf, dtype=np.uint8).reshape(n, L)
    quals = np.frombuffer(qual_buf, dtype=np.uint8).reshape(n, L)

    def write(path, lo, hi, tag):
        with open(path, "wb") as fh:
            for i in range(lo, hi):
                fh.write(b"@r%d%s extra words\n%s\n+\n%s\n" % (i, tag, seqs[i].tobytes(), quals[i].tobytes()))

    whole, m1, m2 = (str(tmp_path / x) for x in ("w.fq", "m_1.fq", "m_2.fq"))
    write(whole, 0, n, b"")
    write(m1, 0, n // 2 + 3000, b"/1")
    write(m2, 3000, n, b"/2")
    gz = str(tmp_path / "w.fq.gz")
    with gzip.open(gz, "wb", compresslevel=1) as fh:
        fh.write(open(whole, "rb").read())

    def run(fn):
        r = fn()
        out = ([myread_tuple(x) for x in r.Vpart], [myread_tuple(x) for x in r.Jpart], {k: r.info[k] for k in ("n_reads", "vpart", "jpart", "left", "rescued", "the_err")})
        r.close()
        return out

    base = {"plain": run(lambda: c.mainflow(whole)), "paired": run(lambda: c.mainflow_paired(m1, m2)), "gz": run(lambda: c.mainflow(gz))}
    assert base["plain"][2]["n_reads"] == n and base["plain"][0] == base["gz"][0] and len(base["plain"][0]) > 10000
    os.environ["CATT_PIN_WINDOW"] = str(4 << 20)
    os.environ["CATT_TRACE_WINDOW"] = "1"
    try:
        assert run(lambda: c.mainflow(whole)) == base["plain"]
        assert run(lambda: c.mainflow_paired(m1, m2)) == base["paired"]
        os.environ["CATT_ZLIB_ONLY"] = "1"               # inflated by zlib beforehand, then copied through the window block by block
        assert run(lambda: c.mainflow(gz)) == base["gz"]
    finally:
        for k in ("CATT_PIN_WINDOW", "CATT_ZLIB_ONLY", "CATT_TRACE_WINDOW"):
            os.environ.pop(k, None)
    c.close()


def test_one_million_reads_digest_parity(trb, O, orefs):
    """1 M synthetic 150-bp reads (several Stage A chunks, ~650 k Vpart reads, the k-mer table and extension at size): both full
    lists, the lists after filter!(cdr3 != "None"), every counter and a 4-rank GPU group against the oracle, compared through the
    canonical text digest (seq, qual, V name, J name, cdr3, able of every read, order sensitive).  profiles/parity_at_scale.py
    repeats this at 10 M reads and on five other chain / length configurations."""
    from catt_b200.host import Catt
    V, J = orefs
    n, length = 1_000_000, 150
    full = Catt(chain_cfg=trb, device=0)
    only = Catt(chain_cfg=trb, device=0, only_cdr3=True)
    grp = Catt(chain_cfg=trb, device=0, only_cdr3=True, ngpu=4, devices=[0, 0, 0, 0])
    try:
        bufs = full.synth_reads(n, length)
        seqs, quals = O.synth_reads(V, J, n, length)
        assert np.array_equal(seqs.reshape(-1), bufs[2]) and np.array_equal(quals.reshape(-1), bufs[4]), "the two generators disagree"
        ref = O.run_flat(V, J, O.make_config(trb), seqs, quals, threads=len(os.sched_getaffinity(0)))
        rf, ro, rg = full.mainflow_buffers(n, *bufs), only.mainflow_buffers(n, *bufs), grp.mainflow_buffers(n, *bufs)
        assert rf.text_digest() == ref.digest(False), "full lists differ from the oracle"
        assert ro.text_digest() == ref.digest(True), "lists after filter! differ from the oracle"
        assert rg.text_digest() == ref.digest(True) and rg.digest() == ro.digest(), "the 4-rank group differs"
        for k in ("kmer", "the_err", "v_hits", "j_hits", "v_final", "j_final", "share", "pair_seq_mismatch", "full_pushed", "vpart", "jpart",
                  "direct", "left", "rescued", "cand_v", "cand_j", "cells_v", "cells_j"):
            assert rf.info[k] == ref.info[k] == ro.info[k] == rg.info[k], (k, rf.info[k], ref.info[k], ro.info[k], rg.info[k])
        assert rf.info["vpart"] > 500_000 and rf.info["rescued"] > 50
    finally:
        full.close(); only.close(); grp.close()
