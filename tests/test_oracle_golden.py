"""CPU tests: pin the oracle against every golden the reference ships for this path."""
import os

import pytest

from oracle import downstream as D
from oracle import oracle as O


@pytest.fixture(scope="module")
def refs(trb):
    return O.RefSet(trb["vregion"]), O.RefSet(trb["jregion"])


@pytest.fixture(scope="module")
def golden(fx):
    return D.load_golden_csv(os.path.join(fx, "aha.TRB.CDR3.CATT.csv"))


def run(fx, trb, refs, **kw):
    cfg = O.make_config(trb, **kw)
    R = O.run_file(refs[0], refs[1], cfg, os.path.join(fx, "testSample.fq"))
    tab = D.clonotype_table(R.Vpart, R.Jpart, R.info["the_err"], trb.get("Dregion"), O.best_d)
    return R, tab


def test_golden_csv_reproduced_exactly(fx, trb, refs, golden):
    """aha.TRB.CDR3.CATT.csv = reference output of testSample.fq: 721 rows, 752 reads, as a multiset of rows
    (allowance 0 in assignV is what the bundled golden was produced with)."""
    gold, gprob = golden
    R, tab = run(fx, trb, refs, allow_v=0)
    assert len(gold) == 721 and sum(gold.values()) == 752
    assert tab == gold
    ptab = D.load_prob(os.path.join(fx, "prob.csv"))
    for key in gold:
        assert abs(D.prob_of(key[0], ptab) - gprob[key]) < 1.1e-5


def test_log_checkpoints(fx, trb, refs):
    """the numbers the reference's own log lines print for testSample.fq."""
    R, _ = run(fx, trb, refs, allow_v=0)
    i = R.info
    assert (i["v_hits"], i["j_hits"]) == (2653, 1752)
    assert (i["v_gapped"], i["j_gapped"]) == (17, 1)
    assert (i["v_tied"], i["j_tied"]) == (1279, 15)
    assert (i["v_final"], i["j_final"], i["share"]) == (2462, 1747, 1110)
    assert (i["pair_seq_mismatch"], i["full_pushed"]) == (367, 708)
    assert (i["kmer"], i["avg_len"]) == (23, 105.0)
    assert abs(i["the_err"] - 0.00482) < 1e-5
    assert (i["vpart"], i["jpart"], i["direct"], i["left"], i["rescued"]) == (1035, 396, 750, 681, 2)
    assert (i["cand_v"], i["cand_j"]) == (54840, 4137)


def test_head_allowance(fx, trb, refs, golden):
    """HEAD calls real_score(rd, ref, 3) (catt.jl:100): 720 identical rows, 10 extra, one count 2 vs 1."""
    gold, _ = golden
    R, tab = run(fx, trb, refs, allow_v=3)
    assert sum(1 for k, n in tab.items() if gold.get(k) == n) == 720
    assert len([k for k in tab if k not in gold]) == 10
    assert (R.info["vpart"], R.info["direct"], R.info["left"], R.info["rescued"]) == (1080, 761, 715, 2)
    assert (len(tab), sum(tab.values())) == (731, 763)


@pytest.mark.parametrize("T", [2, 3, 4, 8])
def test_thread_order_emulation(fx, trb, refs, golden, T):
    """The NR % nthreads record order (catt.jl:228) must not change the golden table."""
    _, tab = run(fx, trb, refs, allow_v=0, nthreads=T)
    assert tab == golden[0]


def test_dregion_known_answers(fx, trb, golden):
    """721 known answers for selBestMatchD (Jtool.jl:40-52): gap(k) = -(4+k)."""
    for (aa, nn, v, j, d), _ in golden[0].items():
        assert O.best_d(nn, trb["Dregion"]) == d


def test_pac_matches_fasta_and_lrand48(fx):
    """The shipped bwa .pac decodes to the FASTA; IGHV's N's equal lrand48()&3 after srand48(11)."""
    from catt_b200.refg import REFG
    for sp in REFG.values():
        for ch in sp.values():
            for key in ("vregion", "jregion"):
                rs = O.RefSet(os.path.join(fx, "resource", ch["CDR3"][key]))
                n_from_pac, mism = rs.pac_stats()
                assert mism == 0, (ch["CDR3"][key], mism)
    ighv = O.RefSet(os.path.join(fx, "resource", REFG["hs"]["IGH"]["CDR3"]["vregion"]))
    assert ighv.pac_stats() == (11, 0)


def test_motifs_compile():
    from catt_b200.refg import REFG
    for sp in REFG.values():
        for ch in sp.values():
            c = ch["CDR3"]
            for k in ("cmotif", "fmotif", "innerC", "innerF"):
                n, ln = O.motif_expand(c[k])
                assert n >= 1 and ln >= 1
    assert O.motif_expand("(LR|YF|YI|YL|YQ|YR)C(A|S|T|V|G|R|P|D)") == (48, 4)
    assert O.motif_expand("[FTYH]{1}FG[ADENPQS]{1}G") == (1, 5)


def test_hash64_and_sw_basics():
    M = (1 << 64) - 1

    def h64(k):                      # bwa's hash_64, restated independently
        k = (k + (~(k << 32) & M)) & M
        k ^= k >> 22
        k = (k + (~(k << 13) & M)) & M
        k ^= k >> 8
        k = (k + (k << 3)) & M
        k ^= k >> 15
        k = (k + (~(k << 27) & M)) & M
        k ^= k >> 31
        return k
    for k in (0, 1, 2, 7921, 123456789, 2 ** 40 + 17):
        assert O.hash64(k) == h64(k)
    s, qi, tj = O.sw("ACGTACGTACGTAAAA", "TTACGTACGTACGTCC")
    assert (s, qi, tj) == (12, 11, 13)
    s, f = O.sw_facts("AAAAACCCCCGGGGGTTTTTACGTACGTAC" + "G" + "TTGCATTGCAAGGCCTTAAGGCCAATTGGC",
                      "AAAAACCCCCGGGGGTTTTTACGTACGTAC" + "TTGCATTGCAAGGCCTTAAGGCCAATTGGC")
    assert s == 60 - 7 and f["n_gap_ops"] == 1 and f["nm"] == 1 and f["mi"] == 61


def test_empty_and_short_inputs(trb, refs, tmp_path):
    p = tmp_path / "empty.fq"
    p.write_text("")
    R = O.run_file(refs[0], refs[1], O.make_config(trb), str(p))
    assert R.info["n_reads"] == 0 and R.Vpart == [] and R.Jpart == []
    p = tmp_path / "short.fq"
    p.write_text("@a\nACGT\n+\nIIII\n@b\nNNNNNNNNNNNNNNNNNNNNNNNN\n+\nIIIIIIIIIIIIIIIIIIIIIIII\n")
    R = O.run_file(refs[0], refs[1], O.make_config(trb), str(p))
    assert R.info["n_reads"] == 2 and R.info["v_hits"] == 0 and R.Vpart == []
