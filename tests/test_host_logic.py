"""CPU tests of the pieces around the CUDA path: the oracle's paired-end restatement, the reference arm of bench.py (the JSON
contract the driver parses), and the chain tables mirrored from reference.jl."""
import json
import os
import subprocess
import sys

import pytest

from util import synth_reads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def test_oracle_multifile_reduces_to_single_file(O, trb):
    """run_mem_files with one file == run_mem; with two files the lists are per-file results joined before the pool pass."""
    V, J = O.RefSet(trb["vregion"]), O.RefSet(trb["jregion"])
    names, seqs, quals = synth_reads(1500, [V.record(i) for i in range(V.nrec)], [J.record(i) for i in range(J.nrec)], 100, seed=12)
    cfg = O.make_config(trb)
    one = O.run_mem(V, J, cfg, names, seqs, quals)
    same = O.run_mem_files(V, J, cfg, [(names, seqs, quals)])
    tup = lambda r: (r.seq, r.qual, r.vs, r.js, r.cdr3)
    assert [tup(r) for r in one.Vpart] == [tup(r) for r in same.Vpart] and [tup(r) for r in one.Jpart] == [tup(r) for r in same.Jpart]
    a, b = (names[:800], seqs[:800], quals[:800]), (names[800:], seqs[800:], quals[800:])
    two = O.run_mem_files(V, J, cfg, [a, b])
    # per-file alignment restarts the read ordinal (bwa's tie-break), so only file-independent facts are compared here
    assert two.info["v_hits"] == one.info["v_hits"] and two.info["j_hits"] == one.info["j_hits"]
    assert two.info["vpart"] + two.info["jpart"] > 0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints ONE JSON line with the keys the driver reads (CPU only, tiny sample)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-sample", "1500"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "reads_per_sec" and d["unit"] == "reads/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_chain_tables_cover_the_reference(fx):
    """refg.py mirrors reference.jl: every (species, chain) shipped in the fixture resolves to existing FASTA files and motifs."""
    from catt_b200.refg import chain_config
    for species, chain in [("hs", "TRB"), ("hs", "TRA"), ("hs", "IGH"), ("hs", "TRG"), ("hs", "TRD"), ("ms", "TRB"), ("pig", "TRB")]:
        cc = chain_config(species, chain, "CDR3", os.path.join(fx, "resource"))
        assert os.path.exists(cc["vregion"]) and os.path.exists(cc["jregion"])
        for k in ("cmotif", "fmotif", "innerC", "innerF"):
            assert isinstance(cc[k], str) and cc[k]
        assert isinstance(cc["coffset"], int) and isinstance(cc["foffset"], int)
