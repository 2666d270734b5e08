"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/catt_b200.h declares, and
refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    import __graft_entry__ as g
    g.build()


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "catt_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(catt_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    _build()
    from catt_b200 import _lib
    L = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"libcatt_b200.so does not export {s}"
    assert sorted(_lib.EXPORTS) == syms


def test_struct_layouts_match_header():
    """The ctypes mirror against the C compiler's view of include/catt_b200.h: tests/abi_c/roundtrip.c (plain C11, built by
    build()) prints sizeof / offsetof of every public struct."""
    import subprocess
    _build()
    from catt_b200 import _lib
    out = subprocess.check_output([os.path.join(ROOT, "tests", "abi_c", "roundtrip"), "layout"], text=True)
    c_view = {k: int(v) for k, v in (ln.split() for ln in out.strip().split("\n"))}
    py = {"catt_config": _lib.Config, "catt_read": _lib.Read, "catt_info": _lib.Info, "catt_clonotype": _lib.Clonotype}
    seen = 0
    for name, val in c_view.items():
        kind, field = name.split(".")
        if kind == "sizeof":
            if field == "catt_aln":
                assert _lib.ALN_DTYPE.itemsize == val
            else:
                assert C.sizeof(py[field]) == val, (field, C.sizeof(py[field]), val)
        elif kind == "catt_aln":
            assert _lib.ALN_DTYPE.fields[field][1] == val, (name, _lib.ALN_DTYPE.fields[field][1], val)
        else:
            assert getattr(py[kind], field).offset == val, (name, getattr(py[kind], field).offset, val)
        seen += 1
    assert seen > 90
    # every ctypes field is covered by the C listing (padding / reserved fields aside)
    for kind, cls in py.items():
        for f, _ in cls._fields_:
            assert f.startswith(("reserved", "pad_")) or f"{kind}.{f}" in c_view, f"{kind}.{f} is not checked against the header"


@pytest.mark.gpu
def test_c_client_roundtrip(fx, trb):
    """A plain-C caller fills catt_config by value, runs the reference's sample through catt_run_file and walks catt_read -
    same counters and the same checksum over (seq, qual, V, J, cdr3) as the ctypes host; also through a 2-rank group."""
    import subprocess
    _build()
    from catt_b200.host import Catt
    fq = os.path.join(fx, "testSample.fq")
    ctx = Catt(chain_cfg=trb, device=0, only_cdr3=True, thread=2)
    res = ctx.mainflow(fq)
    h = 0xcbf29ce484222325

    def fnv(h, b):
        for x in b:
            h = ((h ^ x) * 0x100000001b3) & ((1 << 64) - 1)
        return h
    n_cdr3 = 0
    for part in (res.Vpart, res.Jpart):
        for r in part:
            for s in (r.seq, r.qual, r.vs, r.js):
                h = fnv(h, s.encode())
            if r.cdr3 != "None":
                h = fnv(h, r.cdr3.encode()); n_cdr3 += 1
    first = res.Vpart[0].cdr3 if res.Vpart else None
    want_d = ctx.d_list.index(next(d for d in ctx.d_list if d[0] == ctx.selBestMatchD([first])[0])) if first and ctx.selBestMatchD([first])[0] != "None" else -1
    exe = os.path.join(ROOT, "tests", "abi_c", "roundtrip")
    for ngpu in (1, 2):
        out = subprocess.check_output([exe, "run", os.path.join(fx, "resource"), fq, str(ngpu)], text=True).split()
        got = dict(zip(out[0::2], out[1::2]))
        assert int(got["reads"]) == res.info["n_reads"] == 7922
        assert int(got["kmer"]) == res.info["kmer"] and float(got["the_err"]) == pytest.approx(res.info["the_err"], abs=1e-12)
        assert (int(got["vpart"]), int(got["jpart"]), int(got["rescued"])) == (res.info["vpart"], res.info["jpart"], res.info["rescued"])
        assert int(got["returned_v"]) == len(res.Vpart) and int(got["returned_j"]) == len(res.Jpart) and int(got["with_cdr3"]) == n_cdr3
        assert int(got["best_d"]) == want_d
        assert int(got["n_ranks"]) == (0 if ngpu == 1 else ngpu)
        assert got["checksum"] == f"{h:016x}", (ngpu, got["checksum"], f"{h:016x}")
    ctx.close()


def test_no_cpu_fallback(trb):
    """Without a CUDA device the context refuses to exist: the product path must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _build()
    from catt_b200 import _lib
    from catt_b200.host import Catt
    with pytest.raises(_lib.CattError) as ei:
        Catt(chain_cfg=trb)
    assert ei.value.code == -3 and "no CPU fallback" in str(ei.value)


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under catt_b200/ may reference it."""
    for dp, _, fns in os.walk(os.path.join(ROOT, "catt_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{fn} mentions the oracle"
