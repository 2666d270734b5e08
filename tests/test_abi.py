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
    from catt_b200 import _lib
    assert _lib.ALN_DTYPE.itemsize == 32
    assert C.sizeof(_lib.Read) == 48
    assert C.sizeof(_lib.Info) == 16 + 8 + 20 * 8 + 8 * 8 + 2 * 8 + 5 * 8 + 2 * 8 + 8


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
