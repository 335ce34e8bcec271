"""The C-ABI library loads on a CPU-only box and exports every symbol include/alens_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "alens_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(alens_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from alens_b200.capi import EXPORTS

    assert sorted(EXPORTS) == declared_symbols()


def test_library_exports_every_declared_symbol(alens_lib):
    for name in declared_symbols():
        assert hasattr(alens_lib.dll, name), name
    assert "sm_100a" in alens_lib.version()


def test_no_cpu_fallback(alens_lib):
    """Without a CUDA device context creation must fail loudly (no silent CPU path)."""
    import torch

    import alens_b200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(alens_b200.AlensError) as e:
        alens_b200.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    """Nothing under alens_b200/ or include/ may reference oracle/ (parity rule)."""
    bad = []
    for base in ("alens_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dp:
                continue
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                    with open(os.path.join(dp, fn)) as f:
                        txt = f.read()
                    if re.search(r"(import|from)\s+oracle|oracle/|liboracle|pyoracle", txt):
                        # docstrings that state the rule are fine; code references are not
                        for ln in txt.splitlines():
                            if re.search(r"(^\s*(import|from)\s+oracle)|liboracle|pyoracle\.", ln):
                                bad.append((fn, ln.strip()))
    assert not bad, bad


def test_block_layout_matches_reference_record():
    from alens_b200.capi import BLOCK_DTYPE

    off = {n: BLOCK_DTYPE.fields[n][1] for n in BLOCK_DTYPE.names}
    # SimToolbox/Constraint/ConstraintBlock.hpp:30-48 on x86-64
    assert (off["delta0"], off["gamma"], off["gammaLB"], off["gidI"], off["globalIndexJ"]) == (0, 8, 16, 24, 36)
    assert (off["oneSide"], off["bilateral"], off["kappa"], off["normI"], off["labJ"], off["stress"]) == \
        (40, 41, 48, 56, 176, 200)
    assert BLOCK_DTYPE.itemsize == 272
