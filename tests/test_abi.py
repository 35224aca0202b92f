"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol that
include/fnetgpu.h declares; without a GPU the product fails loudly (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "fnetgpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fnetgpu_[a-z_0-9]+)\s*\(", txt)))


def test_header_symbols_exported():
    from fortnet_b200._lib import lib, SYMBOLS
    l = lib()
    decl = _declared_symbols()
    assert len(decl) >= 20
    for s in decl:
        assert hasattr(l, s), "libfnetgpu.so does not export %s" % s
    assert sorted(SYMBOLS) == decl


def test_no_cpu_fallback():
    import torch
    import fortnet_b200 as fb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(fb.FnetGpuError):
        fb.Context()


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "fortnet_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), "%s references the oracle" % f
