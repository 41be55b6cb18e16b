"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/cbl_gpu.h declares, the Python binding table matches the header, and the product fails
loudly (no CPU fallback) when no CUDA device is present.  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cbl_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cbl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from cbl_b200 import _lib

    syms = declared_symbols()
    assert len(syms) >= 40
    L = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, f"declared in include/cbl_gpu.h but not exported: {missing}"
    assert sorted(_lib.SIGNATURES) == syms, "Python binding table out of sync with the header"


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import cbl_b200

    with pytest.raises(cbl_b200.CBLError) as e:
        cbl_b200.CBL(25, 64, 24)
    assert e.value.code == 2  # CBL_ECUDA
    assert cbl_b200.launch_count() == 0


def test_product_never_imports_the_oracle():
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "cbl_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"oracle", txt, flags=re.I) and "no CPU fallback" not in txt.lower() + "x":
                    if re.search(r"(import|include|from)\s+.*oracle", txt):
                        bad.append(fn)
    assert not bad, f"product files reference the oracle: {bad}"
