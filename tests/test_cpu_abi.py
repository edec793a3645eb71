"""CPU-side checks: the C-ABI library loads and exports every symbol include/nlv_b200.h declares."""
import ctypes
import os

import pytest


def test_library_exports_declared_symbols():
    from nlvsgg_b200 import _C
    if not os.path.exists(_C.SO_PATH):
        _C.build()
    lib = ctypes.CDLL(_C.SO_PATH)
    names = _C.declared_symbols()
    assert len(names) >= 5
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in nlv_b200.h but not exported: {missing}"
    lib.nlv_version.restype = ctypes.c_int
    assert lib.nlv_version() >= 100


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nlvsgg_b200")
    for d, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(d, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{f} imports the oracle"


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from nlvsgg_b200 import ops
    with pytest.raises(RuntimeError):
        ops.draw_union_boxes(torch.zeros(2, 8))
