"""CPU-only: the C-ABI shared libraries load and export every symbol include/aardvark_b200.h declares.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "aardvark_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:avk|orc)_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_boundary():
    names = _declared()
    for must in ("avk_create", "avk_set_reference", "avk_compare_batch", "avk_merge_batch", "avk_wfa_ed_batch", "avk_destroy"):
        assert must in names


def test_cuda_library_exports_every_declared_symbol():
    from aardvark_b200 import lib
    so = lib.build()
    h = ctypes.CDLL(so)
    for name in _declared():
        if name.startswith("avk_"):
            assert hasattr(h, name), f"{name} declared in include/aardvark_b200.h but not exported"
    assert set(lib.EXPORTED_SYMBOLS) == {n for n in _declared() if n.startswith("avk_")}


def test_oracle_library_exports_declared_symbols():
    import oracle_py
    h = oracle_py.lib()
    for name in _declared():
        if name.startswith("orc_"):
            assert hasattr(h, name)


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: creating a solver on a box without CUDA must raise, not degrade."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from aardvark_b200.lib import AvkError, Solver
    with pytest.raises(AvkError):
        Solver(0)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "aardvark_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(import|from)\s+\S*orac", src, flags=re.M), f"{fn} imports the oracle"
            assert "liboracle" not in src and "oracle_py" not in src, f"{fn} references the oracle library"
