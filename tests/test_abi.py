"""The C-ABI library loads and exports every symbol include/ncme.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ncme.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ncme_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(pkg):
    lib = pkg.load_library()
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/ncme.h but not exported: {missing}"


def test_binding_covers_header(pkg):
    from numcme_jl_b200 import _lib
    missing = [n for n in _declared() if n not in _lib.SIGNATURES]
    assert not missing, f"no ctypes signature for: {missing}"


def test_version(pkg):
    assert pkg.load_library().ncme_version() == 100


def test_fails_loudly_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.NcmeError, match="no CPU fallback"):
        pkg.Context(0)


def test_product_does_not_import_oracle():
    pkgdir = os.path.join(ROOT, "numcme.jl_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "oracle/" not in txt, f"{f} references oracle/"
