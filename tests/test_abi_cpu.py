"""CPU tests of the drop-in boundary: the shared library loads and exports every symbol declared in
include/uwu_b200.h; the ctypes signature table covers the whole header (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "uwu_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(uwu_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from uwudiff_b200 import build

    return build.build()


def test_header_symbols_exported(built_lib):
    L = ctypes.CDLL(built_lib)
    syms = _header_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/uwu_b200.h but not exported"


def test_ctypes_table_covers_header(built_lib):
    from uwudiff_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _header_symbols()
    L = _lib.lib()
    assert L.uwu_version() >= 100
    assert L.uwu_launch_count() == 0


def test_invalid_arguments_fail_loudly(built_lib):
    from uwudiff_b200 import _lib

    L = _lib.lib()
    d = _lib.GemmDesc()
    rc = L.uwu_gemm(ctypes.byref(d), None)
    assert rc == -1 and b"null operand" in L.uwu_last_error()
    with pytest.raises(_lib.UwuError):
        _lib.check(rc, "uwu_gemm")
    n = _lib.NoiseDesc()
    n.B, n.n_per = 2, 8
    assert L.uwu_noise_fwd(ctypes.byref(n), None) == -1
    assert L.uwu_wmse_workspace_floats(4, 1 << 20) == 4 * 128


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (no CPU fallback through the checker)."""
    pkg = os.path.join(ROOT, "uwudiff_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                assert "from oracle" not in txt and "import oracle" not in txt, f


def test_ops_refuse_cpu_tensors(built_lib):
    import torch

    from uwudiff_b200 import _lib, ops

    with pytest.raises(_lib.UwuError):
        ops.wmse_fwd(torch.zeros(2, 4), torch.zeros(2, 4), None)
