"""CPU: the C-ABI library loads without a GPU and exports every function include/mico_b200.h declares.
No compute calls here (there is no device); argument validation paths that return before touching CUDA are
exercised to check the error convention (negative code + mico_last_error message)."""
import ctypes as C
import os
import re

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "mico_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(mico_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from mico_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(_lib.lib, n)]
    assert not missing, f"declared in include/mico_b200.h but not exported: {missing}"


def test_version_and_error_convention():
    from mico_b200 import _lib
    assert _lib.lib.mico_version() >= 100
    rc = _lib.lib.mico_gemm_bf16(None, None)
    assert rc == -1
    assert b"invalid argument" in _lib.lib.mico_last_error()
    g = _lib.GemmArgs()
    g.a, g.b, g.out = 16, 16, 16
    g.M, g.N, g.K, g.lda, g.ldb = 128, 128, 60, 60, 60      # pitch not a multiple of 8 elements
    assert _lib.lib.mico_gemm_bf16(C.byref(g), None) == -1
    assert b"lda" in _lib.lib.mico_last_error()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under mico_b200/ may reference it."""
    pkg = os.path.join(REPO, "mico_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(root, f)
                assert "/root/reference" not in txt, os.path.join(root, f)
