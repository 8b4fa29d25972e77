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


def test_new_entry_points_validate_before_touching_cuda():
    """Round-1 (f) entry points: bad arguments return MICO_ERR_INVALID_ARG (-1) with a message, on a box without a GPU."""
    from mico_b200 import _lib
    L = _lib.lib
    assert L.mico_set_reserved_sms(3) == -1 and b"reserved SMs" in L.mico_last_error()       # odd: not whole TPCs
    assert L.mico_set_reserved_sms(0) == 0
    f4 = (C.c_float * 4)(0.5, 0.5, 0.5, 0.5)
    assert L.mico_resize_normalize(None, 1, 1, 3, 8, 8, C.c_int64(192), None, 4, 4, f4, f4, 1, None) == -1
    assert L.mico_resize_normalize(C.c_void_p(16), 1, 1, 5, 8, 8, C.c_int64(320), C.c_void_p(16), 4, 4, f4, f4, 1, None) == -1   # C > 4
    assert L.mico_colsum2_bf16(C.c_void_p(16), C.c_int64(24), 8, 12, 0, 8, C.c_void_p(16), C.c_void_p(16), C.c_void_p(16),
                               C.c_size_t(1 << 20), None) == -1                                                  # n0 % 8 != 0
    assert b"invalid argument" in L.mico_last_error()
    assert L.mico_adamw_multi(C.c_void_p(16), C.c_void_p(16), C.c_void_p(16), 1, 16384, C.c_void_p(16), 17, C.c_float(1.0),
                              C.c_double(1.0), None) == -1                                                       # > 16 groups
    assert L.mico_layernorm_bwd_workspace(16448, 1408) >= 148 * 2 * 1408 * 4


def test_round2_late_entry_points_validate_before_touching_cuda():
    """asum_out (bias gradient from the weight-gradient GEMM) and the fused-dropout LayerNorm backward refuse what their kernels
    cannot do with MICO_ERR_UNSUPPORTED (-3) / MICO_ERR_INVALID_ARG (-1) and a message -- callers then take the separate passes."""
    from mico_b200 import _lib
    L = _lib.lib
    g = _lib.GemmArgs()
    g.a, g.b, g.out = 256, 256, 256
    g.M, g.N, g.K, g.lda, g.ldb, g.ldo = 512, 768, 64, 512, 768, 768          # N % 256 == 0: no room for the ones columns
    g.a_mn_major = g.b_mn_major = 1
    g.out_fp32, g.alpha = 1, 1.0
    g.asum_out, g.ones = 256, 256
    assert L.mico_gemm_bf16(C.byref(g), None) == -3 and b"asum_out" in L.mico_last_error()
    g.ones = 0
    assert L.mico_gemm_bf16(C.byref(g), None) == -1 and b"ones" in L.mico_last_error()
    p = C.c_void_p(256)
    z = C.c_int64(0)
    # D = 2048 is outside the block-per-row kernel: the dropout mask cannot be fused
    rc = L.mico_layernorm_bwd_dropout(p, 0, C.c_int64(2048), None, z, p, C.c_int64(2048), p, p, p, None, z, p, C.c_int64(2048), p,
                                      C.c_int64(2048), None, 0, p, p, 0, None, 8, 2048, p, C.c_size_t(1 << 30), C.c_float(0.1),
                                      C.c_uint64(1), C.c_uint64(0), None)
    assert rc == -3 and b"dropout" in L.mico_last_error()
    ks, rounds, ms = C.c_int(0), C.c_int(0), C.c_double(0)
    assert L.mico_gemm_plan(12, 5, 256, 128, 8, 74, 0, C.byref(ks), C.byref(rounds), C.byref(ms), None, 0) == -1   # 12 % 5 != 0


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under mico_b200/ may reference it."""
    pkg = os.path.join(REPO, "mico_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(root, f)
                assert "/root/reference" not in txt, os.path.join(root, f)
