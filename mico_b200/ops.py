"""Python call wrappers over the C-ABI (include/mico_b200.h).

PyTorch is used here for device memory, streams and shapes only: every function hands raw
device pointers to libmico_b200.so and raises if the library reports an error.  There is no
fallback path.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_GELU_BWD, ACT_NONE, ACT_QUICK_GELU,  # noqa: F401
                   ACT_QUICK_GELU_BWD, MicoError, check, lib)

BF16 = torch.bfloat16
F32 = torch.float32


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t, dtype, name):
    if not t.is_cuda:
        raise MicoError(f"{name}: expected a CUDA tensor (mico_b200 has no CPU path)")
    if t.dtype != dtype:
        raise MicoError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if t.dim() >= 1 and t.stride(-1) != 1:
        raise MicoError(f"{name}: innermost dimension must be contiguous")


def gemm(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=BF16, bias=None, residual=None,
         row_scale=None, rows_per_group=0, act=ACT_NONE, aux_out=None, aux_in=None,
         accumulate=False, alpha=1.0, remap=None, residual_bcast=False, out_rows=None):
    """out[M,N] = epilogue(alpha * A . B^T); see MicoGemmArgs in include/mico_b200.h.

    a: bf16 2-D. K-major [M,K] (a_mn=False) or MN-major [K,M] (a_mn=True).
    b: bf16 2-D. K-major [N,K] (b_mn=False) or MN-major [K,N] (b_mn=True).
    """
    _req(a, BF16, "a")
    _req(b, BF16, "b")
    if a.dim() != 2 or b.dim() != 2:
        raise MicoError("gemm operands must be 2-D")
    if a_mn:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    if K != Kb:
        raise MicoError(f"gemm: contraction mismatch {K} vs {Kb}")
    if out is None:
        rows = out_rows if out_rows is not None else M
        out = torch.empty((rows, N), device=a.device, dtype=out_dtype)
    g = _lib.GemmArgs()
    g.a, g.lda, g.a_mn_major = a.data_ptr(), a.stride(0), int(a_mn)
    g.b, g.ldb, g.b_mn_major = b.data_ptr(), b.stride(0), int(b_mn)
    g.M, g.N, g.K = M, N, K
    if out.dtype not in (BF16, F32):
        raise MicoError("gemm: out must be bf16 or fp32")
    g.out, g.ldo, g.out_fp32 = out.data_ptr(), out.stride(-2), int(out.dtype == F32)
    if bias is not None:
        _req(bias, F32, "bias")
        g.bias = bias.data_ptr()
    if residual is not None:
        _req(residual, F32, "residual")
        g.residual, g.ldr = residual.data_ptr(), residual.stride(-2)
    if row_scale is not None:
        _req(row_scale, F32, "row_scale")
        g.row_scale, g.rows_per_group = row_scale.data_ptr(), int(rows_per_group)
    g.act = int(act)
    if aux_out is not None:
        _req(aux_out, BF16, "aux_out")
        g.aux_out, g.ld_aux_out = aux_out.data_ptr(), aux_out.stride(-2)
    if aux_in is not None:
        _req(aux_in, BF16, "aux_in")
        g.aux_in, g.ld_aux_in = aux_in.data_ptr(), aux_in.stride(-2)
    g.accumulate = int(accumulate)
    g.alpha = float(alpha)
    if remap is not None:
        g.remap_gin, g.remap_gout, g.remap_off = remap
        g.residual_bcast = int(residual_bcast)
    check(lib.mico_gemm_bf16(C.byref(g), _stream()), "mico_gemm_bf16")
    return out
