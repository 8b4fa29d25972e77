"""Python call wrappers over the C-ABI (include/mico_b200.h).

PyTorch is used here for device memory, streams and shapes only: every function hands raw
device pointers to libmico_b200.so and raises if the library reports an error.  There is no
fallback path.
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_GELU_BWD, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX, ACT_NONE, ACT_QUICK_GELU,  # noqa: F401
                   ACT_QUICK_GELU_BWD, ACT_QUICK_GELU_SAVE_GRAD, MicoError, check, lib)

BF16 = torch.bfloat16
F32 = torch.float32


_CHECK_IDS = os.environ.get("MICO_CHECK_IDS", "0") == "1"


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
         accumulate=False, alpha=1.0, remap=None, residual_bcast=False, out_rows=None, asum_out=None):
    """out[M,N] = epilogue(alpha * A . B^T); see MicoGemmArgs in include/mico_b200.h.

    a: bf16 2-D. K-major [M,K] (a_mn=False) or MN-major [K,M] (a_mn=True).
    b: bf16 2-D. K-major [N,K] (b_mn=False) or MN-major [K,N] (b_mn=True).
    asum_out (fp32 [M], weight gradients only, see gemm_fuses_asum): also receives sum_k A(m,k) = the bias gradient.
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
    if asum_out is not None:
        _req(asum_out, F32, "asum_out")
        if asum_out.numel() != M or not asum_out.is_contiguous():
            raise MicoError("gemm: asum_out must be a contiguous fp32 [M] tensor")
        g.asum_out, g.ones = asum_out.data_ptr(), _ones_rows(K, a.device).data_ptr()
    check(lib.mico_gemm_bf16(C.byref(g), _stream()), "mico_gemm_bf16")
    return out


_ONES = {}


def _ones_rows(K, device):
    """bf16 [>= K][64] of ones: the B tile behind the bias-gradient columns of a weight-gradient GEMM (MicoGemmArgs::ones)"""
    t = _ONES.get(device)
    if t is None or t.shape[0] < K:
        t = torch.ones((max(K, 1 << 18), 64), device=device, dtype=BF16)     # 32 MB: never regrown at the MiCo shapes
        _ONES[device] = t
    return t


def gemm_fuses_asum(M, N):
    """True when the weight gradient dW[M, N] = dY^T X can also return the bias gradient (asum_out) from the same pass: the
    last 256-wide N tile has room for the 32 ones columns (gemm.cu).  MICO_GEMM_ASUM=0 switches the fusion off (A/B)."""
    if os.environ.get("MICO_GEMM_ASUM", "1") == "0":
        return False
    return M >= 256 and N % 128 == 0 and 96 < N % 256 <= 160 and (N % 256) % 32 == 0


# ----------------------------------------------------------------------------- attention
def _bhsd_strides(t):
    """t: 4-D view indexed [b, s, h, d] (any strides, d contiguous) -> (ptr, bs, rs, hs)."""
    if t.dim() != 4 or t.stride(3) != 1:
        raise MicoError("attention operands must be 4-D [B,S,H,D] views with contiguous D")
    return t.data_ptr(), t.stride(0), t.stride(1), t.stride(2)


def kv_groups(kv_index, n_kv):
    """CSR inverse of kv_index (int32 [B] -> K/V entry): (ptr [n_kv + 1], list [B]) int32 device tensors -- K/V entry e is
    read by the query batch entries list[ptr[e]:ptr[e+1]] (MicoAttnArgs::grp_ptr / grp_list)."""
    idx = kv_index.long()
    order = torch.argsort(idx, stable=True).to(torch.int32)
    ptr = torch.zeros(n_kv + 1, device=kv_index.device, dtype=torch.int64)
    ptr[1:] = torch.cumsum(torch.bincount(idx, minlength=n_kv), 0)
    return ptr.to(torch.int32), order


def _attn_args(q, k, v, o, lse, mask, scale, dropout=None, kv_index=None, groups=None):
    """dropout: None or (p, seed) -- attention-probability dropout regenerated from the seed in the backward.
    kv_index: int32 [B] device tensor: query batch entry b attends to K/V batch entry kv_index[b] (k, v have fewer
    batch entries than q); groups = kv_groups(kv_index, k.shape[0]) for the backward pass."""
    a = _lib.AttnArgs()
    B, Sq, H, D = q.shape
    Sk = k.shape[1]
    if kv_index is not None:
        if kv_index.dtype != torch.int32 or kv_index.numel() != B or not kv_index.is_cuda or not kv_index.is_contiguous():
            raise MicoError("kv_index must be a contiguous int32 CUDA tensor with one entry per query batch entry")
        a.kv_index, a.n_kv = kv_index.data_ptr(), k.shape[0]
        if groups is not None:
            a.grp_ptr, a.grp_list = groups[0].data_ptr(), groups[1].data_ptr()
    elif k.shape[0] != B:
        raise MicoError("attention: q and k batch sizes differ and no kv_index was given")
    for name, t in (("q", q), ("k", k), ("v", v), ("o", o)):
        _req(t, BF16, name)
    a.q, a.q_bs, a.q_rs, a.q_hs = _bhsd_strides(q)
    a.k, a.k_bs, a.k_rs, a.k_hs = _bhsd_strides(k)
    a.v, a.v_bs, a.v_rs, a.v_hs = _bhsd_strides(v)
    a.o, a.o_bs, a.o_rs, a.o_hs = _bhsd_strides(o)
    if lse is not None:
        _req(lse, F32, "lse")
        a.lse = lse.data_ptr()
    if mask is not None:
        _req(mask, F32, "mask")
        if mask.dim() == 2:      # [B, Sk] key padding mask
            a.mask, a.mask_bs, a.mask_qs = mask.data_ptr(), mask.stride(0), 0
        elif mask.dim() == 3:    # [B, Sq, Sk]
            a.mask, a.mask_bs, a.mask_qs = mask.data_ptr(), mask.stride(0), mask.stride(1)
        elif mask.dim() == 4:    # [G, H, Sq, Sk]: per-head bias shared by batch entries b % G (Swin windows)
            if mask.shape[1] != H or B % mask.shape[0]:
                raise MicoError("4-D mask must be [G,H,Sq,Sk] with B % G == 0")
            a.mask, a.mask_bs, a.mask_hs, a.mask_qs = mask.data_ptr(), mask.stride(0), mask.stride(1), mask.stride(2)
            a.mask_bmod = mask.shape[0]
        else:
            raise MicoError("mask must be [B,Sk], [B,Sq,Sk] or [G,H,Sq,Sk] additive fp32")
    a.B, a.H, a.Sq, a.Sk, a.D = B, H, Sq, Sk, D
    a.scale = float(scale)
    if dropout is not None and dropout[0] > 0.0:
        a.dropout_p, a.dropout_seed = float(dropout[0]), int(dropout[1]) & (2 ** 64 - 1)
    return a


def attention_fwd(q, k, v, scale, mask=None, out=None, need_lse=True, dropout=None, kv_index=None):
    """q,k,v: bf16 views [B,S,H,D] (e.g. slices of a fused qkv buffer). Returns (o [B,Sq,H,D], lse [B,H,Sq])."""
    B, Sq, H, D = q.shape
    if out is None:
        out = torch.empty((B, Sq, H, D), device=q.device, dtype=BF16)
    lse = torch.empty((B, H, Sq), device=q.device, dtype=F32) if need_lse else None
    a = _attn_args(q, k, v, out, lse, mask, scale, dropout, kv_index)
    check(lib.mico_attention_fwd(C.byref(a), _stream()), "mico_attention_fwd")
    return out, lse


def attention_bwd(q, k, v, o, lse, dout, scale, mask=None, dq=None, dk=None, dv=None, dmask=None, dropout=None,
                  kv_index=None, groups=None):
    """dmask: optional zero-initialised fp32 tensor shaped like mask; receives the gradient of the additive bias.
    kv_index / groups: shared K/V entries (see _attn_args); dk, dv then have k.shape[0] batch entries."""
    B, Sq, H, D = q.shape
    Sk, Bk = k.shape[1], k.shape[0]
    dq = torch.empty((B, Sq, H, D), device=q.device, dtype=BF16) if dq is None else dq
    dk = torch.empty((Bk, Sk, H, D), device=q.device, dtype=BF16) if dk is None else dk
    dv = torch.empty((Bk, Sk, H, D), device=q.device, dtype=BF16) if dv is None else dv
    delta = torch.empty((B, H, Sq), device=q.device, dtype=F32)
    if kv_index is not None and groups is None:
        groups = kv_groups(kv_index, Bk)
    a = _attn_args(q, k, v, o, lse, mask, scale, dropout, kv_index, groups)
    _req(dout, BF16, "dout")
    a.dout, a.do_bs, a.do_rs, a.do_hs = _bhsd_strides(dout)
    a.delta = delta.data_ptr()
    a.dq, a.dq_bs, a.dq_rs, a.dq_hs = _bhsd_strides(dq)
    a.dk, a.dk_bs, a.dk_rs, a.dk_hs = _bhsd_strides(dk)
    a.dv, a.dv_bs, a.dv_rs, a.dv_hs = _bhsd_strides(dv)
    check(lib.mico_attention_bwd(C.byref(a), _stream()), "mico_attention_bwd")
    if dmask is not None:
        _req(dmask, F32, "dmask")
        if mask is None or dmask.shape != mask.shape or dmask.stride() != mask.stride():
            raise MicoError("dmask must have the layout of mask")
        check(lib.mico_attention_dmask(C.byref(a), _ptr(dmask), _stream()), "mico_attention_dmask")
    return dq, dk, dv


# ----------------------------------------------------------------------------- layernorm
_ws_cache = {}


def workspace(nbytes, device):
    """Grow-only scratch buffer per device (stream-ordered reuse on the current stream)."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), device=device, dtype=torch.uint8)
        _ws_cache[key] = buf
    return buf


def layernorm_fwd(x, gamma, beta, eps, out_bf16=True, out_f32=False, save_stats=True):
    """x: [M,D] fp32 or bf16.  Returns (y_bf16|None, y_f32|None, mean|None, rstd|None)."""
    M, D = x.shape
    if x.dtype not in (F32, BF16):
        raise MicoError("layernorm: x must be fp32 or bf16")
    _req(gamma, F32, "gamma")
    _req(beta, F32, "beta")
    yb = torch.empty((M, D), device=x.device, dtype=BF16) if out_bf16 else None
    yf = torch.empty((M, D), device=x.device, dtype=F32) if out_f32 else None
    mean = torch.empty(M, device=x.device, dtype=F32) if save_stats else None
    rstd = torch.empty(M, device=x.device, dtype=F32) if save_stats else None
    check(lib.mico_layernorm_fwd(_ptr(x), int(x.dtype == BF16), C.c_int64(x.stride(0)), _ptr(gamma), _ptr(beta),
                                 _ptr(yb), _ptr(yf), C.c_int64(D), _ptr(mean), _ptr(rstd), M, D, C.c_float(eps),
                                 _stream()), "mico_layernorm_fwd")
    return yb, yf, mean, rstd


def layernorm_bwd_fuses_colsum(D):
    """True when mico_layernorm_bwd can also emit the column sums of its scaled output (block-per-row kernel)."""
    return 512 <= D <= 1536 and D % 4 == 0


def layernorm_bwd_fuses_dropout(D, dy):
    """True when mico_layernorm_bwd can mask its bf16 output with the forward pass's hidden-dropout mask and sum its columns
    (block-per-row kernel, fp32 upstream gradient).  MICO_LN_BWD_DROPOUT=0 switches the fusion off (A/B)."""
    return layernorm_bwd_fuses_colsum(D) and dy.dtype == F32 and os.environ.get("MICO_LN_BWD_DROPOUT", "1") != "0"


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, *, dres=None, want_f32=True, want_bf16=False,
                  row_scale=None, rows_per_group=0, accumulate=False, dy2=None, colsum_out=None, dropout=None):
    """Returns (dx_f32|None, dx_bf16|None); writes dgamma/dbeta (fp32 [D]).  colsum_out (fp32 [D]): also receives the
    column sums of the scaled output (the upstream linear layer's bias gradient).  dropout = (p, seed, site_offset): the
    bf16 output and colsum_out carry the hidden-dropout mask of the forward pass (mico_layernorm_bwd_dropout; see
    layernorm_bwd_fuses_dropout)."""
    M, D = x.shape
    _req(x, F32, "x")
    dx = torch.empty((M, D), device=x.device, dtype=F32) if want_f32 else None
    dxb = torch.empty((M, D), device=x.device, dtype=BF16) if want_bf16 else None
    nws = lib.mico_layernorm_bwd_workspace(M, D)
    ws = workspace(nws, x.device)
    if dy2 is not None:
        _req(dy2, BF16, "dy2")
    if dropout is not None and dropout[0] > 0:
        p_, seed, site = dropout
        check(lib.mico_layernorm_bwd_dropout(
            _ptr(dy), int(dy.dtype == BF16), C.c_int64(dy.stride(0)), _ptr(dy2),
            C.c_int64(dy2.stride(0) if dy2 is not None else 0), _ptr(x), C.c_int64(x.stride(0)),
            _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dres),
            C.c_int64(dres.stride(0) if dres is not None else 0), _ptr(dx), C.c_int64(D), _ptr(dxb),
            C.c_int64(D), _ptr(row_scale), int(rows_per_group), _ptr(dgamma), _ptr(dbeta),
            int(accumulate), _ptr(colsum_out), M, D, _ptr(ws), C.c_size_t(ws.numel()), C.c_float(p_),
            C.c_uint64(int(seed) & (2 ** 64 - 1)), C.c_uint64(int(site)), _stream()), "mico_layernorm_bwd_dropout")
        return dx, dxb
    check(lib.mico_layernorm_bwd(_ptr(dy), int(dy.dtype == BF16), C.c_int64(dy.stride(0)), _ptr(dy2),
                                 C.c_int64(dy2.stride(0) if dy2 is not None else 0), _ptr(x), C.c_int64(x.stride(0)),
                                 _ptr(mean), _ptr(rstd), _ptr(gamma), _ptr(dres),
                                 C.c_int64(dres.stride(0) if dres is not None else 0), _ptr(dx), C.c_int64(D), _ptr(dxb),
                                 C.c_int64(D), _ptr(row_scale), int(rows_per_group), _ptr(dgamma), _ptr(dbeta),
                                 int(accumulate), _ptr(colsum_out), M, D, _ptr(ws), C.c_size_t(ws.numel()), _stream()),
          "mico_layernorm_bwd")
    return dx, dxb


# ----------------------------------------------------------------------------- helpers
def cast_bf16(src, dst=None):
    _req(src, F32, "src")
    if not src.is_contiguous():
        raise MicoError("cast_bf16: src must be contiguous")
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=BF16)
    check(lib.mico_cast_f32_to_bf16(_ptr(src), _ptr(dst), C.c_int64(src.numel()), _stream()), "mico_cast_f32_to_bf16")
    return dst


def cast_f32_from_bf16(src, dst):
    """bf16 contiguous -> fp32 contiguous (dst written in place)"""
    _req(src, BF16, "src")
    _req(dst, F32, "dst")
    if src.numel() != dst.numel() or not src.is_contiguous() or not dst.is_contiguous():
        raise MicoError("cast_f32_from_bf16: contiguous tensors of equal size expected")
    check(lib.mico_cast_bf16_to_f32(_ptr(src), _ptr(dst), C.c_int64(src.numel()), _stream()), "mico_cast_bf16_to_f32")
    return dst


def scale_(x, scale_dev=None, scale_host=1.0):
    """x *= scale_dev[0] * scale_host in place (fp32, contiguous, 16-byte aligned)."""
    _req(x, F32, "x")
    if not x.is_contiguous():
        raise MicoError("scale_: contiguous tensor expected")
    if x.numel() == 0:
        return x
    if scale_dev is not None:
        _req(scale_dev, F32, "scale_dev")
    check(lib.mico_scale_f32(_ptr(x), _ptr(scale_dev), C.c_float(scale_host), C.c_int64(x.numel()), _stream()), "mico_scale_f32")
    return x


def rope(x, cos, sin, inverse=False):
    """x: contiguous [B, T, H, d]; fp32 -> rotated bf16 (forward), or bf16 -> inverse-rotated fp32 (gradient)."""
    if not x.is_contiguous() or x.dim() != 4 or x.dtype not in (F32, BF16):
        raise MicoError("rope: contiguous fp32 / bf16 [B, T, H, d] expected")
    _req(cos, F32, "cos")
    _req(sin, F32, "sin")
    B, T, H, d = x.shape
    if tuple(cos.shape) != (T - 1, d) or tuple(sin.shape) != (T - 1, d) or not cos.is_contiguous() or not sin.is_contiguous():
        raise MicoError("rope: tables must be contiguous [T-1, d]")
    y = torch.empty(x.shape, device=x.device, dtype=F32 if x.dtype == BF16 else BF16)
    check(lib.mico_rope(_ptr(x), int(x.dtype == BF16), _ptr(y), int(y.dtype == BF16), _ptr(cos), _ptr(sin), B, T, H, d,
                        int(inverse), _stream()), "mico_rope")
    return y


def swiglu(u1, u2, dg=None):
    """fp32 contiguous: silu(u1) * u2, or (d u1, d u2) when dg is given."""
    for n_, t in (("u1", u1), ("u2", u2)):
        _req(t, F32, n_)
        if not t.is_contiguous():
            raise MicoError("swiglu: contiguous tensors expected")
    out0 = torch.empty_like(u1)
    out1 = torch.empty_like(u1) if dg is not None else None
    check(lib.mico_swiglu(_ptr(u1), _ptr(u2), _ptr(dg), _ptr(out0), _ptr(out1), C.c_int64(u1.numel()), _stream()), "mico_swiglu")
    return (out0, out1) if dg is not None else out0


def cast_bf16_2d(src, ldd):
    """fp32 [rows, cols] (row pitch src.stride(0)) -> bf16 [rows, ldd], zero-filled past cols."""
    _req(src, F32, "src")
    rows, cols = src.shape
    dst = torch.empty((rows, ldd), device=src.device, dtype=BF16)
    check(lib.mico_cast_f32_to_bf16_2d(_ptr(src), C.c_int64(src.stride(0)), rows, cols, _ptr(dst), C.c_int64(ldd),
                                       _stream()), "mico_cast_f32_to_bf16_2d")
    return dst


def colsum2(x, n0, gap, n1, out0, out1):
    """out0[c] = sum_m x[m, c] (c < n0); out1[c] = sum_m x[m, n0 + gap + c] (c < n1): one launch for two column ranges."""
    _req(x, BF16, "x")
    M = x.shape[0]
    ws = workspace(lib.mico_colsum_workspace(M, n0 + n1), x.device)
    check(lib.mico_colsum2_bf16(_ptr(x), C.c_int64(x.stride(0)), M, int(n0), int(gap), int(n1), _ptr(out0), _ptr(out1),
                                _ptr(ws), C.c_size_t(ws.numel()), _stream()), "mico_colsum2_bf16")


def colsum(x, out=None, accumulate=False):
    """out[n] = sum_m x[m,n]; accumulate: False/0 overwrite, True/1 add, 2 subtract."""
    _req(x, BF16, "x")
    M, N = x.shape
    if out is None:
        out = torch.empty(N, device=x.device, dtype=F32)
    ws = workspace(lib.mico_colsum_workspace(M, N), x.device)
    check(lib.mico_colsum_bf16(_ptr(x), C.c_int64(x.stride(0)), M, N, _ptr(out), int(accumulate), _ptr(ws),
                               C.c_size_t(ws.numel()), _stream()), "mico_colsum_bf16")
    return out


def batch_sum(x, B, out=None, accumulate=False):
    """x: fp32 contiguous, viewed as [B, R] -> out[R]."""
    _req(x, F32, "x")
    R = x.numel() // B
    if out is None:
        out = torch.empty(R, device=x.device, dtype=F32)
    check(lib.mico_batch_sum_f32(_ptr(x), B, C.c_int64(R), _ptr(out), int(accumulate), _stream()), "mico_batch_sum_f32")
    return out


def patchify(img, P, Kpad, replicate_channel=False, out=None, tokens_per_img=0, token_off=0):
    """img: fp32 [B,C,H,W] contiguous (or [B,H,W] with replicate_channel -> 3 identical channels).
    tokens_per_img > 0: image b's patches land at rows b*tokens_per_img + token_off.., other rows zeroed."""
    _req(img, F32, "img")
    if replicate_channel:
        B, H, W = img.shape
        Cc, img_stride, chan_stride = 3, img.stride(0), 0
    else:
        B, Cc, H, W = img.shape
        img_stride, chan_stride = img.stride(0), img.stride(1)
    if img.stride(-1) != 1 or img.stride(-2) != W:
        raise MicoError("patchify: image rows must be contiguous")
    rows = B * (tokens_per_img if tokens_per_img > 0 else (H // P) * (W // P))
    if out is None:
        out = torch.empty((rows, Kpad), device=img.device, dtype=BF16)
    check(lib.mico_patchify(_ptr(img), C.c_int64(img_stride), C.c_int64(chan_stride), B, Cc, H, W, P, Kpad,
                            int(tokens_per_img), int(token_off), _ptr(out), _stream()), "mico_patchify")
    return out


def cls_pos_row(cls_token, pos0, x, B, T, D):
    """x: fp32 [B*T, D]; writes rows b*T with cls_token + pos0."""
    check(lib.mico_cls_pos_row(_ptr(cls_token), _ptr(pos0), _ptr(x), C.c_int64(T * D), B, D, _stream()), "mico_cls_pos_row")


def scale_cast_bf16(x, row_scale=None, rows_per_group=0, out=None):
    _req(x, F32, "x")
    M, D = x.shape
    if out is None:
        out = torch.empty((M, D), device=x.device, dtype=BF16)
    check(lib.mico_scale_cast_bf16(_ptr(x), C.c_int64(x.stride(0)), _ptr(row_scale), int(rows_per_group), _ptr(out),
                                   C.c_int64(out.stride(0)), M, D, _stream()), "mico_scale_cast_bf16")
    return out


def drop_path_scales(drop_prob, B, seed, offset=0):
    """drop_prob: fp32 device [L] -> fp32 [L, 2, B] DropPath multipliers (mask / keep_prob)."""
    _req(drop_prob, F32, "drop_prob")
    L = drop_prob.numel()
    out = torch.empty((L, 2, B), device=drop_prob.device, dtype=F32)
    check(lib.mico_drop_path_scales(_ptr(drop_prob), L, B, C.c_uint64(seed), C.c_uint64(offset), _ptr(out), _stream()),
          "mico_drop_path_scales")
    return out


# ----------------------------------------------------------------------------- per-family device timing
PROF_KINDS = ("gemm", "attention_fwd", "attention_bwd", "layernorm_fwd", "layernorm_bwd", "other")


def profile_enable(on=True):
    check(lib.mico_profile_enable(int(bool(on))), "mico_profile_enable")


def profile_collect():
    """-> {family: dict(ms=, work=, calls=)}; synchronises the device."""
    n = len(PROF_KINDS)
    ms, work, cnt = (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
    check(lib.mico_profile_collect(ms, work, cnt, n), "mico_profile_collect")
    return {k: dict(ms=ms[i], work=work[i], calls=int(cnt[i])) for i, k in enumerate(PROF_KINDS)}


# ----------------------------------------------------------------------------- text head / loss kernels
I64 = torch.int64


def embedding_gather(ids, word, pos, typ, S, type_ids=None, pos_ids=None, pos_offset=0):
    """ids: int64 [M] -> fp32 [M, D] = word[ids] + type[type_ids|0] + pos[pos_ids | m % S + pos_offset]."""
    _req(ids, I64, "ids")
    M, D = ids.numel(), word.shape[1]
    # torch's nn.Embedding raises on an out-of-range index; the kernel clamps, so the position range is checked here (host
    # arithmetic only) and MICO_CHECK_IDS=1 adds the device-side range check of the token ids (one reduction + sync)
    if pos_ids is None and int(S) + int(pos_offset) > pos.shape[0]:
        raise MicoError(f"embedding_gather: sequence length {S} + offset {pos_offset} exceeds the {pos.shape[0]} position embeddings")
    if _CHECK_IDS and M:
        lo, hi = int(ids.min()), int(ids.max())
        if lo < 0 or hi >= word.shape[0]:
            raise MicoError(f"embedding_gather: token id range [{lo}, {hi}] outside the vocabulary of {word.shape[0]}")
    out = torch.empty((M, D), device=ids.device, dtype=F32)
    check(lib.mico_embedding_gather(_ptr(ids), _ptr(type_ids), _ptr(pos_ids), int(pos_offset), _ptr(word), _ptr(pos),
                                    _ptr(typ), _ptr(out), M, int(S), D, word.shape[0], pos.shape[0], typ.shape[0], _stream()),
          "mico_embedding_gather")
    return out


def embedding_scatter_add(dx, ids, dtable):
    _req(dx, F32, "dx")
    check(lib.mico_embedding_scatter_add(_ptr(dx), _ptr(ids), _ptr(dtable), dx.shape[0], dx.shape[1], dtable.shape[0],
                                         _stream()), "mico_embedding_scatter_add")


def cross_entropy_fwd(logits, labels, ignore_index=-100, label_smoothing=0.0):
    """-> (stats [2] = (loss, n_valid), lse [M])"""
    M, V = logits.shape
    _req(labels, I64, "labels")
    row_loss = torch.empty(M, device=logits.device, dtype=F32)
    lse = torch.empty(M, device=logits.device, dtype=F32)
    stats = torch.empty(2, device=logits.device, dtype=F32)
    check(lib.mico_cross_entropy_fwd(_ptr(logits), int(logits.dtype == BF16), C.c_int64(logits.stride(0)), _ptr(labels),
                                     C.c_int64(ignore_index), C.c_float(label_smoothing), _ptr(row_loss), _ptr(lse),
                                     _ptr(stats), M, V, _stream()), "mico_cross_entropy_fwd")
    return stats, lse


def cross_entropy_bwd(logits, labels, lse, grad, stats, ignore_index=-100, label_smoothing=0.0, out_dtype=F32):
    M, V = logits.shape
    ldd = (V + 7) // 8 * 8
    buf = torch.empty((M, ldd), device=logits.device, dtype=out_dtype)
    d = buf[:, :V]
    check(lib.mico_cross_entropy_bwd(_ptr(logits), int(logits.dtype == BF16), C.c_int64(logits.stride(0)), _ptr(labels),
                                     C.c_int64(ignore_index), C.c_float(label_smoothing), _ptr(lse), _ptr(grad), _ptr(stats),
                                     _ptr(d), int(out_dtype == BF16), C.c_int64(ldd), M, V, _stream()),
          "mico_cross_entropy_bwd")
    return d


def ce_chunk_update(logits, col0, labels, run_max, run_sum, label_logit, first):
    """one vocabulary chunk of the LM-head loss: logits fp32 [M, Vc] view (columns col0.. of the full vocabulary)"""
    _req(logits, F32, "logits")
    _req(labels, I64, "labels")
    M, Vc = logits.shape
    check(lib.mico_ce_chunk_update(_ptr(logits), C.c_int64(logits.stride(0)), int(col0), Vc, _ptr(labels), _ptr(run_max),
                                   _ptr(run_sum), _ptr(label_logit), M, int(bool(first)), _stream()), "mico_ce_chunk_update")


def ce_chunk_finalize(run_max, run_sum, label_logit, labels, V, ignore_index=-100):
    """-> (stats [2] = (loss, n_valid), lse [M])"""
    M = run_max.numel()
    row_loss = torch.empty(M, device=run_max.device, dtype=F32)
    lse = torch.empty(M, device=run_max.device, dtype=F32)
    stats = torch.empty(2, device=run_max.device, dtype=F32)
    check(lib.mico_ce_chunk_finalize(_ptr(run_max), _ptr(run_sum), _ptr(label_logit), _ptr(labels), C.c_int64(ignore_index),
                                     int(V), _ptr(row_loss), _ptr(lse), _ptr(stats), M, _stream()), "mico_ce_chunk_finalize")
    return stats, lse


def ce_chunk_grad(logits, col0, labels, lse, grad, stats, V, out, ignore_index=-100):
    """recomputed fp32 logits chunk [M, Vc] -> bf16 dlogits chunk written into out[:, :Vc]"""
    _req(logits, F32, "logits")
    _req(out, BF16, "out")
    M, Vc = logits.shape
    check(lib.mico_ce_chunk_grad(_ptr(logits), C.c_int64(logits.stride(0)), int(col0), Vc, _ptr(labels),
                                 C.c_int64(ignore_index), int(V), _ptr(lse), _ptr(grad), _ptr(stats), _ptr(out),
                                 C.c_int64(out.stride(0)), M, _stream()), "mico_ce_chunk_grad")


def l2norm_fwd(x, eps=1e-12):
    _req(x, F32, "x")
    M, D = x.shape
    y = torch.empty_like(x)
    norm = torch.empty(M, device=x.device, dtype=F32)
    check(lib.mico_l2norm_fwd(_ptr(x), _ptr(y), _ptr(norm), M, D, C.c_float(eps), _stream()), "mico_l2norm_fwd")
    return y, norm


def l2norm_bwd(y, dy, norm):
    M, D = y.shape
    dx = torch.empty_like(y)
    check(lib.mico_l2norm_bwd(_ptr(y), _ptr(dy), _ptr(norm), _ptr(dx), M, D, _stream()), "mico_l2norm_bwd")
    return dx


def sgemm(a, b, *, a_t=False, b_t=False, bias=None, alpha=1.0, alpha_dev=None, alpha_recip=False, out=None,
          accumulate=False):
    """fp32: out[M,N] (+)= s * A . B^T + bias.  a: [M,K] ([K,M] if a_t), b: [N,K] ([K,N] if b_t); any 2-D strides."""
    _req2 = lambda t, n: (_ for _ in ()).throw(MicoError(f"{n}: fp32 CUDA 2-D tensor expected")) \
        if (not t.is_cuda or t.dtype != F32 or t.dim() != 2) else None
    _req2(a, "a")
    _req2(b, "b")
    if a_t:
        K, M = a.shape
        a_sm, a_sk = a.stride(1), a.stride(0)
    else:
        M, K = a.shape
        a_sm, a_sk = a.stride(0), a.stride(1)
    if b_t:
        Kb, N = b.shape
        b_sn, b_sk = b.stride(1), b.stride(0)
    else:
        N, Kb = b.shape
        b_sn, b_sk = b.stride(0), b.stride(1)
    if K != Kb:
        raise MicoError(f"sgemm: contraction mismatch {K} vs {Kb}")
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=F32)
    check(lib.mico_sgemm_strided(_ptr(a), C.c_int64(a_sm), C.c_int64(a_sk), _ptr(b), C.c_int64(b_sn), C.c_int64(b_sk),
                                 _ptr(out), C.c_int64(out.stride(0)), _ptr(bias), M, N, K, C.c_float(alpha), _ptr(alpha_dev),
                                 int(alpha_recip), int(accumulate), _stream()), "mico_sgemm_strided")
    return out


def dot(a, b, alpha=1.0, out=None, accumulate=False):
    if out is None:
        out = torch.empty((), device=a.device, dtype=F32)
    check(lib.mico_dot_f32(_ptr(a), _ptr(b), C.c_int64(a.numel()), C.c_float(alpha), _ptr(out), int(accumulate), _stream()),
          "mico_dot_f32")
    return out


def gelu_f32(x, dy=None):
    _req(x, F32, "x")
    out = torch.empty_like(x)
    check(lib.mico_gelu_f32(_ptr(x), _ptr(dy), _ptr(out), C.c_int64(x.numel()), _stream()), "mico_gelu_f32")
    return out


def fbank(wave, window, mel, frame_shift=160, in_scale=32768.0, preemph=0.97, log_floor=1.1920928955078125e-07,
          norm_sub=0.0, norm_mul=1.0):
    """wave: fp32 [n_clips, n_samples] -> fp32 [n_clips, n_frames, num_mel] log-mel (Kaldi fbank semantics)."""
    _req(wave, F32, "wave")
    _req(window, F32, "window")
    _req(mel, F32, "mel")
    n_clips, n_samples = wave.shape
    frame_len = window.numel()
    if n_samples < frame_len:
        raise MicoError("fbank: clip shorter than one frame")
    n_frames = 1 + (n_samples - frame_len) // frame_shift
    num_mel = mel.shape[0]
    if mel.shape[1] != 257 or not mel.is_contiguous():
        raise MicoError("fbank: mel bank must be contiguous [num_mel, 257]")
    out = torch.empty((n_clips, n_frames, num_mel), device=wave.device, dtype=F32)
    check(lib.mico_fbank(_ptr(wave), C.c_int64(wave.stride(0)), n_clips, n_samples, frame_len, int(frame_shift), _ptr(window),
                         _ptr(mel), num_mel, C.c_float(in_scale), C.c_float(preemph), C.c_float(log_floor),
                         C.c_float(norm_sub), C.c_float(norm_mul), _ptr(out), C.c_int64(n_frames * num_mel), _stream()),
          "mico_fbank")
    return out


def dropout(x, p, seed, site_offset, res=None, out_f32=True, out_bf16=False):
    """[res +] x * Bernoulli(1-p)/(1-p) mask of (seed, site_offset + flat index).  x fp32 or bf16, contiguous."""
    if not x.is_contiguous() or (res is not None and not res.is_contiguous()):
        raise MicoError("dropout: contiguous tensors expected")
    yf = torch.empty(x.shape, device=x.device, dtype=F32) if out_f32 else None
    yb = torch.empty(x.shape, device=x.device, dtype=BF16) if out_bf16 else None
    check(lib.mico_dropout(_ptr(x), int(x.dtype == BF16), _ptr(res), _ptr(yf), _ptr(yb), C.c_int64(x.numel()), C.c_float(p),
                           C.c_uint64(int(seed) & (2 ** 64 - 1)), C.c_uint64(int(site_offset)), _stream()), "mico_dropout")
    return yf, yb
