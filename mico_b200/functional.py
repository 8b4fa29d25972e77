"""Autograd wrappers over the C-ABI kernels for the small ops around the towers (heads, losses, fusion inputs).

Each Function's forward/backward is a short launch sequence over mico_b200.ops; torch supplies buffers and the
autograd graph only.  The large stacks (ViT tower, BERT encoder) are single autograd nodes of their own
(eva_vit.py, bert.py); these wrappers serve mico.py's heads (mico.py:36-52, 386-403 in the reference) and the
losses (data/model/vast.py:394-462, 485-512).
"""
import torch

from . import ops
from .ops import ACT_GELU, ACT_GELU_BWD, ACT_NONE, BF16, F32, MicoError


def _flat2d(x):
    return x.reshape(-1, x.shape[-1])


class _LinearTC(torch.autograd.Function):
    """y = act(x W^T + b) on the tcgen05 GEMM (bf16 operands, fp32 accumulate).  x: [..., K] fp32 or bf16."""

    @staticmethod
    def forward(ctx, x, weight, bias, act, out_dtype, residual=None):
        shp = x.shape
        x2 = _flat2d(x)
        xb = x2 if x2.dtype == BF16 else ops.scale_cast_bf16(x2.contiguous())
        if not xb.is_contiguous():
            xb = xb.contiguous()
        wb = ops.cast_bf16(weight.detach().contiguous())
        pre = None
        if act == ACT_GELU:
            pre = torch.empty((xb.shape[0], weight.shape[0]), device=x.device, dtype=BF16)
        res2 = None
        if residual is not None:      # fused residual add (fp32) in the GEMM epilogue
            res2 = _flat2d(residual).contiguous().float()
            out_dtype = F32
        y = ops.gemm(xb, wb, bias=None if bias is None else bias.detach(), act=act, aux_out=pre, out_dtype=out_dtype,
                     residual=res2)
        ctx.save_for_backward(xb, wb, pre)
        ctx.meta = (shp, x.dtype, bias is not None, act)
        ctx.has_res = residual is not None
        return y.view(*shp[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xb, wb, pre = ctx.saved_tensors
        shp, xdtype, has_bias, act = ctx.meta
        dy2 = _flat2d(dy)
        dyb = dy2 if dy2.dtype == BF16 else ops.scale_cast_bf16(dy2.contiguous().float() if dy2.dtype != F32 else dy2.contiguous())
        if not dyb.is_contiguous():
            dyb = dyb.contiguous()
        if act != ACT_NONE:
            raise MicoError("fused activations are only used inside the tower launch sequences")
        dw = ops.gemm(dyb, xb, a_mn=True, b_mn=True, out_dtype=F32)
        db = ops.colsum(dyb) if has_bias else None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dyb, wb, b_mn=True, out_dtype=F32 if xdtype == F32 else BF16).view(shp)
        return dx, dw, db, None, None, (dy if ctx.has_res else None)


def linear_tc(x, weight, bias=None, out_dtype=F32, residual=None):
    return _LinearTC.apply(x, weight, bias, ACT_NONE, out_dtype, residual)


class _LinearF32(torch.autograd.Function):
    """y = x W^T + b in fp32 SIMT (small heads: Contra_head, Match_head, fused contra heads)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        shp = x.shape
        x2 = _flat2d(x).float()
        y = ops.sgemm(x2, weight.detach(), bias=None if bias is None else bias.detach())
        ctx.save_for_backward(x2, weight)
        ctx.meta = (shp, bias is not None)
        return y.view(*shp[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        shp, has_bias = ctx.meta
        dy2 = _flat2d(dy).float().contiguous()
        dw = ops.sgemm(dy2.t(), x2.t())                                                       # [N,K] = dy^T x
        db = None
        if has_bias:
            ones = torch.ones((1, dy2.shape[0]), device=dy2.device, dtype=F32)
            db = ops.sgemm(ones, dy2.t()).view(-1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.sgemm(dy2, weight.detach().t()).view(shp)                                # [M,K] = dy W
        return dx, dw, db


def linear_f32(x, weight, bias=None):
    return _LinearF32.apply(x, weight, bias)


def linear(x, weight, bias=None):
    """Generic nn.Linear replacement: tensor-core path for large row counts, fp32 SIMT for head-sized problems."""
    rows = x.numel() // x.shape[-1]
    if rows >= 256 and weight.shape[1] % 8 == 0 and weight.shape[0] % 8 == 0:
        return linear_tc(x, weight, bias)
    return linear_f32(x, weight, bias)


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype=F32):
        shp = x.shape
        x2 = _flat2d(x).contiguous()
        if x2.dtype != F32:
            x2 = x2.float()
        yb, y, mean, rstd = ops.layernorm_fwd(x2, weight.detach(), bias.detach(), eps, out_bf16=out_dtype == BF16,
                                              out_f32=out_dtype == F32)
        ctx.save_for_backward(x2, mean, rstd, weight)
        ctx.shp = shp
        return (yb if out_dtype == BF16 else y).view(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, weight = ctx.saved_tensors
        dy2 = _flat2d(dy).contiguous()
        if dy2.dtype not in (F32, BF16):
            dy2 = dy2.float()
        dg, db = torch.empty_like(weight), torch.empty_like(weight)
        dx, _ = ops.layernorm_bwd(dy2, x2, mean, rstd, weight.detach(), dg, db)
        return dx.view(ctx.shp), dg, db, None, None


def layer_norm(x, weight, bias, eps, out_dtype=F32):
    """LayerNorm over the last dim of an fp32 tensor; out_dtype=bf16 emits the next GEMM's operand directly."""
    return _LayerNorm.apply(x, weight, bias, eps, out_dtype)


class _GeluF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        xc = x.contiguous().float()
        ctx.save_for_backward(xc)
        return ops.gelu_f32(xc)

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        return ops.gelu_f32(xc, dy.contiguous().float())


def gelu(x):
    """exact-erf GELU (mico.py:22-33) on a small fp32 tensor"""
    return _GeluF32.apply(x)


class _L2Normalize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps):
        x2 = _flat2d(x).contiguous().float()
        y, norm = ops.l2norm_fwd(x2, eps)
        ctx.save_for_backward(y, norm)
        ctx.shp = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        y, norm = ctx.saved_tensors
        return ops.l2norm_bwd(y, _flat2d(dy).contiguous().float(), norm).view(ctx.shp), None


def normalize(x, eps=1e-12):
    """F.normalize(x, dim=-1)"""
    return _L2Normalize.apply(x, eps)


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index, label_smoothing, grad_dtype):
        l2 = _flat2d(logits)
        if l2.stride(-1) != 1:
            l2 = l2.contiguous()
        lab = labels.reshape(-1).contiguous()
        stats, lse = ops.cross_entropy_fwd(l2, lab, ignore_index, label_smoothing)
        ctx.save_for_backward(l2, lab, lse, stats)
        ctx.meta = (logits.shape, ignore_index, label_smoothing, grad_dtype)
        return stats[0].clone()

    @staticmethod
    def backward(ctx, g):
        l2, lab, lse, stats = ctx.saved_tensors
        shp, ignore_index, ls, grad_dtype = ctx.meta
        d = ops.cross_entropy_bwd(l2, lab, lse, g.contiguous().float().reshape(1), stats, ignore_index, ls,
                                  out_dtype=grad_dtype)
        return d.view(shp) if d.is_contiguous() else d.reshape(shp), None, None, None, None


def cross_entropy(logits, labels, ignore_index=-100, label_smoothing=0.0, grad_dtype=F32):
    """F.cross_entropy(logits, labels, ignore_index=, label_smoothing=) with mean reduction."""
    return _CrossEntropy.apply(logits, labels, ignore_index, label_smoothing, grad_dtype)


class _ContrastiveLogits(torch.autograd.Function):
    """sim = a . b_all^T / temp  (vast.py:405-408).  b_all comes from concat_all_gather (no gradient); a and temp do."""

    @staticmethod
    def forward(ctx, a, b_all, temp):
        a2, b2 = a.contiguous().float(), b_all.contiguous().float()
        sim = ops.sgemm(a2, b2, alpha_dev=temp.detach().reshape(1), alpha_recip=True)
        ctx.save_for_backward(a2, b2, temp, sim)
        return sim

    @staticmethod
    def backward(ctx, dsim):
        a2, b2, temp, sim = ctx.saved_tensors
        dsim = dsim.contiguous().float()
        t = temp.detach().reshape(1)
        da = ops.sgemm(dsim, b2.t(), alpha_dev=t, alpha_recip=True)                 # [M,K] = dsim . b_all / temp
        # d temp = - sum(dsim * sim) / temp   (sim is already divided by temp)
        dt = ops.dot(dsim, sim, alpha=-1.0)
        dt = ops.sgemm(dt.reshape(1, 1), torch.ones((1, 1), device=dt.device, dtype=F32), alpha_dev=t, alpha_recip=True)
        return da, None, dt.reshape(temp.shape)


def contrastive_logits(a, b_all, temp):
    return _ContrastiveLogits.apply(a, b_all, temp)


class _Attention(torch.autograd.Function):
    """o = softmax(scale * q k^T + mask) v on the fused tcgen05 kernels; q,k,v: bf16 [B,S,H,D] views (any strides)."""

    @staticmethod
    def forward(ctx, q, k, v, scale, mask):
        o, lse = ops.attention_fwd(q, k, v, scale, mask=mask, need_lse=True)
        ctx.save_for_backward(q, k, v, o, lse, mask if mask is not None else torch.empty(0))
        ctx.scale, ctx.has_mask = scale, mask is not None
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse, mask = ctx.saved_tensors
        do = do.contiguous() if do.dtype == BF16 else do.to(BF16).contiguous()
        dmask = torch.zeros_like(mask) if (ctx.has_mask and ctx.needs_input_grad[4]) else None
        dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, do, ctx.scale, mask=mask if ctx.has_mask else None, dmask=dmask)
        return dq, dk, dv, None, dmask


def attention(q, k, v, scale, mask=None):
    return _Attention.apply(q, k, v, scale, mask)


class _LinearGeluTC(torch.autograd.Function):
    """y = act(x W^T + b) with the activation fused in the GEMM epilogue; backward applies act' in an fp32 helper."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        shp = x.shape
        x2 = _flat2d(x)
        xb = x2 if x2.dtype == BF16 else ops.scale_cast_bf16(x2.contiguous())
        if not xb.is_contiguous():
            xb = xb.contiguous()
        wb = ops.cast_bf16(weight.detach().contiguous())
        pre = torch.empty((xb.shape[0], weight.shape[0]), device=x.device, dtype=BF16)
        y = ops.gemm(xb, wb, bias=None if bias is None else bias.detach(), act=act, aux_out=pre)
        ctx.save_for_backward(xb, wb, pre)
        ctx.meta = (shp, x.dtype, bias is not None, act)
        return y.view(*shp[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xb, wb, pre = ctx.saved_tensors
        shp, xdtype, has_bias, act = ctx.meta
        dyb = _flat2d(dy)
        if dyb.dtype != BF16:
            dyb = dyb.to(BF16)
        dyb = dyb.contiguous()
        # dpre = dy * act'(pre) = (dy . I) with the *_BWD epilogue: an identity GEMM would waste FLOPs, so use the
        # elementwise fp32 helper for erf-GELU (QuickGELU only occurs inside the CLIP tower's own launch sequence)
        if act != ACT_GELU:
            raise MicoError("linear_act: only exact-erf GELU is available outside the towers")
        dpre = ops.scale_cast_bf16(ops.gelu_f32(pre.float(), dyb.float()))
        dw = ops.gemm(dpre, xb, a_mn=True, b_mn=True, out_dtype=F32)
        db = ops.colsum(dpre) if has_bias else None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dpre, wb, b_mn=True, out_dtype=F32 if xdtype == F32 else BF16).view(shp)
        return dx, dw, db, None


def linear_gelu(x, weight, bias=None):
    """bf16 GELU(x W^T + b) on the tensor-core GEMM (returns bf16)."""
    return _LinearGeluTC.apply(x, weight, bias, ACT_GELU)


class _Rope(torch.autograd.Function):
    """Rotary embedding of q / k (EVA02: eva_vit_model.py:314-322): fp32 [B, T, H, d] -> rotated bf16 attention operand."""

    @staticmethod
    def forward(ctx, x, cos, sin):
        ctx.save_for_backward(cos, sin)
        return ops.rope(x.contiguous().float(), cos, sin)

    @staticmethod
    def backward(ctx, dy):
        cos, sin = ctx.saved_tensors
        dyb = dy.contiguous() if dy.dtype == BF16 else dy.to(BF16).contiguous()
        return ops.rope(dyb, cos, sin, inverse=True), None, None


def rope(x, cos, sin):
    return _Rope.apply(x, cos, sin)


class _SwiGLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u1, u2):
        a, b = u1.contiguous().float(), u2.contiguous().float()
        ctx.save_for_backward(a, b)
        return ops.swiglu(a, b)

    @staticmethod
    def backward(ctx, dg):
        a, b = ctx.saved_tensors
        return ops.swiglu(a, b, dg.contiguous().float())


def swiglu(u1, u2):
    """silu(u1) * u2 (eva_vit_model.py:201-224)"""
    return _SwiGLU.apply(u1, u2)


class _CastBF16(torch.autograd.Function):
    """fp32 -> bf16 operand with an fp32 gradient (v of the EVA02 attention)."""

    @staticmethod
    def forward(ctx, x):
        return ops.scale_cast_bf16(_flat2d(x).contiguous().float()).view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        return dy.float()


def cast_bf16(x):
    return _CastBF16.apply(x)


class _PatchEmbed(torch.autograd.Function):
    """Conv2d(k = s = P) as im2col + tcgen05 GEMM (eva_vit_model.py:440-447): pixels (B, C, H, W) -> (B, gh*gw, D) fp32."""

    @staticmethod
    def forward(ctx, x, weight, bias, P):
        B = x.shape[0]
        k = weight[0].numel()
        kpad = (k + 63) // 64 * 64
        cols = ops.patchify(x.contiguous().float(), P, kpad)
        wb = ops.cast_bf16_2d(weight.detach().reshape(weight.shape[0], -1), kpad)
        y = ops.gemm(cols, wb, out_dtype=F32, bias=None if bias is None else bias.detach())
        ctx.save_for_backward(cols)
        ctx.meta = (k, weight.shape, bias is not None)
        return y.view(B, -1, weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        (cols,) = ctx.saved_tensors
        k, wshape, has_bias = ctx.meta
        dyb = ops.scale_cast_bf16(_flat2d(dy).contiguous().float())
        dw = torch.empty((wshape[0], k), device=dy.device, dtype=F32)
        ops.gemm(dyb, cols[:, :k], a_mn=True, b_mn=True, out=dw)
        db = ops.colsum(dyb) if has_bias else None
        return None, dw.view(wshape), db, None


def patch_embed(x, weight, bias, P):
    return _PatchEmbed.apply(x, weight, bias, P)
