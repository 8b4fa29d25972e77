"""EVA-CLIP vision tower (ViT-g/14 family) on the sm_100a kernels.

Host-side mirror of the reference's ``EVAVisionTransformer`` (model/evaclip/eva_vit_model.py:488-659)
for the configuration MiCo uses (EVA01-CLIP-g-14: pre-norm blocks, no layer-scale, no RoPE, no relative
position bias, GELU MLP, q/v bias without k bias).  Same constructor keywords, same ``state_dict`` keys
and shapes, same ``forward(x, return_all_features)`` contract -- but the whole tower is ONE autograd node
whose forward and backward are explicit launch sequences over the C-ABI kernels (include/mico_b200.h):

  forward, per block (eva_vit_model.py:409-424):
    LN1 (fp32 residual stream -> bf16)                       mico_layernorm_fwd
    qkv = h . Wqkv^T + cat(q_bias, 0, v_bias)                mico_gemm_bf16           (eva:305-310)
    o   = softmax(q k^T / sqrt(d)) v  read in place from qkv mico_attention_fwd       (eva:340-361)
    x  += drop_path(o . Wproj^T + b)                         mico_gemm_bf16 epilogue  (eva:363, 420)
    LN2; a = GELU(h . W1^T + b1) (pre-activation kept)       mico_gemm_bf16 epilogue  (eva:191-193)
    x  += drop_path(a . W2^T + b2)                           mico_gemm_bf16 epilogue  (eva:197, 421)
  backward: the same kernels with MN-major operands (dgrad / wgrad need no transposes in HBM), the GELU
  derivative fused into the fc2-dgrad epilogue, the residual add and the DropPath scaling of the next
  branch gradient fused into the LayerNorm backward.

The residual stream is fp32; GEMM operands are bf16 with fp32 accumulation in TMEM.  There is no PyTorch
compute on this path: torch supplies device buffers, the current stream and the autograd hook.
"""
import math

import torch
import torch.nn as nn

from . import ops
from .ops import ACT_GELU_SAVE_GRAD, ACT_MUL_AUX, ACT_QUICK_GELU_SAVE_GRAD, BF16, F32, MicoError

_BLOCK_KEYS = ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.q_bias", "attn.v_bias", "attn.proj.weight",
               "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight",
               "mlp.fc2.bias")
_N1W, _N1B, _QKVW, _QB, _VB, _PW, _PB, _N2W, _N2B, _F1W, _F1B, _F2W, _F2B = range(13)
_NBLK = len(_BLOCK_KEYS)
# tower-level parameters come first in the flat list
_CLS, _POS, _PEW, _PEB, _NW, _NB, _LPW, _LPB = range(8)   # _LPW/_LPB: ln_pre of the OpenAI-CLIP tower (clip.py:245)
_NTOP = 8


def trunc_normal_(t, std=0.02):
    """timm-style truncated normal, cut at +-2 (absolute) like eva_vit_model.py:41-118 with a=-2, b=2."""
    return nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0)


class _ParamLinear(nn.Module):
    """Parameter holder with nn.Linear's state_dict layout; the math runs in the tower's launch sequence."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        trunc_normal_(self.weight, std=0.02)

    def forward(self, x):
        from .functional import linear
        return linear(x, self.weight, self.bias)


class LayerNorm(nn.Module):
    """LayerNorm parameter holder (evaclip/transformer.py:121-127); callable through the K2 kernel."""

    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.normalized_shape = (dim,)
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, x):
        from .functional import layer_norm
        return layer_norm(x, self.weight, self.bias, self.eps)


class PatchEmbed(nn.Module):
    """eva_vit_model.py:427-448: Conv2d(in_chans, embed_dim, k=s=patch).  Holds ``proj.weight/bias``."""

    class _Proj(nn.Module):
        def __init__(self, in_chans, embed_dim, patch):
            super().__init__()
            self.weight = nn.Parameter(torch.empty(embed_dim, in_chans, patch, patch))
            self.bias = nn.Parameter(torch.zeros(embed_dim))
            k = 1.0 / math.sqrt(in_chans * patch * patch)   # nn.Conv2d default init range
            nn.init.uniform_(self.weight, -k, k)
            nn.init.uniform_(self.bias, -k, k)

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.patch_shape = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.patch_shape[0] * self.patch_shape[1]
        self.proj = PatchEmbed._Proj(in_chans, embed_dim, patch_size)


class Attention(nn.Module):
    """Parameter holder for eva_vit_model.py:226-291 (qkv without bias + separate q_bias / v_bias)."""

    def __init__(self, dim, num_heads, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = _ParamLinear(dim, dim * 3, bias=False)
        if qkv_bias:
            self.q_bias = nn.Parameter(torch.zeros(dim))
            self.v_bias = nn.Parameter(torch.zeros(dim))
        else:
            raise MicoError("EVA tower without qkv bias is not a MiCo configuration")
        self.proj = _ParamLinear(dim, dim)


class Mlp(nn.Module):
    """Parameter holder for eva_vit_model.py:167-199 (subln=False: ffn_ln is Identity)."""

    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = _ParamLinear(in_features, hidden_features)
        self.fc2 = _ParamLinear(hidden_features, in_features)


class Block(nn.Module):
    """eva_vit_model.py:368-424 with gamma_1 = None, postnorm = False."""

    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, drop_path, eps):
        super().__init__()
        self.norm1 = LayerNorm(dim, eps)
        self.attn = Attention(dim, num_heads, qkv_bias)
        self.drop_prob = float(drop_path)
        self.norm2 = LayerNorm(dim, eps)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class _Bf16Cache:
    """bf16 copies of the fp32 master weights (the GEMM operands), refreshed when a parameter changes.
    The reference gets the same effect from torch.autocast (data/utils/pipeline.py:43)."""

    def __init__(self):
        self._c = {}

    def get(self, p, key, pad_to=None):
        ver = (p.data_ptr(), p._version, p.device)
        hit = self._c.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        src = p.detach()
        if pad_to is not None:
            w = ops.cast_bf16_2d(src.reshape(src.shape[0], -1), pad_to)
        else:
            w = ops.cast_bf16(src.contiguous())
        self._c[key] = (ver, w, p, pad_to is not None)
        return w

    def sinks(self):
        """(parameter, bf16 copy, key) of every unpadded entry: a fused optimizer step (mico_b200.optim.AdamW) writes the
        updated bf16 operand itself, then calls mark_fresh()."""
        return [(e[2], e[1], k) for k, e in self._c.items() if len(e) == 4 and not e[3] and e[1].numel() == e[2].numel()]

    def mark_fresh(self, updated=None):
        """Re-stamp the entries whose bf16 copy the fused optimizer step just wrote.  `updated`: ids of the parameters
        it actually updated (a parameter it skipped -- grad None, frozen -- may have been changed by someone else since
        the last cast and must be re-cast on its next use); None = every sink entry."""
        for k, e in list(self._c.items()):
            if len(e) == 4 and not e[3] and e[1].numel() == e[2].numel():
                p = e[2]
                if updated is not None and id(p) not in updated:
                    continue
                self._c[k] = ((p.data_ptr(), p._version, p.device), e[1], p, False)

    def qkv_bias(self, qb, vb, key):
        """cat(q_bias, zeros_like(v_bias), v_bias) (eva_vit_model.py:307), rebuilt when either changes."""
        ver = (qb.data_ptr(), qb._version, vb.data_ptr(), vb._version)
        hit = self._c.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        b = torch.cat((qb.detach(), torch.zeros_like(vb), vb.detach()))
        self._c[key] = (ver, b)
        return b


class _TowerFn(torch.autograd.Function):
    """The whole tower as one autograd node: (pixels, drop-path scales, parameters...) -> tokens."""

    @staticmethod
    def forward(ctx, tower, keep, x, dp_scales, *params):
        y, saved = tower._launch_forward(x, dp_scales, params, keep=keep)
        ctx.tower = tower
        ctx.saved = saved
        ctx.params = params
        ctx.dp_scales = dp_scales
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.saved is None:
            raise MicoError("tower backward called but activations were not kept")
        grads = ctx.tower._launch_backward(dy, ctx.saved, ctx.params, ctx.dp_scales)
        ctx.saved = None
        return (None, None, None, None) + tuple(grads)


class EVAVisionTransformer(nn.Module):
    """Drop-in for model/evaclip/eva_vit_model.py:488 (EVA01-g configuration)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., norm_layer=None, init_values=None, patch_dropout=0., use_abs_pos_emb=True,
                 use_rel_pos_bias=False, use_shared_rel_pos_bias=False, rope=False, use_mean_pooling=True,
                 init_scale=0.001, grad_checkpointing=False, xattn=False, postnorm=False, pt_hw_seq_len=16,
                 intp_freq=False, naiveswiglu=False, subln=False, eps=1e-6):
        super().__init__()
        unsupported = dict(init_values=init_values, patch_dropout=patch_dropout, use_rel_pos_bias=use_rel_pos_bias,
                           use_shared_rel_pos_bias=use_shared_rel_pos_bias, rope=rope, postnorm=postnorm,
                           naiveswiglu=naiveswiglu, subln=subln, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate,
                           qk_scale=qk_scale, use_mean_pooling=use_mean_pooling)
        bad = {k: v for k, v in unsupported.items() if v}
        if bad or not use_abs_pos_emb or not qkv_bias:
            raise NotImplementedError(f"EVA tower options outside the EVA01-g path: {bad}")
        if embed_dim % num_heads or (embed_dim // num_heads) % 8:
            raise NotImplementedError("head_dim must be a multiple of 8 (16-byte rows for TMA)")
        if norm_layer is not None:   # reference passes partial(LayerNorm, eps=1e-6) (evaclip/model.py:124)
            eps = getattr(norm_layer, "keywords", {}).get("eps", eps)
        self.image_size = img_size
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.eps = eps
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        num_patches = self.patch_embed.num_patches
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, embed_dim))
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth, device="cpu")]   # eva_vit_model.py:533
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, qkv_bias, dpr[i], eps)
                                     for i in range(depth)])
        self.norm = LayerNorm(embed_dim, eps)
        self.fc_norm = None
        self.head = _ParamLinear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        trunc_normal_(self.pos_embed, std=.02)
        trunc_normal_(self.cls_token, std=.02)
        with torch.no_grad():   # fix_init_weight (eva_vit_model.py:563-573)
            for i, blk in enumerate(self.blocks):
                blk.attn.proj.weight.div_(math.sqrt(2.0 * (i + 1)))
                blk.mlp.fc2.weight.div_(math.sqrt(2.0 * (i + 1)))
            if isinstance(self.head, _ParamLinear):
                self.head.weight.mul_(init_scale)
                self.head.bias.mul_(init_scale)
        self.grad_checkpointing = grad_checkpointing
        # with grad_checkpointing: what each block keeps besides its input (see _launch_forward) -- counted from the last
        # block backwards: n `light` blocks (qkv, o, lse, x1), then n `qkv` blocks (qkv, o, lse), then n `attn` blocks
        # (o, lse; -1 = every remaining block)
        self.ckpt_light_blocks = 0
        self.ckpt_qkv_blocks = 0
        self.ckpt_attn_blocks = 0
        self.flat_grad = None            # optional persistent fp32 gradient buffer (mico_b200.dp.FlatGrads)
        self._bf16 = _Bf16Cache()
        # K of the patch-embed GEMM padded to a multiple of 64 (one 128-byte swizzle atom of bf16)
        k = in_chans * patch_size * patch_size
        self._kpad = (k + 63) // 64 * 64
        self._injected_dp = None
        # launch-sequence variant (the OpenAI-CLIP tower in clip_vit.py flips these)
        self._act, self._act_bwd = ACT_GELU_SAVE_GRAD, ACT_MUL_AUX   # fc1 epilogue stores gelu'(x); fc2 dgrad multiplies by it
        self._full_qkv_bias = False      # EVA: cat(q_bias, 0, v_bias); CLIP: in_proj_bias [3D]
        self._ln_pre = False
        # "philox": all DropPath multipliers from one counter-based launch; "torch": the reference's own
        # bernoulli_ calls in the reference's order (bit-identical masks to a reference run with the same seed)
        self.drop_path_rng = "philox"
        # data-parallel hook: called during backward with each finished, contiguous slice of the flat fp32 gradient
        # buffer (one per block, last block first; then the tower-level parameters) so that a caller can overlap the
        # gradient reduction (data/utils/pipeline.py:93-99) with the rest of the backward pass
        self.grad_bucket_hook = None
        self.grad_begin_hook = None      # called when the tower's backward starts (see _launch_backward)
        self._dp_rates = None
        self._dp_calls = 0

    # ------------------------------------------------------------------ reference-compatible helpers
    def get_num_layers(self):
        return len(self.blocks)

    def get_cast_dtype(self):
        return self.blocks[0].mlp.fc2.weight.dtype

    def lock(self, unlocked_groups=0, freeze_bn_stats=False):
        assert unlocked_groups == 0, 'partial locking not currently supported for this model'
        for param in self.parameters():
            param.requires_grad = False

    def set_grad_checkpointing(self, enable=True):
        self.grad_checkpointing = enable

    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    # ------------------------------------------------------------------ parameters in launch order
    def _flat_params(self):
        top = [self.cls_token, self.pos_embed, self.patch_embed.proj.weight, self.patch_embed.proj.bias,
               self.norm.weight, self.norm.bias, None, None]
        for blk in self.blocks:
            top += [blk.norm1.weight, blk.norm1.bias, blk.attn.qkv.weight, blk.attn.q_bias, blk.attn.v_bias,
                    blk.attn.proj.weight, blk.attn.proj.bias, blk.norm2.weight, blk.norm2.bias,
                    blk.mlp.fc1.weight, blk.mlp.fc1.bias, blk.mlp.fc2.weight, blk.mlp.fc2.bias]
        return top

    @staticmethod
    def flat_grad_sizes(params):
        """elements reserved per parameter in the flat gradient buffer (16-byte aligned slices, launch order)"""
        return [((p.numel() if p is not None else 0) + 3) // 4 * 4 for p in params]

    _flat_grad_written = False
    # class-level defaults (the OpenAI-CLIP subclass builds itself without this __init__)
    flat_grad = None
    ckpt_light_blocks = 0
    ckpt_qkv_blocks = 0
    ckpt_attn_blocks = 0
    grad_begin_hook = None

    def invalidate_weight_cache(self):
        """Drop the bf16 operand copies (an optimizer step that bypasses tensor versioning, or a benchmark
        that wants the per-step cast of changed weights inside the timed region)."""
        self._bf16 = _Bf16Cache()

    def bf16_weight_sinks(self):
        return self._bf16.sinks()

    def mark_weights_fresh(self, updated=None):
        self._bf16.mark_fresh(updated)

    def inject_drop_path_scales(self, scales):
        """Parity hook: use these (depth, 2, B) DropPath multipliers (mask / keep_prob) for the next
        training-mode forward instead of drawing them."""
        self._injected_dp = scales

    def _draw_drop_path(self, B, device):
        """DropPath draws in the reference's order -- attention branch then MLP branch of each block with
        p > 0, ``x.new_empty((B,1,1)).bernoulli_(keep)`` then ``div_(keep)`` (eva_vit_model.py:121-138) --
        so that a run seeded like the reference sees the same masks."""
        if self._injected_dp is not None:
            dp, self._injected_dp = self._injected_dp, None
            return dp.to(device=device, dtype=F32).contiguous()
        if not self.training or all(b.drop_prob == 0.0 for b in self.blocks):
            return None
        if self.drop_path_rng == "philox":   # one launch for the whole tower (K12)
            if self._dp_rates is None or self._dp_rates.device != device:
                self._dp_rates = torch.tensor([b.drop_prob for b in self.blocks], device=device, dtype=F32)
            self._dp_calls += 1
            return ops.drop_path_scales(self._dp_rates, B, torch.initial_seed() & (2 ** 63 - 1), self._dp_calls)
        dp = torch.ones(len(self.blocks), 2, B, device=device, dtype=F32)
        for i, blk in enumerate(self.blocks):
            if blk.drop_prob > 0.0:
                keep = 1.0 - blk.drop_prob
                for j in range(2):
                    m = torch.empty((B, 1, 1), device=device, dtype=F32).bernoulli_(keep)
                    if keep > 0.0:
                        m.div_(keep)
                    dp[i, j] = m.view(B)
        return dp

    # ------------------------------------------------------------------ forward / backward launch sequences
    def _ckpt_levels(self, L):
        """Checkpoint level per block (0 = input only ... 3 = light), assigned from the last block backwards."""
        n3 = min(L, max(0, int(self.ckpt_light_blocks)))
        n2 = min(L - n3, max(0, int(self.ckpt_qkv_blocks)))
        n1 = int(self.ckpt_attn_blocks)
        n1 = L - n3 - n2 if n1 < 0 else min(L - n3 - n2, n1)
        return [0] * (L - n3 - n2 - n1) + [1] * n1 + [2] * n2 + [3] * n3

    def _block_forward(self, xr, i, params, dp, B, T, keep, level=0):
        """One pre-norm block (eva_vit_model.py:409-424) as a launch sequence; returns (x_out, tensors kept for backward).
        keep = False with a checkpoint level > 0: nothing but the level's tensors survives the call -- ("ck", xr, qkv, o,
        lse, x1) with None for what the level drops (1: o, lse; 2: + qkv; 3: + x1)."""
        D, H = self.embed_dim, self.num_heads
        d = D // H
        M = B * T
        c = self._bf16
        scale = d ** -0.5
        base = _NTOP + i * _NBLK
        p = [t.detach() for t in params[base:base + _NBLK]]
        h, _, mean1, rstd1 = ops.layernorm_fwd(xr, p[_N1W], p[_N1B], self.eps, save_stats=keep)
        if self._full_qkv_bias:
            qkv_bias = params[base + _QB].detach()
        else:
            qkv_bias = c.qkv_bias(params[base + _QB], params[base + _VB], ("qkvb", i))
        qkv = ops.gemm(h, c.get(params[base + _QKVW], ("qkv", i)), bias=qkv_bias)
        qkv5 = qkv.view(B, T, 3, H, d)
        o, lse = ops.attention_fwd(qkv5[:, :, 0], qkv5[:, :, 1], qkv5[:, :, 2], scale, need_lse=keep or level > 0)
        s_attn = dp[i, 0] if dp is not None else None
        s_mlp = dp[i, 1] if dp is not None else None
        x1 = ops.gemm(o.view(M, D), c.get(params[base + _PW], ("proj", i)), out_dtype=F32, bias=p[_PB], residual=xr,
                      row_scale=s_attn, rows_per_group=T)
        h2, _, mean2, rstd2 = ops.layernorm_fwd(x1, p[_N2W], p[_N2B], self.eps, save_stats=keep)
        w1 = c.get(params[base + _F1W], ("fc1", i))
        pre = torch.empty((M, w1.shape[0]), device=xr.device, dtype=BF16) if keep else None
        a = ops.gemm(h2, w1, bias=p[_F1B], act=self._act, aux_out=pre)
        x2 = ops.gemm(a, c.get(params[base + _F2W], ("fc2", i)), out_dtype=F32, bias=p[_F2B], residual=x1,
                      row_scale=s_mlp, rows_per_group=T)
        if keep:
            rec = (xr, mean1, rstd1, h, qkv, o, lse, x1, mean2, rstd2, h2, pre, a)
        elif level > 0:
            rec = ("ck", xr, qkv if level >= 2 else None, o, lse, x1 if level >= 3 else None)
        else:
            rec = None
        return x2, rec

    def _block_rebuild(self, rec, i, params, dp, B, T):
        """Rebuild the 13-tensor block record from a checkpointed one (see _launch_forward): whatever the block's level
        dropped is recomputed from what it kept with the forward pass's own launches, so every tensor is bit-identical to
        the forward pass.  The attention kernel never runs again (o and lse are kept from level 1 on)."""
        _, xr, qkv, o, lse, x1 = rec
        D, H = self.embed_dim, self.num_heads
        d = D // H
        c = self._bf16
        base = _NTOP + i * _NBLK
        p = [t.detach() for t in params[base:base + _NBLK]]
        h, _, mean1, rstd1 = ops.layernorm_fwd(xr, p[_N1W], p[_N1B], self.eps, save_stats=True)
        if qkv is None:
            if self._full_qkv_bias:
                qkv_bias = params[base + _QB].detach()
            else:
                qkv_bias = c.qkv_bias(params[base + _QB], params[base + _VB], ("qkvb", i))
            qkv = ops.gemm(h, c.get(params[base + _QKVW], ("qkv", i)), bias=qkv_bias)
        if x1 is None:
            x1 = ops.gemm(o.view(B * T, D), c.get(params[base + _PW], ("proj", i)), out_dtype=F32, bias=p[_PB], residual=xr,
                          row_scale=dp[i, 0] if dp is not None else None, rows_per_group=T)
        h2, _, mean2, rstd2 = ops.layernorm_fwd(x1, p[_N2W], p[_N2B], self.eps, save_stats=True)
        w1 = self._bf16.get(params[base + _F1W], ("fc1", i))
        pre = torch.empty((h2.shape[0], w1.shape[0]), device=xr.device, dtype=BF16)
        a = ops.gemm(h2, w1, bias=p[_F1B], act=self._act, aux_out=pre)
        return (xr, mean1, rstd1, h, qkv, o, lse, x1, mean2, rstd2, h2, pre, a)

    def _launch_forward(self, x, dp, params, keep):
        P = self.patch_embed.patch_size[0]
        D, H = self.embed_dim, self.num_heads
        d = D // H
        xs = x if isinstance(x, (tuple, list)) else (x,)
        B = sum(t.shape[0] for t in xs)
        T = self.patch_embed.num_patches + 1
        M = B * T
        dev = xs[0].device
        c = self._bf16
        # K1 patch embedding: im2col -> GEMM with bias + broadcast pos_embed in the epilogue; row 0 = cls + pos[0]
        # a 3-D input (B,H,W) is one channel replicated three times (forward_audio_encoder, mico.py:139-143): the
        # im2col kernel reads the same plane for every channel instead of materialising repeat(1,1,3,1,1).
        # Several pixel batches (video frames, spectrogram planes, depth maps of one omni-modal step) share one pass:
        # each is unfolded into its own row range of the same operand.
        cols = torch.empty((M, self._kpad), device=dev, dtype=BF16)
        r0 = 0
        for t in xs:
            ops.patchify(t, P, self._kpad, tokens_per_img=T, token_off=1, replicate_channel=(t.dim() == 3),
                         out=cols[r0 * T:(r0 + t.shape[0]) * T])
            r0 += t.shape[0]
        w_pe = c.get(params[_PEW], "pe", pad_to=self._kpad)
        pos = params[_POS].detach().reshape(T, D)
        pe_bias = params[_PEB].detach() if params[_PEB].numel() else None      # clip.py:239 conv1 has no bias
        xr = ops.gemm(cols, w_pe, out_dtype=F32, bias=pe_bias, residual=pos, remap=(T, T, 0), residual_bcast=True)
        ops.cls_pos_row(params[_CLS].detach().reshape(-1), pos, xr, B, T, D)
        saved = {"cols": cols, "B": B, "blocks": []} if keep else None
        if self._ln_pre:     # clip.py:283: x = ln_pre(x) becomes the residual stream
            x0 = xr
            _, xr, m0, r0 = ops.layernorm_fwd(x0, params[_LPW].detach(), params[_LPB].detach(), self.eps, out_bf16=False,
                                              out_f32=True, save_stats=keep)
            if keep:
                saved["ln_pre"] = (x0, m0, r0)
        # eva_vit_model.py:635-637: with grad_checkpointing keep only each block's input and recompute the block in the
        # backward pass.  Memory permitting, blocks keep more than that (HBM traded for recompute, most valuable first):
        #   level 1 `attn`  + attention output and log-sum-exp (2.2 KB per token): the attention kernel is not re-run
        #   level 2 `qkv`   + qkv (8.4 KB per token): LN1 is re-run, the qkv GEMM is not
        #   level 3 `light` + x1 (5.6 KB per token): the proj GEMM is not re-run; LN2 and fc1 + GELU always are
        L = len(self.blocks)
        ckpt = keep and self.grad_checkpointing
        levels = self._ckpt_levels(L) if ckpt else None
        for i in range(L):
            x2, rec = self._block_forward(xr, i, params, dp, B, T, keep and not ckpt, levels[i] if ckpt else 0)
            if keep:
                saved["blocks"].append((xr,) if (ckpt and levels[i] == 0) else rec)
            xr = x2
        _, y, mean, rstd = ops.layernorm_fwd(xr, params[_NW].detach(), params[_NB].detach(), self.eps,
                                             out_bf16=False, out_f32=True, save_stats=keep)
        if keep:
            saved["final"] = (xr, mean, rstd)
        return y.view(B, T, D), saved

    def _launch_backward(self, dy, saved, params, dp):
        D, H = self.embed_dim, self.num_heads
        d = D // H
        B = saved["B"]
        T = self.patch_embed.num_patches + 1
        M = B * T
        L = len(self.blocks)
        dev = dy.device
        c = self._bf16
        scale = d ** -0.5
        grads = [None] * len(params)
        # one flat fp32 gradient buffer in parameter order (block i's gradients are contiguous: a data-parallel
        # caller can reduce them bucket by bucket while earlier blocks are still in backward)
        sizes = self.flat_grad_sizes(params)
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + n)
        # A data-parallel / optimizer layer may own the buffer (mico_b200.dp.FlatGrads): parameters' .grad are then
        # persistent views into it, this pass WRITES them in place (one tower backward per zero_grad) and autograd
        # receives no parameter gradients -- so a bucket all-reduce issued from the hook below really reduces p.grad
        # (with autograd-owned gradients AccumulateGrad may clone the returned views).
        owned = self.flat_grad is not None
        if owned:
            flat = self.flat_grad
            if flat.numel() != offs[-1] or flat.device != dev:
                raise MicoError("tower.flat_grad does not match the parameter layout (use mico_b200.dp.FlatGrads)")
            if self._flat_grad_written:
                raise MicoError("flat-gradient mode takes ONE tower backward per zero_grad (fold all modalities into one "
                                "forward_multi call)")
            self._flat_grad_written = True
        else:
            flat = torch.empty(offs[-1], device=dev, dtype=F32)
        self._last_flat_grad = (flat, offs)

        def pgrad(idx):
            g = flat[offs[idx]:offs[idx] + params[idx].numel()].view(params[idx].shape)
            if not owned:
                grads[idx] = g
            return g

        def branch_scale(i, j):
            return dp[i, j] if (dp is not None and i >= 0) else None

        if self.grad_begin_hook is not None:
            # every autograd node created after the tower's forward has run by now (the engine orders ready nodes by
            # creation sequence, newest first): all gradients outside the tower are final -- a caller can start reducing them
            self.grad_begin_hook()
        dy = dy.contiguous().view(M, D)
        if dy.dtype != F32:
            dy = dy.float()
        xl, mean, rstd = saved.pop("final")
        # gradient wrt the residual stream after the last block, plus its bf16 DropPath-scaled copy that
        # enters the last MLP branch
        # the bias gradient of the linear layer that consumes a LayerNorm-backward output (column sums of the scaled bf16
        # copy) is produced by the same kernel when the width allows it
        fuse_cs = ops.layernorm_bwd_fuses_colsum(D)
        dx, dxb = ops.layernorm_bwd(dy, xl, mean, rstd, params[_NW].detach(), pgrad(_NW), pgrad(_NB),
                                    want_bf16=True, row_scale=branch_scale(L - 1, 1), rows_per_group=T,
                                    colsum_out=pgrad(_NTOP + (L - 1) * _NBLK + _F2B) if (fuse_cs and L > 0) else None)
        del xl, mean, rstd
        blocks = saved["blocks"]
        for i in range(L - 1, -1, -1):
            rec = blocks.pop()
            if len(rec) == 1:        # checkpointed: recompute this block's forward from its input
                rec = self._block_forward(rec[0], i, params, dp, B, T, True)[1]
            elif isinstance(rec[0], str):   # level 1-3 checkpoint: recompute what the level dropped (never the attention)
                rec = self._block_rebuild(rec, i, params, dp, B, T)
            xr, mean1, rstd1, h, qkv, o, lse, x1, mean2, rstd2, h2, pre, a = rec
            del rec
            base = _NTOP + i * _NBLK
            p = [t.detach() for t in params[base:base + _NBLK]]
            # ---- MLP branch: x2 = x1 + s * (a W2^T + b2)
            ops.gemm(dxb, a, a_mn=True, b_mn=True, out=pgrad(base + _F2W))           # dW2 = dY^T a
            if not fuse_cs:
                ops.colsum(dxb, out=pgrad(base + _F2B))
            dpre = ops.gemm(dxb, c.get(params[base + _F2W], ("fc2", i)), b_mn=True, act=self._act_bwd, aux_in=pre)
            del a, pre
            if ops.gemm_fuses_asum(dpre.shape[1], D):      # dW1 and db1 = column sums of dpre from one pass over dpre
                ops.gemm(dpre, h2, a_mn=True, b_mn=True, out=pgrad(base + _F1W), asum_out=pgrad(base + _F1B))
            else:
                ops.gemm(dpre, h2, a_mn=True, b_mn=True, out=pgrad(base + _F1W))         # dW1
                ops.colsum(dpre, out=pgrad(base + _F1B))
            dh2 = ops.gemm(dpre, c.get(params[base + _F1W], ("fc1", i)), b_mn=True)
            del dpre, h2
            dx1, dx1b = ops.layernorm_bwd(dh2, x1, mean2, rstd2, p[_N2W], pgrad(base + _N2W), pgrad(base + _N2B),
                                          dres=dx, want_bf16=True, row_scale=branch_scale(i, 0), rows_per_group=T,
                                          colsum_out=pgrad(base + _PB) if fuse_cs else None)
            del dh2, x1, dx
            # ---- attention branch: x1 = x + s * (o Wp^T + bp)
            ops.gemm(dx1b, o.view(M, D), a_mn=True, b_mn=True, out=pgrad(base + _PW))
            if not fuse_cs:
                ops.colsum(dx1b, out=pgrad(base + _PB))
            do = ops.gemm(dx1b, c.get(params[base + _PW], ("proj", i)), b_mn=True)
            del dx1b
            dqkv = torch.empty_like(qkv)
            q5, g5 = qkv.view(B, T, 3, H, d), dqkv.view(B, T, 3, H, d)
            ops.attention_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], o, lse, do.view(B, T, H, d), scale,
                              dq=g5[:, :, 0], dk=g5[:, :, 1], dv=g5[:, :, 2])
            del do, o, lse, qkv
            if ops.gemm_fuses_asum(3 * D, D):              # dWqkv and the q / v bias gradients from one pass over dqkv
                if self._full_qkv_bias:
                    ops.gemm(dqkv, h, a_mn=True, b_mn=True, out=pgrad(base + _QKVW), asum_out=pgrad(base + _QB))
                else:
                    bsum = torch.empty(3 * D, device=dev, dtype=F32)
                    ops.gemm(dqkv, h, a_mn=True, b_mn=True, out=pgrad(base + _QKVW), asum_out=bsum)
                    pgrad(base + _QB).copy_(bsum[:D])
                    pgrad(base + _VB).copy_(bsum[2 * D:])                          # k has no bias (eva_vit_model.py:307)
            else:
                ops.gemm(dqkv, h, a_mn=True, b_mn=True, out=pgrad(base + _QKVW))
                if self._full_qkv_bias:
                    ops.colsum(dqkv, out=pgrad(base + _QB))
                else:
                    ops.colsum2(dqkv, D, D, D, pgrad(base + _QB), pgrad(base + _VB))   # k has no bias (eva_vit_model.py:307)
            dh = ops.gemm(dqkv, c.get(params[base + _QKVW], ("qkv", i)), b_mn=True)
            del dqkv, h
            dx, dxb = ops.layernorm_bwd(dh, xr, mean1, rstd1, p[_N1W], pgrad(base + _N1W), pgrad(base + _N1B),
                                        dres=dx1, want_bf16=True, row_scale=branch_scale(i - 1, 1), rows_per_group=T,
                                        colsum_out=pgrad(base - _NBLK + _F2B) if (fuse_cs and i > 0) else None)
            del dh, dx1, xr
            if self.grad_bucket_hook is not None:      # block i's gradients are final and contiguous: reduce them now
                self.grad_bucket_hook(flat[offs[base]:offs[base + _NBLK]])
        # ---- patch embedding / cls / pos (eva_vit_model.py:613-619); pixels get no gradient
        cols = saved.pop("cols")
        if self._ln_pre:
            x0, m0, r0 = saved.pop("ln_pre")
            dx, dxb = ops.layernorm_bwd(dx, x0, m0, r0, params[_LPW].detach(), pgrad(_LPW), pgrad(_LPB), want_bf16=True)
        k = params[_PEW][0].numel()
        ops.gemm(dxb, cols[:, :k], a_mn=True, b_mn=True, out=pgrad(_PEW).view(D, k))   # cls rows of `cols` are zero
        if params[_PEB].numel():
            gb = ops.colsum(dxb, out=pgrad(_PEB))                      # all rows ...
            ops.colsum(dxb.view(B, T * D)[:, :D], out=gb, accumulate=2)   # ... minus the cls rows
        gpos = ops.batch_sum(dx, B, out=pgrad(_POS).view(-1))
        pgrad(_CLS).view(-1).copy_(gpos[:D])
        if self.grad_bucket_hook is not None:
            self.grad_bucket_hook(flat[:offs[_NTOP]])
        return grads

    # ------------------------------------------------------------------ public forward
    def forward_features(self, x, return_all_features=False):
        if x.dim() == 4 and x.stride(1) == 0 and x.shape[1] == 3:
            x = x[:, 0]                  # expanded single channel -> replicate inside the im2col kernel
        if x.dim() not in (3, 4):
            raise MicoError("EVAVisionTransformer expects (B, C, H, W) pixels (or (B, H, W) for one replicated channel)")
        B, Himg, Wimg = x.shape[0], x.shape[-2], x.shape[-1]
        assert Himg == self.patch_embed.img_size[0] and Wimg == self.patch_embed.img_size[1], \
            f"Input image size ({Himg}*{Wimg}) doesn't match model ({self.patch_embed.img_size[0]}*{self.patch_embed.img_size[1]})."
        if not x.is_cuda:
            raise MicoError("mico_b200 runs on CUDA (sm_100a) only; move the module and inputs to the GPU")
        x = x.contiguous()
        if x.dtype != F32:
            x = x.float()
        dp = self._draw_drop_path(B, x.device)
        flat = self._flat_params()
        keep = torch.is_grad_enabled() and any(p.requires_grad for p in flat if p is not None)   # keep activations?
        flat = [p if p is not None else x.new_empty(0) for p in flat]    # slots a variant does not have
        y = _TowerFn.apply(self, keep, x, dp, *flat)
        if not return_all_features:
            return y[:, 0]      # fc_norm is None when use_mean_pooling=False (eva_vit_model.py:643-648)
        return y

    def forward_multi(self, inputs):
        """Several pixel batches -- (B_i, C, H, W), or (B_i, H, W) for one replicated channel (spectrograms) -- through ONE
        pass of the tower (one weight read, one backward, one gradient buffer); returns the list of (B_i, T, D) token
        tensors.  Same result per sample as separate forward(x, return_all_features=True) calls (every op is
        per-sample; DropPath masks are drawn per sample either way)."""
        xs = []
        for x in inputs:
            if x.dim() == 4 and x.stride(1) == 0 and x.shape[1] == 3:
                x = x[:, 0]
            if x.dim() not in (3, 4) or not x.is_cuda:
                raise MicoError("forward_multi expects CUDA (B, C, H, W) / (B, H, W) tensors")
            assert x.shape[-2] == self.patch_embed.img_size[0] and x.shape[-1] == self.patch_embed.img_size[1]
            x = x.contiguous()
            xs.append(x if x.dtype == F32 else x.float())
        B = sum(x.shape[0] for x in xs)
        dp = self._draw_drop_path(B, xs[0].device)
        flat = self._flat_params()
        keep = torch.is_grad_enabled() and any(p.requires_grad for p in flat if p is not None)
        flat = [p if p is not None else xs[0].new_empty(0) for p in flat]
        y = _TowerFn.apply(self, keep, tuple(xs), dp, *flat)
        return list(torch.split(y, [x.shape[0] for x in xs], dim=0))

    def forward(self, x, return_all_features=False):
        if return_all_features:
            return self.forward_features(x, return_all_features)
        x = self.forward_features(x)
        return self.head(x)
