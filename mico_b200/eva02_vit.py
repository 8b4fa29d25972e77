"""EVA02 vision towers (``evaclip02_base`` / ``evaclip02_large``) on the sm_100a kernels: the reference's
``EVAVisionTransformer`` (model/evaclip/eva_vit_model.py:488-650) in its EVA02 configuration -- ``rope=True``,
``naiveswiglu=True``, ``subln=True`` -- with the same constructor keywords and ``state_dict`` keys.

Differences from the EVA01-g tower of ``eva_vit.py``, all in model/evaclip/eva_vit_model.py:
  * Attention (:226-365) with sub-LN: separate ``q_proj`` / ``k_proj`` / ``v_proj`` without bias, ``q_bias`` / ``v_bias`` added to
    q and v (:295-302), rotary position embedding on every token but cls (:314-322; tables of
    model/evaclip/rope.py:79-136 ``VisionRotaryEmbeddingFast``), ``inner_attn_ln`` over the concatenated heads before
    ``proj`` (:360);
  * SwiGLU MLP (:201-224): ``w3(ffn_ln(silu(w1 x) * w2 x))``, hidden = int(dim * mlp_ratio) (2730 for EVA02-L: padded to a
    multiple of 8 for the 16-byte rows the GEMM's TMA loads need -- the padding columns are exact zeros end to end);
  * pre-norm residual blocks as EVA01 (:417-424), DropPath per sample.

Compute: patch embedding = im2col kernel + tcgen05 GEMM; every Linear = tcgen05 GEMM (bf16 operands, fp32 accumulation);
LayerNorm / RoPE / SwiGLU = their CUDA kernels; attention = the fused tcgen05 kernels (head_dim 64).  Each op is one autograd
node (mico_b200/functional.py) -- the EVA01-g tower's single-node launch sequence with checkpoint levels and the flat
gradient buffer is not replicated for this variant; ``grad_checkpointing`` recomputes block by block
(eva_vit_model.py:635-637) through torch.utils.checkpoint.
"""
import math

import torch
import torch.nn as nn
from torch.utils.checkpoint import checkpoint

from . import functional as MF
from .ops import BF16, F32, MicoError


class _Linear(nn.Module):
    def __init__(self, i, o, bias=True):
        super().__init__()
        self.weight = nn.Parameter(nn.init.trunc_normal_(torch.empty(o, i), std=.02, a=-2.0, b=2.0))
        self.bias = nn.Parameter(torch.zeros(o)) if bias else None


class LayerNorm(nn.Module):
    def __init__(self, d, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))

    def forward(self, x, out_dtype=F32):
        return MF.layer_norm(x, self.weight, self.bias, self.eps, out_dtype)


class _Conv(nn.Module):
    def __init__(self, c, d, p):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(d, c, p, p))
        self.bias = nn.Parameter(torch.zeros(d))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))


class PatchEmbed(nn.Module):
    """eva_vit_model.py:427-447"""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.patch_shape = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.patch_shape[0] * self.patch_shape[1]
        self.proj = _Conv(in_chans, embed_dim, patch_size)

    def forward(self, x):
        return MF.patch_embed(x, self.proj.weight, self.proj.bias, self.patch_size[0])


class VisionRotaryEmbeddingFast(nn.Module):
    """model/evaclip/rope.py:79-136: cos / sin tables [G*G, head_dim], state_dict buffers as in the reference (the tower and
    every block's attention hold the SAME module: ``rope.*`` and ``blocks.N.attn.rope.*`` keys)."""

    def __init__(self, dim, pt_seq_len=16, ft_seq_len=None, theta=10000.0):
        super().__init__()
        ft_seq_len = pt_seq_len if ft_seq_len is None else ft_seq_len
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        t = torch.arange(ft_seq_len) / ft_seq_len * pt_seq_len
        f = torch.einsum("i,f->if", t, freqs).repeat_interleave(2, dim=-1)       # [G, dim], pairs interleaved (n r), r = 2
        full = torch.cat((f[:, None, :].expand(ft_seq_len, ft_seq_len, dim), f[None, :, :].expand(ft_seq_len, ft_seq_len, dim)),
                         dim=-1).reshape(ft_seq_len * ft_seq_len, 2 * dim)
        self.register_buffer("freqs_cos", full.cos().contiguous())
        self.register_buffer("freqs_sin", full.sin().contiguous())


class Attention(nn.Module):
    """eva_vit_model.py:226-365 (subln, rope, no relative position bias)"""

    def __init__(self, dim, num_heads, eps, rope):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.rope = rope
        self.q_proj = _Linear(dim, dim, bias=False)
        self.k_proj = _Linear(dim, dim, bias=False)
        self.v_proj = _Linear(dim, dim, bias=False)
        self.q_bias = nn.Parameter(torch.zeros(dim))
        self.v_bias = nn.Parameter(torch.zeros(dim))
        self.inner_attn_ln = LayerNorm(dim, eps)
        self.proj = _Linear(dim, dim)

    def forward(self, h, residual=None):
        """h: bf16 [B, N, C] (norm1 output) -> fp32 [B, N, C]: proj output (+ residual, added in the GEMM epilogue)"""
        B, N, C = h.shape
        H = self.num_heads
        d = C // H
        q = MF.linear_tc(h, self.q_proj.weight, self.q_bias).view(B, N, H, d)
        k = MF.linear_tc(h, self.k_proj.weight, None).view(B, N, H, d)
        v = MF.linear_tc(h, self.v_proj.weight, self.v_bias, out_dtype=BF16).view(B, N, H, d)
        q = MF.rope(q, self.rope.freqs_cos, self.rope.freqs_sin)    # fp32 -> rotated bf16 operand (cls token passes through)
        k = MF.rope(k, self.rope.freqs_cos, self.rope.freqs_sin)
        o = MF.attention(q, k, v, self.scale)                   # bf16 [B, N, H, d]
        o = self.inner_attn_ln(o.view(B, N, C).float(), out_dtype=BF16)
        return MF.linear_tc(o, self.proj.weight, self.proj.bias, residual=residual)


class SwiGLU(nn.Module):
    """eva_vit_model.py:201-224"""

    def __init__(self, dim, hidden, eps):
        super().__init__()
        self.w1 = _Linear(dim, hidden)
        self.w2 = _Linear(dim, hidden)
        self.ffn_ln = LayerNorm(hidden, eps)
        self.w3 = _Linear(hidden, dim)

    def forward(self, h, residual=None):
        """h: bf16 [B, N, C] (norm2 output) -> fp32 [B, N, C] (+ residual, added in the GEMM epilogue)"""
        hid = self.w1.weight.shape[0]
        pad = (-hid) % 8
        w1, b1, w2, b2, w3 = self.w1.weight, self.w1.bias, self.w2.weight, self.w2.bias, self.w3.weight
        if pad:      # zero rows / columns: u1 = u2 = 0 there, silu(0) * 0 = 0, and w3 ignores them
            w1, w2 = nn.functional.pad(w1, (0, 0, 0, pad)), nn.functional.pad(w2, (0, 0, 0, pad))
            b1, b2 = nn.functional.pad(b1, (0, pad)), nn.functional.pad(b2, (0, pad))
            w3 = nn.functional.pad(w3, (0, pad))
        g = MF.swiglu(MF.linear_tc(h, w1, b1), MF.linear_tc(h, w2, b2))
        if pad:
            g = nn.functional.pad(self.ffn_ln(g[..., :hid].contiguous()), (0, pad))
            g = MF.cast_bf16(g)
        else:
            g = self.ffn_ln(g, out_dtype=BF16)
        return MF.linear_tc(g, w3, self.w3.bias, residual=residual)


class Block(nn.Module):
    """eva_vit_model.py:368-424 with gamma_1 = None, postnorm = False"""

    def __init__(self, dim, num_heads, mlp_ratio, drop_path, eps, rope):
        super().__init__()
        self.norm1 = LayerNorm(dim, eps)
        self.attn = Attention(dim, num_heads, eps, rope)
        self.drop_prob = float(drop_path)
        self.norm2 = LayerNorm(dim, eps)
        self.mlp = SwiGLU(dim, int(dim * mlp_ratio), eps)

    def _drop_path(self, y):
        """timm DropPath (eva_vit_model.py:121-138): per-sample keep mask / keep_prob; identity in eval mode"""
        if self.drop_prob == 0.0 or not self.training:
            return y
        keep = 1.0 - self.drop_prob
        m = torch.empty((y.shape[0], 1, 1), device=y.device, dtype=y.dtype).bernoulli_(keep)
        if keep > 0.0:
            m.div_(keep)
        return y * m

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:      # every EVA02-CLIP configuration: residual adds in the GEMM epilogues
            x = self.attn(self.norm1(x, out_dtype=BF16), residual=x)
            return self.mlp(self.norm2(x, out_dtype=BF16), residual=x)
        x = x + self._drop_path(self.attn(self.norm1(x, out_dtype=BF16)))
        return x + self._drop_path(self.mlp(self.norm2(x, out_dtype=BF16)))


class EVA02VisionTransformer(nn.Module):
    """Drop-in for model/evaclip/eva_vit_model.py:488 with rope / naiveswiglu / subln (EVA02-CLIP-B-16, -L-14)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=True, drop_path_rate=0., norm_layer=None, use_mean_pooling=False, init_scale=0.001,
                 grad_checkpointing=False, xattn=True, rope=True, pt_hw_seq_len=16, intp_freq=True, naiveswiglu=True,
                 subln=True, eps=1e-6, **unsupported):
        super().__init__()
        bad = {k: v for k, v in unsupported.items() if v}
        if bad or not (rope and naiveswiglu and subln and qkv_bias) or use_mean_pooling:
            raise NotImplementedError(f"EVA02 tower options outside the EVA02-CLIP-B/L configuration: {bad}")
        if embed_dim % num_heads or (embed_dim // num_heads) % 8:
            raise NotImplementedError("head_dim must be a multiple of 8 (16-byte rows for TMA)")
        if norm_layer is not None:
            eps = getattr(norm_layer, "keywords", {}).get("eps", eps)
        self.image_size = img_size
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        hw = img_size // patch_size
        self.rope = VisionRotaryEmbeddingFast(embed_dim // num_heads // 2, pt_hw_seq_len, hw if intp_freq else None)
        dpr = [v.item() for v in torch.linspace(0, drop_path_rate, depth, device="cpu")]
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads, mlp_ratio, dpr[i], eps, self.rope) for i in range(depth)])
        self.norm = LayerNorm(embed_dim, eps)
        self.fc_norm = None
        self.head = _Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
        with torch.no_grad():   # fix_init_weight (eva_vit_model.py:563-573, the naiveswiglu branch rescales w3)
            for i, blk in enumerate(self.blocks):
                blk.attn.proj.weight.div_(math.sqrt(2.0 * (i + 1)))
                blk.mlp.w3.weight.div_(math.sqrt(2.0 * (i + 1)))
            if isinstance(self.head, _Linear):
                self.head.weight.mul_(init_scale)
                self.head.bias.mul_(init_scale)
        self.grad_checkpointing = grad_checkpointing

    def set_grad_checkpointing(self, enable=True):
        self.grad_checkpointing = enable

    def no_weight_decay(self):
        return {"pos_embed", "cls_token"}

    def forward_features(self, x, return_all_features=False):
        if x.dim() == 3:                 # one channel replicated three times (forward_audio_encoder, mico.py:139-143)
            x = x[:, None].expand(-1, 3, -1, -1)
        if x.dim() != 4 or not x.is_cuda:
            raise MicoError("EVA02VisionTransformer expects CUDA (B, C, H, W) pixels (or (B, H, W) for one replicated channel)")
        assert x.shape[-2] == self.patch_embed.img_size[0] and x.shape[-1] == self.patch_embed.img_size[1], \
            f"Input image size ({x.shape[-2]}*{x.shape[-1]}) doesn't match model ({self.patch_embed.img_size[0]}*{self.patch_embed.img_size[1]})."
        t = self.patch_embed(x)
        t = torch.cat((self.cls_token.expand(t.shape[0], -1, -1), t), dim=1) + self.pos_embed
        for blk in self.blocks:
            if self.grad_checkpointing and torch.is_grad_enabled():
                t = checkpoint(blk, t, use_reentrant=False, preserve_rng_state=True)
            else:
                t = blk(t)
        t = self.norm(t)
        return t if return_all_features else t[:, 0]

    def forward(self, x, return_all_features=False):
        if return_all_features:
            return self.forward_features(x, return_all_features)
        x = self.forward_features(x)
        return MF.linear(x, self.head.weight, self.head.bias) if isinstance(self.head, _Linear) else x
