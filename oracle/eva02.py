"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the EVA02 tower variant (not built in CUDA yet: this oracle and
its golden fixture are the parity target for that work).

Differences from the EVA01-g tower of oracle/eva_vit.py, all in model/evaclip/eva_vit_model.py:
  * Attention (:226-365) with subln: separate q_proj / k_proj / v_proj (no bias; q_bias / v_bias added to q and v), rotary
    position embedding on every token except cls (:314-322), LayerNorm over the concatenated heads (`inner_attn_ln`, :360)
    before `proj`;
  * RoPE tables (model/evaclip/rope.py:79-136, VisionRotaryEmbeddingFast): dim = head_dim / 2, freqs = theta^(-2i/dim),
    positions t = arange(G) / G * pt_seq_len, pairs interleaved (n r) with r = 2, the h table and the w table concatenated ->
    [G*G, head_dim]; rotate_half maps (x1, x2) -> (-x2, x1) on interleaved pairs (rope.py:20-24);
  * SwiGLU MLP (:201-224): w3(ffn_ln(silu(w1 x) * w2 x)), hidden = int(dim * mlp_ratio) (2730 for EVA02-L: not a multiple of 8);
  * Block (:409-424) pre-norm without layer scale as for EVA01.
Pinned by tests/golden/eva02_tiny.pt (reference EVAVisionTransformer with rope / naiveswiglu / subln, xattn=False path).
"""
import math

import torch
import torch.nn.functional as F


def rope_tables(head_dim, grid, pt_seq_len=16, theta=10000.0):
    dim = head_dim // 2
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
    t = torch.arange(grid) / grid * pt_seq_len                         # intp_freq: ft_seq_len = grid
    f = torch.einsum("i,f->if", t, freqs).repeat_interleave(2, dim=-1)  # [G, dim], pairs interleaved
    fh = f[:, None, :].expand(grid, grid, dim)
    fw = f[None, :, :].expand(grid, grid, dim)
    full = torch.cat((fh, fw), dim=-1).reshape(grid * grid, 2 * dim)    # [G*G, head_dim]
    return full.cos(), full.sin()


def rotate_half(x):
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(dim=-1)
    return torch.stack((-x2, x1), dim=-1).reshape(*x.shape[:-2], -1)


def apply_rope(t, cos, sin):
    """t: [B, H, N, d] with the cls token first; rotates tokens 1.. (eva_vit_model.py:314-322)."""
    rot = t[:, :, 1:, :] * cos + rotate_half(t[:, :, 1:, :]) * sin
    return torch.cat((t[:, :, :1, :], rot), dim=-2)


def block(p, pre, x, heads, eps, cos, sin):
    B, N, C = x.shape
    d = C // heads
    h = F.layer_norm(x, (C,), p[pre + "norm1.weight"], p[pre + "norm1.bias"], eps)
    q = F.linear(h, p[pre + "attn.q_proj.weight"], p[pre + "attn.q_bias"])
    k = F.linear(h, p[pre + "attn.k_proj.weight"])
    v = F.linear(h, p[pre + "attn.v_proj.weight"], p[pre + "attn.v_bias"])
    q, k, v = (t.reshape(B, N, heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    a = torch.softmax((q * d ** -0.5) @ k.transpose(-2, -1), dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, N, C)
    o = F.layer_norm(o, (C,), p[pre + "attn.inner_attn_ln.weight"], p[pre + "attn.inner_attn_ln.bias"], eps)
    x = x + F.linear(o, p[pre + "attn.proj.weight"], p[pre + "attn.proj.bias"])
    h = F.layer_norm(x, (C,), p[pre + "norm2.weight"], p[pre + "norm2.bias"], eps)
    g = F.silu(F.linear(h, p[pre + "mlp.w1.weight"], p[pre + "mlp.w1.bias"])) * F.linear(h, p[pre + "mlp.w2.weight"], p[pre + "mlp.w2.bias"])
    g = F.layer_norm(g, (g.shape[-1],), p[pre + "mlp.ffn_ln.weight"], p[pre + "mlp.ffn_ln.bias"], eps)
    return x + F.linear(g, p[pre + "mlp.w3.weight"], p[pre + "mlp.w3.bias"])


def forward_features(p, x, cfg):
    """p: reference state_dict of EVAVisionTransformer(rope, naiveswiglu, subln); x: [B, 3, H, W] -> [B, 1 + G*G, width]."""
    P, C, heads, eps = cfg["patch"], cfg["width"], cfg["heads"], cfg["eps"]
    G = cfg["image"] // P
    t = F.conv2d(x, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"], stride=P).flatten(2).transpose(1, 2)
    t = torch.cat((p["cls_token"].expand(x.shape[0], -1, -1), t), dim=1) + p["pos_embed"]
    cos, sin = rope_tables(C // heads, G, cfg.get("pt_hw_seq_len", 16))
    for i in range(cfg["depth"]):
        t = block(p, f"blocks.{i}.", t, heads, eps, cos, sin)
    return F.layer_norm(t, (C,), p["norm.weight"], p["norm.bias"], eps)
