"""Oracle: EVA-CLIP vision tower (ViT-g/14 family), fp32 CPU, functional over a state_dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates, op for op,
  * PatchEmbed + cls/pos           model/evaclip/eva_vit_model.py:440-447, 613-619
  * Attention (non-xattn branch)   model/evaclip/eva_vit_model.py:293-365
  * Mlp                            model/evaclip/eva_vit_model.py:190-199
  * Block (pre-norm, no gamma)     model/evaclip/eva_vit_model.py:409-424
  * DropPath                       model/evaclip/eva_vit_model.py:121-138
  * forward_features               model/evaclip/eva_vit_model.py:611-650
  * LayerNorm eps=1e-6             model/evaclip/model.py:124, transformer.py:121-127
Only the EVA01-g configuration is restated (rope/rel-pos/subln/swiglu/postnorm are off
for g-14: model_configs/EVA01-CLIP-g-14.json).

`p` is a mapping name -> tensor using the reference's own key names below `prefix`
(e.g. prefix='vision_encoder.visual.').
"""
import math

import torch
import torch.nn.functional as F


def vit_cfg(width=1408, depth=40, heads=16, mlp=6144, patch=14, image=224, eps=1e-6):
    return dict(width=width, depth=depth, heads=heads, mlp=mlp, patch=patch, image=image, eps=eps)


VIT_G14 = vit_cfg()


def patch_embed(p, prefix, x, cfg):
    """Conv2d(k=s=patch) == per-patch linear map; flatten(2).transpose(1,2) (eva:440-447)."""
    w = p[prefix + "patch_embed.proj.weight"]
    b = p[prefix + "patch_embed.proj.bias"]
    y = F.conv2d(x, w, b, stride=cfg["patch"])
    return y.flatten(2).transpose(1, 2)


def attention(p, prefix, x, heads):
    """eva:305-311, 340-341, 358-363 (q scaled before q@k^T; k has no bias)."""
    B, N, C = x.shape
    w = p[prefix + "qkv.weight"]
    qb, vb = p[prefix + "q_bias"], p[prefix + "v_bias"]
    bias = torch.cat((qb, torch.zeros_like(vb), vb))
    qkv = F.linear(x, w, bias).reshape(B, N, 3, heads, -1).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    d = q.shape[-1]
    q = q * (d ** -0.5)
    att = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, N, -1)
    return F.linear(o, p[prefix + "proj.weight"], p[prefix + "proj.bias"])


def mlp(p, prefix, x):
    """fc2(GELU_erf(fc1(x))) (eva:190-199; ffn_ln is Identity when subln=False)."""
    h = F.linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"])
    h = F.gelu(h)  # nn.GELU default = exact erf
    return F.linear(h, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


def drop_path_scale(keep_mask, keep_prob):
    """DropPath multiplier: mask/keep_prob per sample (eva:121-138, scale_by_keep=True)."""
    if keep_prob > 0.0:
        return keep_mask / keep_prob
    return keep_mask


def block(p, prefix, x, cfg, dp_scale=None):
    """x + dp(attn(norm1(x))); x + dp(mlp(norm2(x)))  (eva:420-421).
    dp_scale: None (eval) or (2, B) per-sample multipliers for the two residual branches."""
    D = x.shape[-1]
    h = F.layer_norm(x, (D,), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], cfg["eps"])
    a = attention(p, prefix + "attn.", h, cfg["heads"])
    if dp_scale is not None:
        a = a * dp_scale[0].view(-1, 1, 1)
    x = x + a
    h = F.layer_norm(x, (D,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], cfg["eps"])
    m = mlp(p, prefix + "mlp.", h)
    if dp_scale is not None:
        m = m * dp_scale[1].view(-1, 1, 1)
    return x + m


def forward_features(p, x, cfg=VIT_G14, prefix="", dp_scales=None, return_all_features=True):
    """eva:611-650 with return_all_features=True (the only mode MiCo uses, mico.py:120).
    dp_scales: None or tensor (depth, 2, B) of DropPath multipliers (training parity)."""
    x = patch_embed(p, prefix, x, cfg)
    B = x.shape[0]
    cls = p[prefix + "cls_token"].expand(B, -1, -1)
    x = torch.cat((cls, x), dim=1) + p[prefix + "pos_embed"]
    for i in range(cfg["depth"]):
        x = block(p, f"{prefix}blocks.{i}.", x, cfg, None if dp_scales is None else dp_scales[i])
    D = x.shape[-1]
    x = F.layer_norm(x, (D,), p[prefix + "norm.weight"], p[prefix + "norm.bias"], cfg["eps"])
    if return_all_features:
        return x
    return x[:, 0]  # fc_norm is None for g-14 (use_mean_pooling False)


def drop_path_rates(depth, rate=0.4):
    """torch.linspace(0, rate, depth) (eva:533)."""
    return [v.item() for v in torch.linspace(0, rate, depth)]


def init_params(cfg, seed=0, prefix="", scale_like_reference=True):
    """Seeded random parameters with the reference's shapes and init statistics
    (trunc_normal .02, proj/fc2 rescaled by 1/sqrt(2*(layer+1)): eva:563-593).  The values are NOT
    bit-identical to the reference's RNG stream; parity tests pass the same dict to both sides."""
    g = torch.Generator().manual_seed(seed)
    W, F_, P = cfg["width"], cfg["mlp"], cfg["patch"]
    n_tok = (cfg["image"] // P) ** 2 + 1

    def tn(*shape, std=0.02):
        return torch.nn.init.trunc_normal_(torch.empty(*shape), std=std, a=-2 * std, b=2 * std, generator=g)

    def small(*shape):  # non-zero biases/affines so that parity exercises every term
        return 0.02 * torch.randn(*shape, generator=g)

    p = {
        prefix + "cls_token": tn(1, 1, W),
        prefix + "pos_embed": tn(1, n_tok, W),
        prefix + "patch_embed.proj.weight": tn(W, 3, P, P),
        prefix + "patch_embed.proj.bias": small(W),
        prefix + "norm.weight": 1.0 + small(W),
        prefix + "norm.bias": small(W),
    }
    for i in range(cfg["depth"]):
        b = f"{prefix}blocks.{i}."
        r = 1.0 / math.sqrt(2.0 * (i + 1)) if scale_like_reference else 1.0
        p.update({
            b + "norm1.weight": 1.0 + small(W), b + "norm1.bias": small(W),
            b + "norm2.weight": 1.0 + small(W), b + "norm2.bias": small(W),
            b + "attn.qkv.weight": tn(3 * W, W),
            b + "attn.q_bias": small(W), b + "attn.v_bias": small(W),
            b + "attn.proj.weight": tn(W, W) * r, b + "attn.proj.bias": small(W),
            b + "mlp.fc1.weight": tn(F_, W), b + "mlp.fc1.bias": small(F_),
            b + "mlp.fc2.weight": tn(W, F_) * r, b + "mlp.fc2.bias": small(W),
        })
    return p
