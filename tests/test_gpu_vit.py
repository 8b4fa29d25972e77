"""GPU parity of the EVA ViT tower (mico_b200.eva_vit) through the C-ABI kernels.

  * against the golden fixture produced by the unmodified reference (tests/golden/eva_vit_tiny.pt):
    width 176 = 2 heads x 88, depth 2, 257 tokens -- eval forward, training forward with the reference's
    DropPath masks, and every parameter gradient;
  * against the CPU oracle at ViT-g width (1408, 16 heads x 88, MLP 6144), 2 blocks, seeded weights.

Tolerances (bf16 operands / fp32 accumulation vs an fp32 reference): each GEMM operand is rounded to bf16
(2^-9 relative), so features agree to ~3e-3 rel-L2 and gradients to ~1e-2 (a chain of ~2x as many rounded
operands plus bf16 dY); the scalar loss to 1e-3 relative as BASELINE.json states."""
import os

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

FEAT_TOL = 5e-3
GRAD_TOL = 2e-2
LOSS_TOL = 1e-3


def _tower_from_cfg(cfg, drop_path_rate=0.0):
    from mico_b200.eva_vit import EVAVisionTransformer
    return EVAVisionTransformer(img_size=cfg["image"], patch_size=cfg["patch"], num_classes=8, use_mean_pooling=False,
                                embed_dim=cfg["width"], depth=cfg["depth"], num_heads=cfg["heads"],
                                mlp_ratio=cfg["mlp"] / cfg["width"], qkv_bias=True, drop_path_rate=drop_path_rate,
                                eps=cfg["eps"])


def test_golden_eval_forward(golden_dir):
    g = torch.load(os.path.join(golden_dir, "eva_vit_tiny.pt"), weights_only=False)
    m = _tower_from_cfg(g["cfg"], 0.4)
    missing, unexpected = m.load_state_dict(g["state_dict"], strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        y = m(g["x"].cuda(), return_all_features=True)
    err = rel_l2(y.cpu(), g["y_eval"])
    print(f"golden eval fwd rel-L2 {err:.3e}")
    assert y.shape == (2, 257, 176) and y.dtype == torch.float32
    assert err < FEAT_TOL


def test_golden_train_forward_backward(golden_dir):
    g = torch.load(os.path.join(golden_dir, "eva_vit_tiny.pt"), weights_only=False)
    m = _tower_from_cfg(g["cfg"], 0.4)
    m.load_state_dict(g["state_dict"], strict=True)
    m = m.cuda().train()
    m.inject_drop_path_scales(g["dp_scales"])
    y = m(g["x"].cuda(), return_all_features=True)
    assert rel_l2(y.detach().cpu(), g["y_train"]) < FEAT_TOL
    loss = y.float().pow(2).mean()
    assert abs(loss.item() - g["loss"].item()) <= LOSS_TOL * abs(g["loss"].item())
    loss.backward()
    worst = ("", 0.0)
    for k, p in m.named_parameters():
        if k.startswith("head."):
            assert p.grad is None
            continue
        e = rel_l2(p.grad.cpu(), g["grads"][k])
        if e > worst[1]:
            worst = (k, e)
        assert e < GRAD_TOL, (k, e)
    print(f"golden train: loss {loss.item():.6f} vs {g['loss'].item():.6f}; worst grad {worst[0]} rel-L2 {worst[1]:.3e}")


@pytest.mark.parametrize("B", [1, 3])
def test_vitg_width_two_blocks_vs_oracle(B):
    from oracle import eva_vit as O
    cfg = O.vit_cfg(depth=2)
    p = O.init_params(cfg, seed=3)
    m = _tower_from_cfg(cfg, 0.0)
    sd = dict(p)
    sd["head.weight"], sd["head.bias"] = m.head.weight.detach(), m.head.bias.detach()
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, 224, 224, generator=gen)
    dp = torch.tensor([[[1.0] * B, [1.0] * B], [[0.0] + [1.25] * (B - 1), [1.25] * B]])
    m.inject_drop_path_scales(dp)
    y = m(x.cuda(), return_all_features=True)
    loss = y.pow(2).mean()
    loss.backward()
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    yr = O.forward_features(pr, x, cfg, dp_scales=dp)
    lr = yr.pow(2).mean()
    lr.backward()
    e = rel_l2(y.detach().cpu(), yr)
    print(f"ViT-g width x2 blocks B={B}: features rel-L2 {e:.3e}, loss {loss.item():.6f} vs {lr.item():.6f}")
    assert e < FEAT_TOL
    assert abs(loss.item() - lr.item()) <= LOSS_TOL * abs(lr.item())
    for k, v in m.named_parameters():
        if k.startswith("head."):
            continue
        ge = rel_l2(v.grad.cpu(), pr[k].grad)
        assert ge < GRAD_TOL, (k, ge)


def test_no_grad_forward_keeps_nothing_and_matches_train_forward():
    from oracle import eva_vit as O
    cfg = O.vit_cfg(width=176, depth=3, heads=2, mlp=352)
    m = _tower_from_cfg(cfg, 0.0).cuda()
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        y0 = m(x, return_all_features=True)
    y1 = m(x, return_all_features=True)
    assert y0.grad_fn is None and y1.grad_fn is not None
    assert torch.equal(y0, y1.detach())
    cls = m.forward_features(x)           # return_all_features=False -> cls token (eva:643-648)
    assert torch.equal(cls.detach(), y1.detach()[:, 0])


def test_activation_checkpointing_gives_identical_gradients():
    """config.checkpointing (eva_vit_model.py:635-637): recomputing each block in backward must not change a bit."""
    from oracle import eva_vit as O
    cfg = O.vit_cfg(width=176, depth=3, heads=2, mlp=352)
    torch.manual_seed(3)
    m = _tower_from_cfg(cfg, 0.3).cuda().train()
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(1)).cuda()
    dp = torch.tensor([[[1.0, 1.0]] * 2, [[1 / 0.85, 0.0]] * 2, [[0.0, 1 / 0.7], [1 / 0.7, 1 / 0.7]]])
    grads = []
    # (checkpointing, light, qkv, attn blocks): no checkpointing, input-only, and every mix of the three richer levels
    for ck, n3, n2, n1 in ((False, 0, 0, 0), (True, 0, 0, 0), (True, 1, 1, 1), (True, 0, 0, -1), (True, 0, 2, 0), (True, 2, 0, 0)):
        m.set_grad_checkpointing(ck)
        m.ckpt_light_blocks, m.ckpt_qkv_blocks, m.ckpt_attn_blocks = n3, n2, n1
        m.zero_grad(set_to_none=True)
        m.inject_drop_path_scales(dp)
        m(x, return_all_features=True).pow(2).mean().backward()
        grads.append({k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 30
    for g in grads[1:]:
        assert g.keys() == grads[0].keys()
        for k in grads[0]:
            assert torch.equal(grads[0][k], g[k]), k
