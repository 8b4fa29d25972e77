"""GPU parity of the OpenAI-CLIP vision tower (mico_b200.clip_vit) against the golden fixture produced by the unmodified
reference model/clip/clip.py VisionTransformer (width 128, 2 layers, patch 16, 197 tokens; QuickGELU, ln_pre, in_proj_bias)."""
import os

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def test_clip_vit_matches_reference(golden_dir):
    from mico_b200.clip_vit import VisionTransformer
    g = torch.load(os.path.join(golden_dir, "clip_vit_tiny.pt"), weights_only=False)
    m = VisionTransformer(input_resolution=224, patch_size=16, width=128, layers=2, heads=2, output_dim=32)
    m.load_state_dict(g["state_dict"], strict=True)
    m = m.cuda().train()
    y = m(g["x"].cuda(), return_all_features=True)
    e = rel_l2(y.detach().cpu(), g["y"])
    y.float().pow(2).mean().backward()
    worst = max((rel_l2(p.grad.cpu(), g["grads"][k]), k) for k, p in m.named_parameters() if k in g["grads"])
    with torch.no_grad():
        pooled = m(g["x"].cuda())
    ep = rel_l2(pooled.cpu(), g["pooled"])
    print(f"clip vit: y {e:.3e} pooled {ep:.3e} worst grad {worst[1]} {worst[0]:.3e}")
    assert y.shape == (2, 197, 128)
    assert e < 5e-3 and ep < 5e-3 and worst[0] < 2e-2
    assert m.proj.grad is None       # proj is not on the return_all_features path
