"""GPU parity of the Swin encoder (mico_b200.swin) against the golden fixture produced by the unmodified reference
model/swin.py (embed_dim 32, depths [2,2], heads [1,2] = head_dim 32 as in Swin-B, window 7, shifted windows, patch merge):
forward_features and every parameter gradient, including the relative-position-bias table (mico_attention_dmask)."""
import os

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def test_swin_matches_reference(golden_dir):
    from mico_b200.swin import SwinTransformer
    g = torch.load(os.path.join(golden_dir, "swin_tiny.pt"), weights_only=False)
    m = SwinTransformer(img_size=56, patch_size=4, in_chans=3, num_classes=0, embed_dim=32, depths=[2, 2], num_heads=[1, 2],
                        window_size=7, mlp_ratio=4., drop_path_rate=0.0)
    missing, unexpected = m.load_state_dict(g["state_dict"], strict=True)
    m = m.cuda().train()
    y = m.forward_features(g["x"].cuda())
    e = rel_l2(y.detach().cpu(), g["y"])
    y.float().pow(2).mean().backward()
    errs = sorted(((rel_l2(p.grad.cpu(), g["grads"][k]), k) for k, p in m.named_parameters() if k in g["grads"]),
                  reverse=True)
    print(f"swin: y {e:.3e}; worst grads {[(f'{a:.2e}', k) for a, k in errs[:3]]}")
    assert y.shape == (3, 49, 64)
    assert e < 5e-3
    assert errs[0][0] < 3e-2
    tbl = [a for a, k in errs if "relative_position_bias_table" in k]
    assert len(tbl) == 4 and max(tbl) < 3e-2
