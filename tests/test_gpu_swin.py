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


def test_swin_b_size_matches_reference(golden_dir):
    """Swin-B dimensions (BASELINE configs[3]: embed 128, depths [2,2,18,2], heads [4,8,16,32], 224 x 224; VERDICT r1: "Swin-B
    shapes were never run against anything"): forward_features, selected gradients and the norm of every parameter gradient
    against tests/golden/swin_b.pt, produced by the unmodified reference model/swin.py on parameters that both sides
    regenerate from per-tensor seeds (oracle.fullsize.seeded_params_)."""
    from mico_b200.swin import SwinTransformer
    from oracle import fullsize as FS
    g = torch.load(os.path.join(golden_dir, "swin_b.pt"), weights_only=False)
    m = FS.seeded_params_(SwinTransformer(**FS.SWIN_B), seed=5)
    assert sum(p.numel() for p in m.parameters()) == g["n_params"]
    m = m.cuda().train()
    y = m.forward_features(FS.swin_b_input().cuda())
    e = rel_l2(y.detach().cpu(), g["y"])
    y.float().pow(2).mean().backward()
    named = dict(m.named_parameters())
    errs = sorted(((rel_l2(named[k].grad.cpu(), v), k) for k, v in g["grads"].items()), reverse=True)
    nerr = sorted(((abs(float(named[k].grad.norm()) - n) / max(n, 1e-12), k) for k, n in g["grad_norms"].items() if n > 1e-9),
                  reverse=True)
    print(f"swin-B: y rel-L2 {e:.3e}; worst selected grads {[(f'{a:.2e}', k) for a, k in errs[:3]]}; "
          f"worst grad-norm error over {len(nerr)} tensors {nerr[0][0]:.2e} ({nerr[0][1]}), median {nerr[len(nerr) // 2][0]:.2e}")
    assert y.shape == (2, 49, 1024)
    assert e < 1e-2                      # 24 blocks in bf16 (the 2-block fixture above: 5e-3)
    assert errs[0][0] < 5e-2
    assert nerr[0][0] < 5e-2 and nerr[len(nerr) // 2][0] < 1e-2
