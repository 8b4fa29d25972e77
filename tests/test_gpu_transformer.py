"""GPU parity of the generic encoder (mico_b200.transformer, mirror of model/transformer.py) against the golden fixture
produced by the unmodified reference (tests/golden/transformer_tiny.pt), pre-norm and post-norm."""
import os

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", ["prenorm", "postnorm"])
def test_transformer_encoder_matches_reference(golden_dir, mode):
    from mico_b200.mico import _AttrDict
    from mico_b200.transformer import TransformerEncoder
    g = torch.load(os.path.join(golden_dir, "transformer_tiny.pt"), weights_only=False)[mode]
    cfg = _AttrDict(hidden_size=128, num_attention_heads=2, intermediate_size=256, num_hidden_layers=2, hidden_dropout=0.0,
                    attention_dropout=0.0, checkpointing=False)
    m = TransformerEncoder(cfg, mode)
    m.load_state_dict(g["state_dict"], strict=True)
    m = m.cuda().train()
    x = g["x"].cuda().requires_grad_(True)
    y, _ = m(x, g["mask"].cuda())
    e = rel_l2(y.detach().cpu(), g["y"])
    y.float().pow(2).sum().backward()
    ex = rel_l2(x.grad.cpu(), g["dx"])
    worst = max((rel_l2(p.grad.cpu(), g["grads"][k]), k) for k, p in m.named_parameters()
                if g["grads"][k].norm() > 1e-6 and not k.endswith("linears.1.bias"))
    print(f"{mode}: y {e:.3e} dx {ex:.3e} worst param grad {worst[1]} {worst[0]:.3e}")
    assert e < 5e-3 and ex < 2e-2 and worst[0] < 2e-2
