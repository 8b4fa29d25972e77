"""GPU parity of the EVA02 tower (mico_b200.eva02_vit: RoPE, separate q / k / v projections, sub-LN, SwiGLU with a hidden
width that is not a multiple of 8) against the golden fixture produced by the unmodified reference EVAVisionTransformer in
its EVA02 configuration (oracle/make_golden.py:gen_eva02; width 128 = 2 heads x 64, depth 2, hidden 341, 257 tokens)."""
import os

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _tower(g, **kw):
    from mico_b200.eva02_vit import EVA02VisionTransformer
    c = g["cfg"]
    m = EVA02VisionTransformer(img_size=c["image"], patch_size=c["patch"], num_classes=0, embed_dim=c["width"], depth=c["depth"],
                               num_heads=c["heads"], mlp_ratio=c["mlp_ratio"], pt_hw_seq_len=c["pt_hw_seq_len"], eps=c["eps"], **kw)
    missing = m.load_state_dict(g["state_dict"], strict=False)
    assert not missing.unexpected_keys and all(".attn.rope." in k for k in missing.missing_keys)   # per-block aliases of rope.*
    return m.cuda()


def test_eva02_tower_matches_reference(golden_dir):
    g = torch.load(os.path.join(golden_dir, "eva02_tiny.pt"), weights_only=False)
    m = _tower(g).eval()
    assert torch.equal(m.rope.freqs_cos.cpu(), g["state_dict"]["rope.freqs_cos"])      # rope.py:79-136 tables, bit for bit
    y = m(g["x"].cuda(), return_all_features=True)
    e = rel_l2(y.detach().cpu(), g["y"])
    y.float().pow(2).mean().backward()
    errs = {k: rel_l2(p.grad.cpu(), g["grads"][k]) for k, p in m.named_parameters() if k in g["grads"]}
    worst = max(errs.items(), key=lambda kv: kv[1])
    print(f"eva02: y {e:.3e} worst grad {worst[0]} {worst[1]:.3e}; all: " + ", ".join(f"{k.split('.', 2)[-1]} {v:.1e}" for k, v in errs.items()))
    assert y.shape == (2, 257, 128) and len(errs) == len(g["grads"])
    assert e < 5e-3 and worst[1] < 2e-2
    cls = m(g["x"].cuda())                       # return_all_features=False -> cls token (num_classes = 0: no head)
    assert torch.equal(cls.detach(), y.detach()[:, 0])


def test_eva02_checkpointing_and_audio_input(golden_dir):
    """grad_checkpointing (eva_vit_model.py:635-637) reproduces the gradients (to fp32 summation order: the any-width LayerNorm
    backward accumulates its weight / bias gradients with atomics); a 3-D input is one channel replicated three times
    (mico.py:139-143)."""
    g = torch.load(os.path.join(golden_dir, "eva02_tiny.pt"), weights_only=False)
    m = _tower(g).train()
    x = g["x"].cuda()
    grads = []
    for ck in (False, True):
        m.set_grad_checkpointing(ck)
        m.zero_grad(set_to_none=True)
        m(x, return_all_features=True).pow(2).mean().backward()
        grads.append({k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 40
    for k in grads[0]:
        assert rel_l2(grads[1][k], grads[0][k]) < 1e-5, k
    spec = x[:, 0]
    with torch.no_grad():
        assert torch.equal(m(spec, return_all_features=True), m(spec[:, None].repeat(1, 3, 1, 1), return_all_features=True))


def test_mico_constructs_eva02_large_shapes():
    """MiCo(vision_encoder_type='evaclip02_large') builds the EVA02-CLIP-L-14 tower (width 1024, 24 blocks, hidden 2730) and
    runs its vision path (2 blocks here through the vision_tower_kwargs test hook)."""
    from mico_b200.mico import MiCo, _AttrDict
    cfg = _AttrDict(vision_encoder_type="evaclip02_large", vision_resolution=224, checkpointing=False, contra_dim=512,
                    max_vision_sample_num=2, max_audio_sample_num=1, max_depth_sample_num=1, beam_size=3, itm_ratio=0.1,
                    max_omni_caption_len=70, max_caption_len=40, max_subtitle_len=70, frame_embedding_type="adaptive",
                    pool_video=False, vision_tower_kwargs=dict(depth=2))
    torch.manual_seed(0)
    with torch.device("cuda"):
        m = MiCo.from_pretrained(cfg, {})
    m = m.cuda().eval()
    blk = m.vision_encoder.visual.blocks[0]
    assert m.vision_dim == 1024 and blk.mlp.w1.weight.shape == (2730, 1024) and blk.attn.num_heads == 16
    px = torch.randn(2, 2, 3, 224, 224, device="cuda")
    out = m.forward_vision_encoder(px)
    assert out.shape == (2, 2, 257, 1024) and torch.isfinite(out).all()
    pooled = m.pool_vision_for_contra(out) if hasattr(m, "pool_vision_for_contra") else None
    assert pooled is None or torch.isfinite(pooled).all()
