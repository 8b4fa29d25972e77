"""GPU parity of the BERT text/fusion encoder (mico_b200.bert) against the golden fixture produced by the unmodified
reference (tests/golden/bert_tiny.pt: hidden 128 = 2 heads x 64, 2 layers, cross-attention, tied decoder) and against
the CPU oracle at bert-base width.  Tolerances as in test_gpu_vit.py: features 5e-3, gradients 2e-2 rel-L2, loss 1e-3."""
import os

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
FEAT_TOL, GRAD_TOL, LOSS_TOL = 5e-3, 2e-2, 1e-3


def _model(cfgd):
    from mico_b200.bert import BertConfig, BertForMaskedLM
    return BertForMaskedLM(BertConfig(**cfgd))


def test_golden_text_only(golden_dir):
    g = torch.load(os.path.join(golden_dir, "bert_tiny.pt"), weights_only=False)
    m = _model(g["cfg"])
    m.load_state_dict(g["state_dict"], strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        h = m.bert(input_ids=g["ids"].cuda(), attention_mask=g["att"].cuda()).last_hidden_state
    e = rel_l2(h.cpu(), g["text_only"])
    print(f"bert text-only rel-L2 {e:.3e}")
    assert e < FEAT_TOL


def test_golden_caption_loss_and_grads(golden_dir):
    g = torch.load(os.path.join(golden_dir, "bert_tiny.pt"), weights_only=False)
    m = _model(g["cfg"])
    m.load_state_dict(g["state_dict"], strict=True)
    m = m.cuda().train()          # dropout probabilities are 0 in this fixture
    enc = g["enc"].cuda().requires_grad_(True)
    out = m(input_ids=g["ids"].cuda(), attention_mask=g["att3"].cuda(), encoder_hidden_states=enc, labels=g["labels"].cuda())
    assert rel_l2(out.sequence_output.detach().cpu(), g["sequence_output"]) < FEAT_TOL
    assert rel_l2(out.logits.detach().cpu(), g["logits"]) < FEAT_TOL
    assert abs(out.loss.item() - g["loss"].item()) <= LOSS_TOL * abs(g["loss"].item())
    out.loss.backward()
    assert rel_l2(enc.grad.cpu(), g["d_enc"]) < GRAD_TOL
    worst = ("", 0.0)
    for k, p in m.named_parameters():
        ref = g["grads"].get(k)
        if ref is None:
            continue
        if k.endswith("self.key.bias"):
            # softmax is invariant to a per-row constant q.b_k, so this gradient is exactly 0 in exact arithmetic:
            # both sides hold only rounding noise -- compare it with the query-bias gradient's scale instead
            qn = g["grads"][k.replace("key.bias", "query.bias")].norm().item()
            assert p.grad.norm().item() < 2e-2 * qn, (k, p.grad.norm().item(), qn)
            continue
        e = rel_l2(p.grad.cpu(), ref)
        worst = max(worst, (k, e), key=lambda t: t[1])
        assert e < GRAD_TOL, (k, e)
    print(f"bert caption: loss {out.loss.item():.5f} vs {g['loss'].item():.5f}; worst grad {worst[0]} {worst[1]:.3e}")


def test_attention_dropout_kernel_matches_masked_reference():
    """fused attention with probability dropout (fwd + bwd, tile kernels and the SIMT tail rows) vs an fp32 reference
    that applies the same counter-based mask to softmax(S) (bert.py:243-247)."""
    from mico_b200 import ops
    from oracle import bert as OB
    for (B, H, Sq, Sk, D) in [(2, 3, 40, 257, 64), (1, 2, 130, 130, 64)]:
        g = torch.Generator().manual_seed(Sq)
        q, k, v, do = (torch.randn(B, s_, H, D, generator=g).to(torch.bfloat16).cuda() for s_ in (Sq, Sk, Sk, Sq))
        p_, seed = 0.25, 123456789
        o, lse = ops.attention_fwd(q, k, v, D ** -0.5, dropout=(p_, seed))
        dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, do, D ** -0.5, dropout=(p_, seed))
        qf, kf, vf = (t.float().cpu().requires_grad_(True) for t in (q, k, v))
        pr = (torch.einsum("bihd,bjhd->bhij", qf, kf) * D ** -0.5).softmax(-1)
        pr = pr * OB.attn_drop_mult(p_, seed, B, H, Sq, Sk)
        ro = torch.einsum("bhij,bjhd->bihd", pr, vf)
        ro.backward(do.float().cpu())
        assert rel_l2(o.cpu(), ro) < 4e-3
        for a, r in ((dq, qf.grad), (dk, kf.grad), (dv, vf.grad)):
            assert rel_l2(a.cpu(), r) < 1e-2


def test_training_with_dropout_matches_masked_oracle(golden_dir):
    """Stock dropout probabilities (hidden 0.1, attention 0.1) in training mode: loss and gradients against the fp32 oracle
    driven by the same counter-based masks."""
    from oracle import bert as OB
    g = torch.load(os.path.join(golden_dir, "bert_tiny.pt"), weights_only=False)
    cfg = dict(g["cfg"], hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    m = _model(cfg)
    m.load_state_dict(g["state_dict"], strict=True)
    m = m.cuda().train()
    m.bert.dropout_seed = 20240607
    enc = g["enc"].cuda().requires_grad_(True)
    out = m(input_ids=g["ids"].cuda(), attention_mask=g["att3"].cuda(), encoder_hidden_states=enc, labels=g["labels"].cuda())
    out.loss.backward()
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in g["state_dict"].items()}
    p["cls.predictions.decoder.weight"] = p["bert.embeddings.word_embeddings.weight"]
    encr = g["enc"].clone().requires_grad_(True)
    loss, logits, seq = OB.masked_lm(p, g["ids"], g["att3"], encr, None, g["labels"], layers=2, heads=2,
                                     drop=(0.1, 0.1, 20240607))
    loss.backward()
    print(f"bert dropout: loss {out.loss.item():.5f} vs {loss.item():.5f} (no-dropout loss {g['loss'].item():.5f})")
    assert abs(loss.item() - g["loss"].item()) > 1e-3 * abs(g["loss"].item())      # the masks really are applied
    assert rel_l2(out.sequence_output.detach().cpu(), seq) < FEAT_TOL
    assert abs(out.loss.item() - loss.item()) <= LOSS_TOL * abs(loss.item())
    assert rel_l2(enc.grad.cpu(), encr.grad) < GRAD_TOL
    for k, v in m.named_parameters():
        if v.grad is None or k.endswith("self.key.bias") or k.endswith("decoder.weight"):
            continue
        assert rel_l2(v.grad.cpu(), p[k].grad) < GRAD_TOL, k


def test_bert_base_width_vs_oracle():
    """bert-base layer shapes (768, 12 heads x 64, FFN 3072, vocab 30522), 2 layers, S=40, S_k=257, 2-D masks both sides."""
    from oracle import bert as OB
    cfgd = dict(num_hidden_layers=2, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    torch.manual_seed(5)
    m = _model(cfgd)
    with torch.no_grad():
        for n, p in sorted(m.named_parameters()):
            if p.dim() == 1:
                p.add_(0.02 * torch.randn(p.shape))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m = m.cuda().train()
    g = torch.Generator().manual_seed(3)
    b, S, Sk = 2, 40, 257
    ids = torch.randint(1, 30522, (b, S), generator=g)
    att = (torch.arange(S)[None] < torch.tensor([40, 23])[:, None]).long()
    ids = ids * att
    enc = torch.randn(b, Sk, 768, generator=g)
    enc_att = (torch.arange(Sk)[None] < torch.tensor([257, 200])[:, None]).long()
    labels = torch.where((torch.rand(b, S, generator=g) < 0.6) & (att > 0), ids, torch.full_like(ids, -100))
    encg = enc.cuda().requires_grad_(True)
    out = m(input_ids=ids.cuda(), attention_mask=att.cuda(), encoder_hidden_states=encg, encoder_attention_mask=enc_att.cuda(),
            labels=labels.cuda())
    out.loss.backward()
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    p["cls.predictions.decoder.weight"] = p["bert.embeddings.word_embeddings.weight"]
    encr = enc.clone().requires_grad_(True)
    loss, logits, seq = OB.masked_lm(p, ids, att, encr, enc_att, labels, layers=2, heads=12)
    loss.backward()
    print(f"bert-base x2: seq {rel_l2(out.sequence_output.detach().cpu(), seq):.3e} logits "
          f"{rel_l2(out.logits.detach().cpu(), logits):.3e} loss {out.loss.item():.5f} vs {loss.item():.5f}")
    assert rel_l2(out.sequence_output.detach().cpu(), seq) < FEAT_TOL
    assert rel_l2(out.logits.detach().cpu(), logits) < FEAT_TOL
    assert abs(out.loss.item() - loss.item()) <= LOSS_TOL * abs(loss.item())
    assert rel_l2(encg.grad.cpu(), encr.grad) < GRAD_TOL
    for k, v in m.named_parameters():
        if v.grad is None or k.endswith("self.key.bias"):    # key-bias gradient is identically zero (see above)
            continue
        assert rel_l2(v.grad.cpu(), p[k].grad) < GRAD_TOL, k


def test_fused_lm_head_loss_matches_materialised_logits():
    """K7: the chunked LM-head + cross-entropy path (no [rows, 30522] logits; 8 vocabulary chunks of 4096) gives the loss and
    every gradient of the path that materialises the logits (same GEMMs, same bf16 dlogits), including ignored rows and a
    chunk boundary label; `.logits` of the lazy output equals the materialised tensor."""
    cfgd = dict(num_hidden_layers=1, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    torch.manual_seed(11)
    m = _model(cfgd).cuda().train()
    g = torch.Generator().manual_seed(12)
    b, S = 3, 24
    ids = torch.randint(1, 30522, (b, S), generator=g)
    labels = torch.where(torch.rand(b, S, generator=g) < 0.6, ids, torch.full_like(ids, -100))
    labels[0, 0], labels[0, 1], labels[1, 0], labels[2, 5] = 4095, 4096, 30521, 0       # chunk edges, last / first token
    res = {}
    for fused in (False, True):
        m.fused_lm_loss = fused
        m.zero_grad(set_to_none=True)
        out = m(input_ids=ids.cuda(), labels=labels.cuda())
        out.loss.backward()
        res[fused] = (out.loss.item(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None},
                      out.logits.detach().float().clone())
    l0, g0, lg0 = res[False]
    l1, g1, lg1 = res[True]
    assert abs(l1 - l0) <= 1e-5 * abs(l0), (l0, l1)
    assert torch.equal(lg0, lg1)
    assert g0.keys() == g1.keys()
    # key.bias: its gradient is exactly 0 in exact arithmetic (see test_golden_caption_loss_and_grads), both sides hold noise.
    # Elsewhere the two paths differ by bf16 rounding only: the materialised path rounds d(LN input) = dlogits . W to bf16 in
    # one GEMM, the chunked path accumulates the eight partial products in fp32 first.
    worst = max((rel_l2(g1[k], g0[k]), k) for k in g0 if not k.endswith("self.key.bias"))
    print(f"fused LM-head loss {l1:.6f} vs {l0:.6f}; worst gradient difference {worst[1]} {worst[0]:.2e}")
    assert worst[0] < 1e-2
    for k in ("cls.predictions.bias", "bert.embeddings.word_embeddings.weight", "cls.predictions.transform.dense.weight"):
        assert rel_l2(g1[k], g0[k]) < 2e-3, k          # the LM head's own parameters: same GEMMs, same bf16 dlogits
    # smaller chunks, not a divisor of the vocabulary
    m.lm_loss_chunk = 1000
    m.zero_grad(set_to_none=True)
    out = m(input_ids=ids.cuda(), labels=labels.cuda())
    assert abs(out.loss.item() - l0) <= 1e-5 * abs(l0)
