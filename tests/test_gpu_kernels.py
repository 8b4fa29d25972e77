"""GPU parity for the non-GEMM kernels: attention fwd, LayerNorm fwd/bwd, helpers.
References are plain PyTorch fp32 ops on the same (bf16-rounded) inputs.
Tolerances: bf16 outputs 4e-3 rel-L2 (one bf16 rounding of P and of the result), fp32 outputs 1e-5."""
import math

import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _randn(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def _attn_ref(q, k, v, scale, mask=None):
    # q,k,v: [B,S,H,D] fp32
    s = torch.einsum("bihd,bjhd->bhij", q, k) * scale
    if mask is not None:
        s = s + (mask[:, None, None, :] if mask.dim() == 2 else mask[:, None])
    p = s.softmax(-1)
    o = torch.einsum("bhij,bjhd->bihd", p, v)
    return o, torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,H,S,D", [(2, 2, 257, 88), (1, 3, 128, 64), (2, 4, 49, 32), (1, 2, 300, 128),
                                     (3, 16, 257, 88), (1, 2, 136, 64), (2, 2, 264, 88), (1, 1, 513, 32),
                                     (1, 2, 144, 88), (1, 2, 145, 64)])
def test_attention_fwd_fused_qkv(B, H, S, D):
    from mico_b200 import ops
    qkv = _randn((B, S, 3, H, D), 1, 1.0, torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    scale = D ** -0.5
    o, lse = ops.attention_fwd(q, k, v, scale)
    ro, rl = _attn_ref(q.float(), k.float(), v.float(), scale)
    torch.cuda.synchronize()
    print(f"attn fwd B{B} H{H} S{S} D{D}: o rel={rel_l2(o, ro):.3e} lse rel={rel_l2(lse, rl):.3e}")
    assert rel_l2(o, ro) < 4e-3
    assert rel_l2(lse, rl) < 1e-5


@pytest.mark.parametrize("Sq,Sk,three_d", [(128, 257, False), (40, 40, True), (70, 2056, False), (128, 128, True)])
def test_attention_fwd_bert_masks(Sq, Sk, three_d):
    """BERT: separate q / k / v projections, additive -10000 masks (bert.py:260, 697-781)."""
    from mico_b200 import ops
    B, H, D = 3, 12, 64
    q = _randn((B, Sq, H, D), 2, 1.0, torch.bfloat16)
    k = _randn((B, Sk, H, D), 3, 1.0, torch.bfloat16)
    v = _randn((B, Sk, H, D), 4, 1.0, torch.bfloat16)
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(max(1, Sk // 2), Sk + 1, (B,), generator=g)
    keep = (torch.arange(Sk)[None, :] < lens[:, None]).float()
    if three_d:
        keep = torch.tril(keep[:, None, :].expand(B, Sq, Sk))
    mask = ((1.0 - keep) * -10000.0).cuda()
    scale = 1.0 / math.sqrt(D)
    o, lse = ops.attention_fwd(q, k, v, scale, mask=mask)
    ro, rl = _attn_ref(q.float(), k.float(), v.float(), scale, mask)
    assert rel_l2(o, ro) < 4e-3
    assert rel_l2(lse, rl) < 1e-5


@pytest.mark.parametrize("M,D,eps", [(514, 1408, 1e-6), (384, 768, 1e-12), (100, 176, 1e-6), (49 * 4, 128, 1e-5)])
def test_layernorm_fwd_bwd(M, D, eps):
    from mico_b200 import ops
    x = _randn((M, D), 1) * 3 + 0.5
    gamma = 1 + _randn((D,), 2, 0.1)
    beta = _randn((D,), 3, 0.1)
    yb, yf, mean, rstd = ops.layernorm_fwd(x, gamma, beta, eps, out_bf16=True, out_f32=True)
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), gr, br, eps)
    assert rel_l2(yf, ref) < 1e-5
    assert rel_l2(yb, ref) < 3e-3
    assert rel_l2(mean, x.mean(-1)) < 1e-5
    dy = _randn((M, D), 4)
    dres = _randn((M, D), 5)
    ref.backward(dy)
    dgamma, dbeta = torch.empty_like(gamma), torch.empty_like(beta)
    T = 7 if M % 7 == 0 else 1
    rs = torch.rand(M // T, device="cuda") + 0.5
    dx, dxb = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, dres=dres, want_f32=True, want_bf16=True,
                                row_scale=rs, rows_per_group=T)
    assert rel_l2(dx, xr.grad + dres) < 1e-5
    assert rel_l2(dxb, (xr.grad + dres) * rs.repeat_interleave(T)[:, None]) < 3e-3
    assert rel_l2(dgamma, gr.grad) < 1e-5
    assert rel_l2(dbeta, br.grad) < 1e-5
    # bf16 dy + accumulate
    dyb = dy.to(torch.bfloat16)
    dg2, db2 = dgamma.clone(), dbeta.clone()
    ops.layernorm_bwd(dyb, x, mean, rstd, gamma, dg2, db2, accumulate=True)
    xr2 = x.clone().requires_grad_(True)
    g2 = gamma.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr2, (D,), g2, beta, eps).backward(dyb.float())
    assert rel_l2(dg2, dgamma + g2.grad) < 1e-5
    # fused bias gradient of the upstream linear layer: column sums of the scaled output (block-per-row kernel widths)
    if ops.layernorm_bwd_fuses_colsum(D):
        cs = torch.empty(D, device="cuda")
        dx3, dxb3 = ops.layernorm_bwd(dyb, x, mean, rstd, gamma, dg2, db2, dres=dres, want_bf16=True, row_scale=rs,
                                      rows_per_group=T, colsum_out=cs)
        want = ((xr2.grad + dres) * rs.repeat_interleave(T)[:, None]).sum(0)
        assert rel_l2(cs, want) < 1e-4
        assert rel_l2(cs, dxb3.float().sum(0)) < 2e-3


def test_helpers():
    from mico_b200 import ops
    x = _randn((1000, 1408), 1)
    assert torch.equal(ops.cast_bf16(x), x.to(torch.bfloat16))
    odd = _randn((12345,), 2)
    assert torch.equal(ops.cast_bf16(odd), odd.to(torch.bfloat16))
    xb = _randn((16448, 6144), 3, 1.0, torch.bfloat16)
    assert rel_l2(ops.colsum(xb), xb.float().sum(0)) < 1e-5
    xb2 = _randn((77, 4224), 4, 1.0, torch.bfloat16)
    acc = torch.ones(4224, device="cuda")
    assert rel_l2(ops.colsum(xb2, out=acc, accumulate=True), 1 + xb2.float().sum(0)) < 1e-5
    f = _randn((5, 257, 1408), 5)
    assert rel_l2(ops.batch_sum(f, 5), f.sum(0).flatten()) < 1e-6
    s = torch.rand(5, device="cuda")
    y = ops.scale_cast_bf16(f.view(-1, 1408), s, 257)
    assert torch.equal(y, (f * s[:, None, None]).view(-1, 1408).to(torch.bfloat16))


def test_patchify_matches_conv():
    from mico_b200 import ops
    B, P, W_ = 3, 14, 176
    img = _randn((B, 3, 224, 224), 1)
    cols = ops.patchify(img, P, 640)
    assert cols.shape == (B * 256, 640)
    assert cols[:, 588:].abs().max().item() == 0
    ref = torch.nn.functional.unfold(img, P, stride=P).transpose(1, 2).reshape(B * 256, 588)
    assert torch.equal(cols[:, :588], ref.to(torch.bfloat16))
    # audio: one channel replicated three times (mico.py:139-143)
    spec = _randn((B, 224, 224), 3)
    c2 = ops.patchify(spec, P, 640, replicate_channel=True)
    ref2 = torch.nn.functional.unfold(spec[:, None].repeat(1, 3, 1, 1), P, stride=P).transpose(1, 2).reshape(B * 256, 588)
    assert torch.equal(c2[:, :588], ref2.to(torch.bfloat16))


@pytest.mark.parametrize("B,H,Sq,Sk,D,masked", [(2, 2, 257, 257, 88, False), (1, 3, 128, 128, 64, False),
                                                (2, 12, 40, 257, 64, True), (2, 4, 49, 49, 32, False),
                                                (1, 2, 300, 200, 96, True), (3, 16, 257, 257, 88, False),
                                                (2, 12, 128, 128, 64, "3d"), (1, 2, 264, 136, 64, True),
                                                (2, 1, 129, 129, 32, "3d"), (1, 3, 144, 385, 88, False),
                                                (1, 2, 16, 2056, 64, True)])
def test_attention_bwd(B, H, Sq, Sk, D, masked):
    """dQ/dK/dV vs autograd of the fp32 reference on the same bf16 inputs.
    Tolerance 1e-2 rel-L2: P, dS and the outputs are each rounded to bf16 once (3 x 2^-9 in quadrature ~ 4e-3)."""
    from mico_b200 import ops
    q = _randn((B, Sq, H, D), 1, 1.0, torch.bfloat16)
    k = _randn((B, Sk, H, D), 2, 1.0, torch.bfloat16)
    v = _randn((B, Sk, H, D), 3, 1.0, torch.bfloat16)
    do = _randn((B, Sq, H, D), 4, 1.0, torch.bfloat16)
    mask = None
    if masked:
        g = torch.Generator().manual_seed(5)
        lens = torch.randint(max(1, Sk // 2), Sk + 1, (B,), generator=g)
        keep = (torch.arange(Sk)[None, :] < lens[:, None]).float()
        if masked == "3d":
            keep = torch.tril(keep[:, None, :].expand(B, Sq, Sk))
        mask = ((1.0 - keep) * -10000.0).cuda()
    scale = D ** -0.5
    o, lse = ops.attention_fwd(q, k, v, scale, mask=mask)
    dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, do, scale, mask=mask)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ro, _ = _attn_ref(qf, kf, vf, scale, mask)
    ro.backward(do.float())
    torch.cuda.synchronize()
    errs = [rel_l2(dq, qf.grad), rel_l2(dk, kf.grad), rel_l2(dv, vf.grad)]
    print(f"attn bwd B{B} H{H} Sq{Sq} Sk{Sk} D{D} mask={masked}: dq {errs[0]:.3e} dk {errs[1]:.3e} dv {errs[2]:.3e}")
    assert max(errs) < 1e-2


def test_attention_timing_vitg():
    """ViT-g shape at bs 64 (B*H = 1024 problems of 257 x 88): coarse timing print, fwd and bwd."""
    from mico_b200 import ops
    B, H, S, D = 64, 16, 257, 88
    qkv = _randn((B, S, 3, H, D), 1, 1.0, torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    do = _randn((B, S, H, D), 2, 1.0, torch.bfloat16)
    dqkv = torch.empty_like(qkv)
    scale = D ** -0.5
    o, lse = ops.attention_fwd(q, k, v, scale)
    for _ in range(2):
        ops.attention_fwd(q, k, v, scale, out=o)
        ops.attention_bwd(q, k, v, o, lse, do, scale, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(10):
        ops.attention_fwd(q, k, v, scale, out=o)
    ev[1].record()
    for _ in range(10):
        ops.attention_bwd(q, k, v, o, lse, do, scale, dq=dqkv[:, :, 0], dk=dqkv[:, :, 1], dv=dqkv[:, :, 2])
    ev[2].record()
    torch.cuda.synchronize()
    fl = 4.0 * B * H * S * S * D
    tf, tb = ev[0].elapsed_time(ev[1]) / 10, ev[1].elapsed_time(ev[2]) / 10
    print(f"attention ViT-g bs64: fwd {tf:.3f} ms ({fl / tf / 1e9:.0f} TFLOP/s)  bwd {tb:.3f} ms ({2.5 * fl / tb / 1e9:.0f} TFLOP/s)")


def test_colsum_two_ranges_single_pass():
    """q and v thirds of a fused [M, 3D] gradient in one launch (k skipped), repeated so that the self-resetting ticket
    counters of the single-pass reduction are exercised; ragged N through the plain entry point."""
    from mico_b200 import ops
    for M, D in ((16448, 1408), (514, 176), (77, 64)):
        x = _randn((M, 3 * D), 9, 1.0, torch.bfloat16)
        for _ in range(3):
            o0, o1 = torch.empty(D, device="cuda"), torch.empty(D, device="cuda")
            ops.colsum2(x, D, D, D, o0, o1)
            assert rel_l2(o0, x[:, :D].float().sum(0)) < 1e-5
            assert rel_l2(o1, x[:, 2 * D:].float().sum(0)) < 1e-5
    y = _randn((1000, 1001 + 7), 10, 1.0, torch.bfloat16)[:, :1001]
    for _ in range(2):
        assert rel_l2(ops.colsum(y), y.float().sum(0)) < 1e-5


@pytest.mark.parametrize("Sq,Sk,p_drop,masked", [(128, 771, 0.0, True), (40, 257, 0.25, True), (130, 200, 0.1, True),
                                                 (128, 771, 0.1, False), (72, 300, 0.25, False)])     # dropout-only kernels (kMode 2)
def test_attention_shared_kv_entries(Sq, Sk, p_drop, masked):
    """kv_index: several query batch entries read ONE K/V entry (fusion encoder: ITM + caption sequences of a sample).
    Forward equals attention against the expanded K/V; dK / dV are the sums over the readers (fp32 reference with the same
    dropout mask)."""
    from mico_b200 import ops
    from oracle import bert as OB
    B, E, H, D = 7, 3, 2, 64
    g = torch.Generator().manual_seed(Sq + Sk)
    q, do = (torch.randn(B, Sq, H, D, generator=g).to(torch.bfloat16).cuda() for _ in range(2))
    k, v = (torch.randn(E, Sk, H, D, generator=g).to(torch.bfloat16).cuda() for _ in range(2))
    idx = torch.tensor([0, 2, 1, 0, 0, 2, 1], dtype=torch.int32)
    mask = torch.zeros(B, Sk)
    mask[:, Sk - 5:] = -10000.0          # per-query-entry key padding
    mask[3, :7] = -10000.0
    if not masked:       # the fusion encoder's cross-attention over visual tokens: no mask, dropout on
        mask.zero_()
    drop = (p_drop, 99) if p_drop > 0 else None
    mk = mask.cuda() if masked else None
    o, lse = ops.attention_fwd(q, k, v, D ** -0.5, mask=mk, dropout=drop, kv_index=idx.cuda())
    dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, do, D ** -0.5, mask=mk, dropout=drop, kv_index=idx.cuda())
    assert dk.shape == k.shape and dv.shape == v.shape
    qf = q.float().cpu().requires_grad_(True)
    kf, vf = (t.float().cpu().requires_grad_(True) for t in (k, v))
    ke, ve = kf[idx.long()], vf[idx.long()]
    pr = (torch.einsum("bihd,bjhd->bhij", qf, ke) * D ** -0.5 + mask[:, None, None, :]).softmax(-1)
    if drop:
        pr = pr * OB.attn_drop_mult(p_drop, 99, B, H, Sq, Sk)
    ro = torch.einsum("bhij,bjhd->bihd", pr, ve)
    ro.backward(do.float().cpu())
    assert rel_l2(o.cpu(), ro) < 4e-3
    for a, r in ((dq, qf.grad), (dk, kf.grad), (dv, vf.grad)):
        assert rel_l2(a.cpu(), r) < 1e-2
    # an entry nobody reads gets zero gradients
    idx2 = torch.tensor([0, 0, 1, 0, 0, 1, 1], dtype=torch.int32).cuda()
    o2, lse2 = ops.attention_fwd(q, k, v, D ** -0.5, kv_index=idx2)
    _, dk2, dv2 = ops.attention_bwd(q, k, v, o2, lse2, do, D ** -0.5, kv_index=idx2)
    assert dk2[2].abs().max().item() == 0 and dv2[2].abs().max().item() == 0 and dk2[0].abs().max().item() > 0


@pytest.mark.parametrize("M,D", [(70, 2048), (33, 2730), (5, 6)])
def test_layernorm_any_width(M, D):
    """Rows wider than the register-resident kernels hold, or not a multiple of 4 (Swin-B patch merging 2048, EVA02-L SwiGLU
    2730): forward statistics / outputs and every backward term against torch's fp32 layer_norm."""
    from mico_b200 import ops
    g = torch.Generator().manual_seed(D)
    x = torch.randn(M, D, generator=g) * 2 + 0.5
    gam, bet = torch.randn(D, generator=g), torch.randn(D, generator=g)
    dy, dres = torch.randn(M, D, generator=g), torch.randn(M, D, generator=g)
    yb, yf, mean, rstd = ops.layernorm_fwd(x.cuda(), gam.cuda(), bet.cuda(), 1e-5, out_bf16=True, out_f32=True)
    xr = x.clone().requires_grad_(True)
    gr, br = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-5)
    ref.backward(dy)
    assert rel_l2(yf.cpu(), ref.detach()) < 1e-5 and rel_l2(yb.float().cpu(), ref.detach()) < 4e-3
    dg, db = torch.empty(D, device="cuda"), torch.empty(D, device="cuda")
    dx, dxb = ops.layernorm_bwd(dy.cuda(), x.cuda(), mean, rstd, gam.cuda(), dg, db, dres=dres.cuda(), want_bf16=True)
    assert rel_l2(dx.cpu(), xr.grad + dres) < 1e-5 and rel_l2(dxb.float().cpu(), xr.grad + dres) < 4e-3
    assert rel_l2(dg.cpu(), gr.grad) < 1e-5 and rel_l2(db.cpu(), br.grad) < 1e-5


@pytest.mark.parametrize("M,D", [(300, 768), (130, 1408)])
def test_layernorm_bwd_with_fused_dropout_mask(M, D):
    """mico_layernorm_bwd_dropout: the bf16 output and its column sums carry the forward pass's hidden-dropout mask (post-LN
    BERT: the gradient of the dense output and its bias gradient) -- against LayerNorm backward + mico_dropout + column sums."""
    from mico_b200 import ops
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, D, generator=g).cuda()
    dy = torch.randn(M, D, generator=g).cuda()
    dy2 = torch.randn(M, D, generator=g).to(torch.bfloat16).cuda()
    gamma = (torch.rand(D, generator=g) + 0.5).cuda()
    _, _, mean, rstd = ops.layernorm_fwd(x, gamma, torch.zeros(D, device="cuda"), 1e-12)
    p_, seed, site = 0.1, 424242, 7 << 40
    assert ops.layernorm_bwd_fuses_dropout(D, dy)
    dg0, db0, dg1, db1 = (torch.empty(D, device="cuda") for _ in range(4))
    cs = torch.empty(D, device="cuda")
    d32, d16 = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg1, db1, want_bf16=True, dy2=dy2, colsum_out=cs, dropout=(p_, seed, site))
    r32, r16 = ops.layernorm_bwd(dy, x, mean, rstd, gamma, dg0, db0, want_bf16=True, dy2=dy2)
    assert torch.equal(d32, r32) and torch.equal(dg0, dg1) and torch.equal(db0, db1)        # the fp32 side is untouched
    mult = ops.dropout(torch.ones(M, D, device="cuda"), p_, seed, site)[0]                  # the mask itself: 0 or 1 / (1 - p)
    ref = r32 * mult
    assert rel_l2(d16, ref) < 4e-3
    assert ((d16 == 0) == (mult == 0)).float().mean().item() > 0.999                        # same elements dropped
    assert 0.05 < (mult == 0).float().mean().item() < 0.15
    assert rel_l2(cs, ref.sum(0)) < 2e-3
