"""GPU parity of the small head / loss kernels (mico_b200.functional) against the same torch ops in fp32."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _r(shape, seed, scale=1.0):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).cuda()


@pytest.mark.parametrize("M,V,ls", [(96, 768, 0.1), (40, 30522, 0.0), (17, 2, 0.0), (32, 256, 0.1)])
def test_cross_entropy(M, V, ls):
    from mico_b200 import functional as MF
    x = _r((M, V), 1, 3.0).requires_grad_(True)
    y = torch.randint(0, V, (M,), generator=torch.Generator().manual_seed(2)).cuda()
    y[::5] = -100
    loss = MF.cross_entropy(x, y, label_smoothing=ls)
    xr = x.detach().clone().requires_grad_(True)
    ref = F.cross_entropy(xr, y, label_smoothing=ls)
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    (loss * 1.7).backward()
    (ref * 1.7).backward()
    assert rel_l2(x.grad, xr.grad) < 1e-5


def test_linear_f32_layernorm_gelu_normalize():
    from mico_b200 import functional as MF
    x = _r((48, 1408), 1).requires_grad_(True)
    w = _r((512, 1408), 2, 0.02).requires_grad_(True)
    b = _r((512,), 3, 0.1).requires_grad_(True)
    g, be = (1 + _r((512,), 4, 0.1)).requires_grad_(True), _r((512,), 5, 0.1).requires_grad_(True)
    y = MF.normalize(MF.layer_norm(MF.gelu(MF.linear_f32(x, w, b)), g, be, 1e-12))
    xr, wr, br, gr, ber = (t.detach().clone().requires_grad_(True) for t in (x, w, b, g, be))
    yr = F.normalize(F.layer_norm(F.gelu(F.linear(xr, wr, br)), (512,), gr, ber, 1e-12), dim=-1)
    assert rel_l2(y, yr) < 1e-5
    dy = _r((48, 512), 6)
    y.backward(dy)
    yr.backward(dy)
    for a, r in ((x, xr), (w, wr), (b, br), (g, gr), (be, ber)):
        assert rel_l2(a.grad, r.grad) < 2e-5


def test_contrastive_logits_and_temperature_grad():
    from mico_b200 import functional as MF
    a = F.normalize(_r((32, 512), 1), dim=-1).requires_grad_(True)
    ball = F.normalize(_r((256, 512), 2), dim=-1)
    temp = torch.tensor(0.07, device="cuda", requires_grad=True)
    sim = MF.contrastive_logits(a, ball, temp)
    ar, tr = a.detach().clone().requires_grad_(True), temp.detach().clone().requires_grad_(True)
    simr = ar @ ball.t() / tr
    assert rel_l2(sim, simr) < 1e-5
    tgt = torch.arange(32, device="cuda") + 64
    MF.cross_entropy(sim, tgt, label_smoothing=0.1).backward()
    F.cross_entropy(simr, tgt, label_smoothing=0.1).backward()
    assert rel_l2(a.grad, ar.grad) < 1e-5
    assert abs(temp.grad.item() - tr.grad.item()) < 1e-4 * abs(tr.grad.item())


def test_linear_tc_matches_fp32_within_bf16():
    from mico_b200 import functional as MF
    x = _r((2, 257, 1408), 1).requires_grad_(True)
    w = _r((768, 1408), 2, 0.02).requires_grad_(True)
    b = _r((768,), 3, 0.1).requires_grad_(True)
    y = MF.linear(x, w, b)
    xr, wr, br = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    yr = F.linear(xr, wr, br)
    assert y.shape == yr.shape and rel_l2(y, yr) < 4e-3
    dy = _r(tuple(yr.shape), 4)
    y.backward(dy)
    yr.backward(dy)
    for a, r in ((x, xr), (w, wr), (b, br)):
        assert rel_l2(a.grad, r.grad) < 6e-3


def test_embedding_gather_and_scatter():
    from mico_b200 import ops
    V, P, D, b, S = 1000, 64, 128, 3, 24
    word, pos, typ = _r((V, D), 1), _r((P, D), 2), _r((2, D), 3)
    ids = torch.randint(0, V, (b, S), generator=torch.Generator().manual_seed(4)).cuda()
    x = ops.embedding_gather(ids.view(-1), word, pos, typ, S)
    ref = word[ids] + typ[0] + pos[:S]
    assert torch.allclose(x.view(b, S, D), ref, atol=1e-6)
    dx = _r((b * S, D), 5)
    gw = torch.zeros_like(word)
    ops.embedding_scatter_add(dx, ids.view(-1), gw)
    refg = torch.zeros_like(word).index_add_(0, ids.view(-1), dx)
    assert rel_l2(gw, refg) < 1e-6
    # a sequence longer than the position table fails loudly, like nn.Embedding (the kernel itself would clamp)
    long_ids = torch.zeros(P + 1, dtype=torch.int64, device="cuda")
    with pytest.raises(Exception, match="position embeddings"):
        ops.embedding_gather(long_ids, word, pos, typ, P + 1)
