"""GPU parity: tcgen05 GEMM (mico_gemm_bf16) vs fp32 math on the same bf16-rounded operands.

Tolerance: fp32-output rel-L2 <= 2e-5 (only accumulation order differs);
bf16-output rel-L2 <= 3e-3 = one bf16 rounding (2^-9 relative) of the fp32 result.
"""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _mk(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


SHAPES = [
    # M, N, K
    (128, 128, 64), (128, 256, 128), (256, 176, 192), (257, 1408, 1408), (300, 4224, 1408),
    (1000, 6144, 1408), (514, 1408, 6144), (64, 512, 768), (96, 2, 768), (130, 30522, 768), (2056, 768, 1408),
]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_tn_fp32_out(M, N, K):
    from mico_b200 import ops
    a, b = _mk((M, K), 1), _mk((N, K), 2, 0.05)
    out = ops.gemm(a, b, out_dtype=torch.float32)
    ref = a.float() @ b.float().t()
    torch.cuda.synchronize()
    err = rel_l2(out, ref)
    print(f"TN {M}x{N}x{K}: rel_l2={err:.3e} max_abs={(out - ref).abs().max().item():.3e}")
    assert err < 2e-5


@pytest.mark.parametrize("a_mn,b_mn", [(False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 256), (1408, 6144, 1032), (4224, 1408, 520),
                                   (520, 1408, 4224), (136, 768, 200)])
def test_gemm_majors(M, N, K, a_mn, b_mn):
    from mico_b200 import ops
    a = _mk((K, M) if a_mn else (M, K), 3)
    b = _mk((K, N) if b_mn else (N, K), 4, 0.05)
    out = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    ref = A @ B
    torch.cuda.synchronize()
    err = rel_l2(out, ref)
    print(f"a_mn={a_mn} b_mn={b_mn} {M}x{N}x{K}: rel_l2={err:.3e}")
    assert err < 2e-5


def test_gemm_bias_gelu_aux_bf16():
    from mico_b200 import ops
    M, N, K = 514, 6144, 1408
    a, b = _mk((M, K), 5), _mk((N, K), 6, 0.03)
    bias = torch.randn(N, device="cuda") * 0.1
    aux = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU, aux_out=aux)
    pre = a.float() @ b.float().t() + bias
    ref = torch.nn.functional.gelu(pre)
    assert rel_l2(aux, pre) < 3e-3
    assert rel_l2(out, ref) < 3e-3
    # against the identically rounded reference the match must be ~exact
    assert rel_l2(out, ref.to(torch.bfloat16)) < 1e-3


def test_gemm_residual_rowscale_fp32():
    from mico_b200 import ops
    B_, T, N, K = 3, 257, 1408, 1408
    M = B_ * T
    a, b = _mk((M, K), 7), _mk((N, K), 8, 0.03)
    bias = torch.randn(N, device="cuda") * 0.1
    res = torch.randn(M, N, device="cuda")
    scale = torch.tensor([1.25, 0.0, 1.0], device="cuda")
    out = ops.gemm(a, b, bias=bias, residual=res, row_scale=scale, rows_per_group=T, out_dtype=torch.float32)
    ref = res + (a.float() @ b.float().t() + bias) * scale.repeat_interleave(T)[:, None]
    assert rel_l2(out, ref) < 2e-5
    # in-place on the residual stream (out aliases residual), as the block epilogue uses it
    x = res.clone()
    ops.gemm(a, b, bias=bias, residual=x, row_scale=scale, rows_per_group=T, out=x)
    assert rel_l2(x, ref) < 2e-5


def test_gemm_gelu_bwd_and_accumulate():
    from mico_b200 import ops
    M, N, K = 514, 6144, 1408
    dy, w = _mk((M, K), 9), _mk((K, N), 10, 0.03)          # w is [K rows][N] -> MN-major B
    u = _mk((M, N), 11)
    out = ops.gemm(dy, w, b_mn=True, act=ops.ACT_GELU_BWD, aux_in=u)
    uf = u.float().requires_grad_(True)
    torch.nn.functional.gelu(uf).backward(dy.float() @ w.float())
    assert rel_l2(out, uf.grad) < 3e-3
    # accumulate into fp32 (gradient accumulation for wgrad)
    acc = torch.randn(M, N, device="cuda")
    ref = acc + 0.5 * (dy.float() @ w.float())
    ops.gemm(dy, w, b_mn=True, out=acc, accumulate=True, alpha=0.5)
    assert rel_l2(acc, ref) < 2e-5


@pytest.mark.parametrize("quick", [False, True])
@pytest.mark.parametrize("M,N,K", [(514, 6144, 1408), (100, 200, 72)])
def test_gemm_act_save_grad_then_mul_aux(M, N, K, quick):
    """fc1 epilogue stores act'(x) (GELU or QuickGELU); the fc2-dgrad epilogue multiplies by it (MUL_AUX)."""
    from mico_b200 import ops
    a, b = _mk((M, K), 21), _mk((N, K), 22, 0.03)
    bias = torch.randn(N, device="cuda") * 0.1
    aux = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    act = ops.ACT_QUICK_GELU_SAVE_GRAD if quick else ops.ACT_GELU_SAVE_GRAD
    out = ops.gemm(a, b, bias=bias, act=act, aux_out=aux)
    pre = (a.float() @ b.float().t() + bias).requires_grad_(True)
    ref = pre * torch.sigmoid(1.702 * pre) if quick else torch.nn.functional.gelu(pre)
    ref.backward(torch.ones_like(ref))
    assert rel_l2(out, ref.detach()) < 3e-3
    assert rel_l2(aux, pre.grad) < 3e-3
    dy, w2 = _mk((M, 136), 23), _mk((136, N), 24, 0.03)
    d = ops.gemm(dy, w2, b_mn=True, act=ops.ACT_MUL_AUX, aux_in=aux)
    assert rel_l2(d, (dy.float() @ w2.float()) * aux.float()) < 3e-3


def test_gemm_patch_remap():
    from mico_b200 import ops
    B_, P, T, N, K = 2, 256, 257, 1408, 640
    a, b = _mk((B_ * P, K), 12), _mk((N, K), 13, 0.03)
    bias = torch.randn(N, device="cuda") * 0.1
    pos = torch.randn(T, N, device="cuda")
    out = torch.zeros(B_ * T, N, device="cuda")
    ops.gemm(a, b, bias=bias, residual=pos, out=out, remap=(P, T, 1), residual_bcast=True)
    ref = torch.zeros(B_, T, N, device="cuda")
    ref[:, 1:] = (a.float() @ b.float().t() + bias).view(B_, P, N) + pos[1:]
    assert rel_l2(out, ref.view(-1, N)) < 2e-5
    assert out.view(B_, T, N)[:, 0].abs().max().item() == 0.0


def test_gemm_full_size_fc1_timing():
    """ViT-g fc1 at bs 64: parity at full size through linearity (gemm(a1+a2) == gemm(a1)+gemm(a2) for
    operands whose sum is exact in bf16) plus a coarse TFLOP/s print."""
    from mico_b200 import ops
    M, N, K = 16448, 6144, 1408
    g = torch.Generator().manual_seed(0)
    a1 = torch.randint(-8, 8, (M, K), generator=g).to(torch.bfloat16).cuda()
    a2 = torch.randint(-8, 8, (M, K), generator=g).to(torch.bfloat16).cuda()
    w = torch.randint(-4, 4, (N, K), generator=g).to(torch.bfloat16).cuda()
    o1 = ops.gemm(a1, w, out_dtype=torch.float32)
    o2 = ops.gemm(a2, w, out_dtype=torch.float32)
    o12 = ops.gemm(a1 + a2, w, out_dtype=torch.float32)
    assert torch.equal(o1 + o2, o12)          # small integers: exact in fp32
    assert torch.equal(o1[:512], (a1[:512].float() @ w.float().t()))
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(a1, w, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(a1, w, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"fc1 16448x6144x1408: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s")


@pytest.mark.parametrize("M,N,K,b_mn", [(18944, 176, 128, False),      # 148 M tiles: the cost model picks 176-wide tiles (16-col tail chunk)
                                        (16448, 1408, 256, False),     # 256-wide tiles, last N tile issues a 128-wide MMA
                                        (16448, 1408, 256, True),      # same with an MN-major B (dgrad layout)
                                        (700, 352, 192, False)])       # partial last M tile, narrow last N tile (96 of 256)
def test_specialised_epilogues_all_kinds(M, N, K, b_mn):
    """Every specialised epilogue kernel (bf16 +/- bias, GELU + GELU' store, x saved derivative, fp32 residual with DropPath
    row scale, fp32) on tile shapes that exercise the narrow last-N MMA, the 176-wide tile's 16-column tail chunk and a
    ragged last M tile (rows clipped by the TMA store), against fp32 math on the same bf16 operands."""
    from mico_b200 import ops
    a = _mk((M, K), 11)
    b = _mk((K, N) if b_mn else (N, K), 12, 0.05)
    acc = a.float() @ (b.float() if b_mn else b.float().t())
    bias = torch.randn(N, device="cuda") * 0.1
    kw = dict(b_mn=b_mn)
    assert rel_l2(ops.gemm(a, b, **kw), acc) < 4e-3                                         # EPI_BF16
    assert rel_l2(ops.gemm(a, b, bias=bias, **kw), acc + bias) < 4e-3                       # EPI_BF16 + bias
    assert rel_l2(ops.gemm(a, b, out_dtype=torch.float32, **kw), acc) < 2e-5                # EPI_F32
    aux = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU_SAVE_GRAD, aux_out=aux, **kw)          # EPI_GELU_SAVE
    pre = (acc + bias).requires_grad_(True)
    ref = torch.nn.functional.gelu(pre)
    ref.sum().backward()
    assert rel_l2(out, ref.detach()) < 4e-3 and rel_l2(aux, pre.grad) < 4e-3
    sav = _mk((M, N), 13)
    assert rel_l2(ops.gemm(a, b, act=ops.ACT_MUL_AUX, aux_in=sav, **kw), acc * sav.float()) < 4e-3     # EPI_MUL_AUX
    T = 37
    res = torch.randn(M, N, device="cuda")
    rs = torch.rand((M + T - 1) // T, device="cuda") + 0.5
    out = ops.gemm(a, b, bias=bias, residual=res, row_scale=rs, rows_per_group=T, out_dtype=torch.float32, **kw)   # EPI_RES32
    ref = res + (acc + bias) * rs.repeat_interleave(T)[:M, None]
    assert rel_l2(out, ref) < 2e-5


@pytest.mark.parametrize("M,N,K", [(768, 768, 8192), (1408, 1408, 16448), (2304, 768, 4096 + 64)])
def test_wgrad_split_k(M, N, K):
    """Weight gradients with few output tiles and a long K loop are split over K (TMA reduce-add of fp32 partial tiles into the
    zeroed output, gemm.cu launch_gemm): dW = dY^T X against an fp32 reference on the same bf16 operands, and against the
    unsplit kernel (MICO_GEMM_SPLITK_MAX is read once per process, so the unsplit result comes from an accumulate launch)."""
    from mico_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    dy = torch.randn(K, M, generator=g).to(torch.bfloat16).cuda()
    x = torch.randn(K, N, generator=g).to(torch.bfloat16).cuda()
    out = torch.full((M, N), 7.0, device="cuda")           # stale contents must not leak into the result
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=out)
    ref = dy.float().t() @ x.float()
    assert rel_l2(out, ref) < 2e-5
    acc = torch.zeros((M, N), device="cuda")
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=acc, accumulate=True)      # generic epilogue, never split
    assert rel_l2(out, acc) < 2e-5       # different fp32 partial-sum grouping over K


@pytest.mark.parametrize("M,N,K,wgrad", [(37000, 1408, 192, False),     # 145 pair groups x (5 full + 1 half tile): 12 rounds of the balanced schedule
                                         (37000, 1408, 192, "dgrad"),   # MN-major B
                                         (4224, 1408, 8192, True),      # qkv weight gradient: split K, last M group half empty
                                         (6144, 1408, 4096, True)])     # fc1 weight gradient
def test_balanced_unit_schedule_exact(M, N, K, wgrad):
    """The balanced work-unit schedule (gemm.cu:GemmPlan: each round's units dealt longest-first to the least-loaded CTA pairs,
    split-K chosen by planned makespan) must cover every (tile, K part) exactly once: small-integer operands make every
    partial sum exact in fp32, so the result equals the fp32 product bit for bit whatever the unit order."""
    from mico_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    if wgrad is True:
        dy = torch.randint(-4, 4, (K, M), generator=g).to(torch.bfloat16).cuda()
        x = torch.randint(-4, 4, (K, N), generator=g).to(torch.bfloat16).cuda()
        out = torch.full((M, N), 3.0, device="cuda")
        ops.gemm(dy, x, a_mn=True, b_mn=True, out=out)
        assert torch.equal(out, dy.float().t() @ x.float())
        return
    a = torch.randint(-8, 8, (M, K), generator=g).to(torch.bfloat16).cuda()
    if wgrad == "dgrad":
        w = torch.randint(-4, 4, (K, N), generator=g).to(torch.bfloat16).cuda()
        ref = a.float() @ w.float()
        out = ops.gemm(a, w, b_mn=True, out_dtype=torch.float32)
    else:
        w = torch.randint(-4, 4, (N, K), generator=g).to(torch.bfloat16).cuda()
        ref = a.float() @ w.float().t()
        out = ops.gemm(a, w, out_dtype=torch.float32)
    assert torch.equal(out, ref)
    res = torch.randint(-64, 64, (M, N), generator=g).float().cuda()
    out2 = ops.gemm(a, w, b_mn=wgrad == "dgrad", residual=res, out_dtype=torch.float32)        # fp32 residual epilogue
    assert torch.equal(out2, ref + res)


@pytest.mark.parametrize("M,N,K", [(6144, 1408, 4096), (4224, 1408, 8192 + 72), (1408, 1408, 1000), (512, 384, 333)])
def test_wgrad_returns_bias_gradient(M, N, K):
    """asum_out: the weight-gradient GEMM dW = dY^T X also returns the bias gradient sum_t dY[t, :] from 32 extra MMA columns
    of its last N tile (B operand = a tile of ones), so the tower backward no longer re-reads dY for a column sum.  Small
    integers: both results are exact in fp32 whatever the split-K order."""
    from mico_b200 import ops
    assert ops.gemm_fuses_asum(M, N)
    g = torch.Generator().manual_seed(M + K)
    dy = torch.randint(-4, 4, (K, M), generator=g).to(torch.bfloat16).cuda()
    x = torch.randint(-4, 4, (K, N), generator=g).to(torch.bfloat16).cuda()
    out = torch.full((M, N), 3.0, device="cuda")
    bsum = torch.full((M,), 5.0, device="cuda")
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=out, asum_out=bsum)
    assert torch.equal(out, dy.float().t() @ x.float())
    assert torch.equal(bsum, dy.float().sum(0))
    # real-valued operands against the separate column-sum kernel
    dy = (torch.randn(K, M, generator=g) * 0.3).to(torch.bfloat16).cuda()
    ops.gemm(dy, x, a_mn=True, b_mn=True, out=out, asum_out=bsum)
    assert rel_l2(bsum, ops.colsum(dy)) < 1e-5
    assert not ops.gemm_fuses_asum(768, 768) and not ops.gemm_fuses_asum(128, 1408)
