#!/usr/bin/env python
"""Tiny-shape sweep of every kernel family behind the C-ABI, meant to run UNDER compute-sanitizer (scripts/sanitize.sh):
tail shapes of the 2-CTA pair GEMM and its epilogues, masked / dropout / shared-K/V attention forward + backward with
remainder rows, LayerNorm, embedding, cross-entropy, fbank, image preprocessing, AdamW.  Shapes are small on purpose (the
sanitizer slows kernels down by 10-1000x); correctness is covered by tests/, this only has to EXECUTE every code path."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mico_b200 import ops
from mico_b200.ops import ACT_GELU_SAVE_GRAD, ACT_MUL_AUX, BF16, F32


def gemms():
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    for (M, N, K) in [(200, 1408, 192), (130, 256, 64), (257, 176, 176), (64, 24, 136), (300, 384, 128)]:
        a, b = r(M, K).to(BF16), r(N, K).to(BF16)
        bias, res = r(N), r(M, N)
        ops.gemm(a, b, bias=bias)                                              # bf16 epilogue
        ops.gemm(a, b, out_dtype=F32, bias=bias, residual=res, row_scale=r((M + 63) // 64).abs(), rows_per_group=64)
        pre = torch.empty(M, N, device="cuda", dtype=BF16)
        y = ops.gemm(a, b, bias=bias, act=ACT_GELU_SAVE_GRAD, aux_out=pre)     # GELU + GELU' store
        ops.gemm(y, b, b_mn=True, act=ACT_MUL_AUX, aux_in=r(M, K).to(BF16))    # dgrad x saved GELU'
        ops.gemm(y, a, a_mn=True, b_mn=True, out_dtype=F32)                    # wgrad (MN-major operands)
        ops.gemm(y, a, a_mn=True, b_mn=True, out=torch.zeros(N, K, device="cuda"), accumulate=True)
    torch.cuda.synchronize()


def gemms_r2n():
    """End of round 2: the balanced unit schedule over several rounds (84 units on 74 CTA pairs, narrow last N tile), split-K
    weight gradients dealt by the plan, the bias gradient from the weight-gradient GEMM's ones columns (asum_out), and the
    dropout-only attention kernels (head_dim 64, no mask)."""
    g = torch.Generator(device="cuda").manual_seed(5)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    a, b = r(3400, 64).to(BF16), r(1408, 64).to(BF16)
    ops.gemm(a, b, bias=r(1408))                                                # 14 pair groups x 6 N tiles: two rounds
    ops.gemm(a, b, out_dtype=F32, bias=r(1408), residual=r(3400, 1408))
    ops.gemm(a, r(64, 1408).to(BF16), b_mn=True)
    for (M, N, K) in [(512, 384, 333), (1408, 1408, 2048), (4224, 1408, 1100)]:
        dy, x = r(K, M).to(BF16), r(K, N).to(BF16)
        out, bsum = torch.empty(M, N, device="cuda"), torch.empty(M, device="cuda")
        ops.gemm(dy, x, a_mn=True, b_mn=True, out=out)                           # planned split-K
        ops.gemm(dy, x, a_mn=True, b_mn=True, out=out, asum_out=bsum)            # + bias gradient
    # LayerNorm backward carrying the hidden-dropout mask + bias column sums (post-LN BERT)
    x, dy = r(130, 768), r(130, 768)
    gam = r(768).abs() + 0.5
    _, _, mean, rstd = ops.layernorm_fwd(x, gam, torch.zeros(768, device="cuda"), 1e-12)
    ops.layernorm_bwd(dy, x, mean, rstd, gam, torch.empty(768, device="cuda"), torch.empty(768, device="cuda"), want_bf16=True,
                      dy2=r(130, 768).to(BF16), colsum_out=torch.empty(768, device="cuda"), dropout=(0.1, 5, 3 << 40))
    q, do = r(3, 128, 2, 64).to(BF16), r(3, 128, 2, 64).to(BF16)
    k, v = r(3, 300, 2, 64).to(BF16), r(3, 300, 2, 64).to(BF16)
    o, lse = ops.attention_fwd(q, k, v, 0.125, dropout=(0.1, 4))
    ops.attention_bwd(q, k, v, o, lse, do, 0.125, dropout=(0.1, 4))
    torch.cuda.synchronize()


def attention():
    g = torch.Generator(device="cuda").manual_seed(1)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g).to(BF16)
    for (B, H, Sq, Sk, D, masked, drop, shared) in [(2, 2, 257, 257, 88, False, None, False), (2, 2, 40, 257, 64, True, (0.1, 7), False),
                                                     (3, 2, 128, 200, 64, False, (0.1, 9), True), (4, 4, 49, 49, 32, True, None, False),
                                                     (1, 2, 130, 130, 64, True, None, False),
                                                     # head_dim 88 with a mask + dropout and a ragged key count: the pipelined dQ kernel's
                                                     # predicated path (attention_bwd_dq.cu) next to the single-buffer dK/dV kernel
                                                     (2, 2, 257, 200, 88, True, (0.1, 3), False), (1, 2, 300, 97, 88, False, None, False)]:
        E = 2 if shared else B
        q, do = r(B, Sq, H, D), r(B, Sq, H, D)
        k, v = r(E, Sk, H, D), r(E, Sk, H, D)
        mask = None
        if masked:
            mask = torch.zeros(B, Sq, Sk, device="cuda")
            mask[:, :, Sk - 3:] = -10000.0
        idx = torch.tensor([0, 1, 0][:B], dtype=torch.int32, device="cuda") if shared else None
        o, lse = ops.attention_fwd(q, k, v, D ** -0.5, mask=mask, dropout=drop, kv_index=idx)
        ops.attention_bwd(q, k, v, o, lse, do, D ** -0.5, mask=mask, dropout=drop, kv_index=idx)
    # Swin: per-(window, head) bias + its gradient
    B, H, S, D = 8, 2, 49, 32
    q, k, v, do = (r(B, S, H, D) for _ in range(4))
    bias = torch.randn(4, H, S, S, device="cuda")
    o, lse = ops.attention_fwd(q, k, v, D ** -0.5, mask=bias)
    ops.attention_bwd(q, k, v, o, lse, do, D ** -0.5, mask=bias, dmask=torch.zeros_like(bias))
    torch.cuda.synchronize()


def rowwise():
    g = torch.Generator(device="cuda").manual_seed(2)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    for (M, D) in [(300, 1408), (130, 768), (77, 176), (50, 96)]:
        x, gam, bet = r(M, D), r(D), r(D)
        yb, yf, mean, rstd = ops.layernorm_fwd(x, gam, bet, 1e-6, out_bf16=True, out_f32=True)
        dg, db = torch.empty(D, device="cuda"), torch.empty(D, device="cuda")
        cs = torch.empty(D, device="cuda") if ops.layernorm_bwd_fuses_colsum(D) else None
        ops.layernorm_bwd(r(M, D), x, mean, rstd, gam, dg, db, dres=r(M, D), want_bf16=True, row_scale=r((M + 9) // 10),
                          rows_per_group=10, dy2=r(M, D).to(BF16), colsum_out=cs)
        ops.colsum(yb)
        ops.scale_cast_bf16(x)
        ops.cast_bf16(x)
        ops.scale_(x.clone(), scale_dev=torch.tensor([0.5], device="cuda"))
        ops.dropout(x, 0.1, 5, 0, res=x, out_f32=True, out_bf16=True)
    ops.colsum2(r(200, 528).to(BF16), 176, 176, 176, torch.empty(176, device="cuda"), torch.empty(176, device="cuda"))
    ops.batch_sum(r(5, 1000), 5)
    img = r(3, 3, 224, 224)
    cols = ops.patchify(img, 14, 640, tokens_per_img=257, token_off=1)
    ops.patchify(r(2, 224, 224), 14, 640, replicate_channel=True, tokens_per_img=257, token_off=1)
    ops.drop_path_scales(torch.tensor([0.0, 0.2, 0.4], device="cuda"), 5, 1234, 1)
    torch.cuda.synchronize()


def heads_and_io():
    g = torch.Generator(device="cuda").manual_seed(3)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    V, D, S = 1000, 128, 24
    ids = torch.randint(0, V, (4 * S,), device="cuda")
    x = ops.embedding_gather(ids, r(V, D), r(64, D), r(2, D), S)
    ops.embedding_scatter_add(x, ids, torch.zeros(V, D, device="cuda"))
    logits = r(4 * S, V)
    labels = torch.randint(0, V, (4 * S,), device="cuda")
    labels[::3] = -100
    stats, lse = ops.cross_entropy_fwd(logits, labels, label_smoothing=0.1)
    ops.cross_entropy_bwd(logits, labels, lse, torch.ones(1, device="cuda"), stats, label_smoothing=0.1, out_dtype=BF16)
    # K7: chunked LM-head loss (three vocabulary chunks of 400 / 400 / 200)
    M = 4 * S
    rm, rs_, ll = torch.empty(M, device="cuda"), torch.empty(M, device="cuda"), torch.zeros(M, device="cuda")
    for c0 in range(0, V, 400):
        ops.ce_chunk_update(logits[:, c0:min(V, c0 + 400)], c0, labels, rm, rs_, ll, first=(c0 == 0))
    st2, lse2 = ops.ce_chunk_finalize(rm, rs_, ll, labels, V)
    dl = torch.empty(M, 400, device="cuda", dtype=BF16)
    for c0 in range(0, V, 400):
        ops.ce_chunk_grad(logits[:, c0:min(V, c0 + 400)], c0, labels, lse2, torch.ones(1, device="cuda"), st2, V, dl)
    # EVA02: rotary embedding (forward / inverse) and SwiGLU (forward / backward)
    cos, sin = r(16, 64), r(16, 64)
    xb = ops.rope(r(2, 17, 2, 64), cos, sin)
    ops.rope(xb, cos, sin, inverse=True)
    ops.swiglu(r(50, 344), r(50, 344))
    ops.swiglu(r(50, 344), r(50, 344), r(50, 344))
    y, nrm = ops.l2norm_fwd(r(9, 512))
    ops.l2norm_bwd(y, r(9, 512), nrm)
    ops.sgemm(r(9, 512), r(17, 512), alpha_dev=torch.tensor([0.07], device="cuda"), alpha_recip=True)
    ops.gelu_f32(r(33, 77))
    from mico_b200.audioprocessor import AudioProcessor
    AudioProcessor(melbins=224, target_length=224, sample_num=3, training=False, device="cuda").batch(0.1 * r(2, 16000))
    from mico_b200 import optim
    ps = [torch.nn.Parameter(r(n)) for n in (33, 4096, 17000)]
    for p in ps:
        p.grad = r(p.numel())
    optim.AdamW(ps, lr=1e-3, weight_decay=0.01).step()
    torch.cuda.synchronize()


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "attention", "rowwise", "heads"]
    for w in which:
        dict(gemm=gemms, gemm_r2n=gemms_r2n, attention=attention, rowwise=rowwise, heads=heads_and_io)[w]()
        print(f"[sanitize_sweep] {w}: done", flush=True)
