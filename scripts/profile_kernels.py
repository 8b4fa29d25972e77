"""Launch each hot kernel a few times on the ViT-g shapes (target for `ncu --set full -k regex:...`): 64 frames by default,
MICO_PROFILE_FRAMES=768 for the omni-modal step's single tower pass (64 samples x 12 frames = 197 376 tokens)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mico_b200 import ops
from mico_b200.ops import ACT_GELU_SAVE_GRAD, ACT_MUL_AUX, BF16, F32

which = sys.argv[1] if len(sys.argv) > 1 else "all"
M, D, F = int(os.environ.get("MICO_PROFILE_FRAMES", "64")) * 257, 1408, 6144
dev = "cuda"
r = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(BF16)
reps = 3
if which in ("all", "gemm"):
    x, w1, w2, a, dy = r(M, D), r(F, D), r(D, F), r(M, F), r(M, D)
    pre = r(M, F)
    bias = torch.randn(F, device=dev)
    for _ in range(reps):
        ops.gemm(x, w1, bias=bias, act=ACT_GELU_SAVE_GRAD, aux_out=pre)   # fc1 fwd (GELU + GELU' store)
        ops.gemm(dy, w2, b_mn=True, act=ACT_MUL_AUX, aux_in=pre)         # fc2 dgrad (* GELU')
        ops.gemm(dy, a, a_mn=True, b_mn=True, out_dtype=F32)             # fc2 wgrad
        ops.gemm(a, w2, out_dtype=F32, residual=torch.zeros(M, D, device=dev))  # fc2 fwd
    if os.environ.get("MICO_PROFILE_FRAMES"):     # round 2: the step's weakest GEMMs at the omni shapes
        wq, wp, dqkv = r(3 * D, D), r(D, D), r(M, 3 * D)
        res = torch.zeros(M, D, device=dev)
        for _ in range(reps):
            ops.gemm(x, wp, out_dtype=F32, bias=torch.zeros(D, device=dev), residual=res)   # proj fwd (+residual), K = 1408
            ops.gemm(dqkv, x, a_mn=True, b_mn=True, out_dtype=F32)                            # qkv wgrad
            ops.gemm(x, w1, bias=bias, act=ACT_GELU_SAVE_GRAD)                                # fc1 fwd, activation only
if which in ("all", "ln"):
    xf = torch.randn(M, D, device=dev)
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    dg, db = torch.empty_like(g), torch.empty_like(b)
    for _ in range(reps):
        yb, _, mean, rstd = ops.layernorm_fwd(xf, g, b, 1e-6)
        ops.layernorm_bwd(yb, xf, mean, rstd, g, dg, db, dres=xf, want_bf16=True)
if which in ("all", "attn"):
    B, H, S, d = 64, 16, 257, 88
    qkv = r(B, S, 3, H, d)
    do = r(B, S, H, d)
    dq = torch.empty_like(qkv)
    for _ in range(reps):
        o, lse = ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], d ** -0.5)
        ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, lse, do, d ** -0.5, dq=dq[:, :, 0], dk=dq[:, :, 1], dv=dq[:, :, 2])
if which in ("attn2",):
    # round 2: (a) ViT-g tower shape, plain; (b) fusion-encoder cross-attention of the omni step's largest group: 256 text
    # sequences x 12 heads x 128 queries against 64 shared K/V entries of 2827 visual tokens, attention dropout 0.1;
    # (c) fusion-encoder self-attention with a 3-D causal mask
    B, H, S, d = 64, 16, 257, 88
    qkv = r(B, S, 3, H, d)
    do = r(B, S, H, d)
    dq = torch.empty_like(qkv)
    o, lse = ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], d ** -0.5)
    ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, lse, do, d ** -0.5, dq=dq[:, :, 0], dk=dq[:, :, 1], dv=dq[:, :, 2])
    Bq, E, H, Sq, Sk, d = 256, 64, 12, 128, 2827, 64
    q, do = r(Bq, Sq, H, d), r(Bq, Sq, H, d)
    kv = r(E, Sk, 2, H, d)
    idx = (torch.arange(Bq, device=dev) % E).to(torch.int32)
    o, lse = ops.attention_fwd(q, kv[:, :, 0], kv[:, :, 1], d ** -0.5, dropout=(0.1, 11), kv_index=idx)
    ops.attention_bwd(q, kv[:, :, 0], kv[:, :, 1], o, lse, do, d ** -0.5, dropout=(0.1, 11), kv_index=idx)
    qkv = r(Bq, Sq, 3, H, d)
    mask = torch.zeros(Bq, Sq, Sq, device=dev).masked_fill_(torch.triu(torch.ones(Sq, Sq, device=dev), 1).bool(), -10000.0)
    o, lse = ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], d ** -0.5, mask=mask, dropout=(0.1, 12))
    ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, lse, do, d ** -0.5, mask=mask, dropout=(0.1, 12))
torch.cuda.synchronize()
