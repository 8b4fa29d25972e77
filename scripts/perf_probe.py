"""Per-kernel timing probe on the ViT-g bs=64 shapes (CUDA events, after warm-up).  Developer tool."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mico_b200 import ops
from mico_b200.ops import ACT_GELU, ACT_GELU_BWD, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX, BF16, F32


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    M, D, F = 64 * 257, 1408, 6144
    dev = "cuda"
    r = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(BF16)
    x, w_qkv, w_p, w1, w2 = r(M, D), r(3 * D, D), r(D, D), r(F, D), r(D, F)
    a = r(M, F)
    pre = r(M, F)
    dy = r(M, D)
    dqkv = r(M, 3 * D)
    bias_f = torch.randn(F, device=dev)
    bias_d = torch.randn(D, device=dev)
    bias_q = torch.randn(3 * D, device=dev)
    res = torch.randn(M, D, device=dev)
    out_f = torch.empty(M, F, device=dev, dtype=BF16)
    out_q = torch.empty(M, 3 * D, device=dev, dtype=BF16)
    out_d32 = torch.empty(M, D, device=dev, dtype=F32)
    out_d = torch.empty(M, D, device=dev, dtype=BF16)
    gw1 = torch.empty(F, D, device=dev, dtype=F32)
    gw2 = torch.empty(D, F, device=dev, dtype=F32)
    gwq = torch.empty(3 * D, D, device=dev, dtype=F32)
    gwp = torch.empty(D, D, device=dev, dtype=F32)
    cases = [
        ("qkv fwd   [M,1408]x[4224,1408]", 2.0 * M * D * 3 * D, lambda: ops.gemm(x, w_qkv, out=out_q, bias=bias_q)),
        ("proj fwd  [M,1408]x[1408,1408] +res", 2.0 * M * D * D, lambda: ops.gemm(x, w_p, out=out_d32, bias=bias_d, residual=res)),
        ("fc1 fwd   [M,1408]x[6144,1408] gelu", 2.0 * M * D * F, lambda: ops.gemm(x, w1, out=out_f, bias=bias_f, act=ACT_GELU, aux_out=pre)),
        ("fc1 fwd plain (no epilogue)", 2.0 * M * D * F, lambda: ops.gemm(x, w1, out=out_f)),
        ("fc1 fwd gelu + store gelu' (tower path)", 2.0 * M * D * F, lambda: ops.gemm(x, w1, out=out_f, bias=bias_f, act=ACT_GELU_SAVE_GRAD, aux_out=pre)),
        ("fc2 dgrad * saved gelu' (tower path)", 2.0 * M * D * F, lambda: ops.gemm(dy, w2, b_mn=True, out=out_f, act=ACT_MUL_AUX, aux_in=pre)),
        ("fc1 fwd bias only", 2.0 * M * D * F, lambda: ops.gemm(x, w1, out=out_f, bias=bias_f)),
        ("fc1 fwd gelu, no pre-activation store", 2.0 * M * D * F, lambda: ops.gemm(x, w1, out=out_f, bias=bias_f, act=ACT_GELU)),
        ("fc2 dgrad plain (no gelu')", 2.0 * M * D * F, lambda: ops.gemm(dy, w2, b_mn=True, out=out_f)),
        ("fc2 fwd   [M,6144]x[1408,6144] +res", 2.0 * M * D * F, lambda: ops.gemm(a, w2, out=out_d32, bias=bias_d, residual=res)),
        ("fc2 dgrad [M,1408]x[1408,6144]mn gelu'", 2.0 * M * D * F, lambda: ops.gemm(dy, w2, b_mn=True, out=out_f, act=ACT_GELU_BWD, aux_in=pre)),
        ("fc1 dgrad [M,6144]x[6144,1408]mn", 2.0 * M * D * F, lambda: ops.gemm(a, w1, b_mn=True, out=out_d)),
        ("qkv dgrad [M,4224]x[4224,1408]mn", 2.0 * M * D * 3 * D, lambda: ops.gemm(dqkv, w_qkv, b_mn=True, out=out_d)),
        ("proj dgrad", 2.0 * M * D * D, lambda: ops.gemm(dy, w_p, b_mn=True, out=out_d)),
        ("fc2 wgrad [1408,6144] K=M", 2.0 * M * D * F, lambda: ops.gemm(dy, a, a_mn=True, b_mn=True, out=gw2)),
        ("fc1 wgrad [6144,1408] K=M", 2.0 * M * D * F, lambda: ops.gemm(a, x, a_mn=True, b_mn=True, out=gw1)),
        ("qkv wgrad [4224,1408] K=M", 2.0 * M * D * 3 * D, lambda: ops.gemm(dqkv, x, a_mn=True, b_mn=True, out=gwq)),
        ("proj wgrad [1408,1408] K=M", 2.0 * M * D * D, lambda: ops.gemm(dy, x, a_mn=True, b_mn=True, out=gwp)),
    ]
    tot = 0.0
    for name, fl, fn in cases:
        t = timeit(fn)
        tot += t
        print(f"{name:44s} {t:8.3f} ms  {fl / t / 1e9:8.1f} TFLOP/s")
    # torch/cuBLAS reference for the fc1 shape (library baseline, not the product)
    t = timeit(lambda: torch.matmul(x, w1.t()))
    print(f"{'cuBLAS fc1 (torch.matmul)':44s} {t:8.3f} ms  {2.0 * M * D * F / t / 1e9:8.1f} TFLOP/s")
    t = timeit(lambda: torch.matmul(dy.t(), a))
    print(f"{'cuBLAS fc2 wgrad (torch.matmul)':44s} {t:8.3f} ms  {2.0 * M * D * F / t / 1e9:8.1f} TFLOP/s")
    # HBM-bound kernels
    xf = torch.randn(M, D, device=dev)
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    t = timeit(lambda: ops.layernorm_fwd(xf, g, b, 1e-6))
    print(f"{'layernorm fwd fp32->bf16':44s} {t:8.3f} ms  {M * D * 6 / t / 1e6:8.1f} GB/s")
    yb, _, mean, rstd = ops.layernorm_fwd(xf, g, b, 1e-6)
    dg, db = torch.empty_like(g), torch.empty_like(b)
    t = timeit(lambda: ops.layernorm_bwd(yb, xf, mean, rstd, g, dg, db, dres=xf, want_bf16=True))
    print(f"{'layernorm bwd (+res, +bf16 copy)':44s} {t:8.3f} ms  {M * D * (2 + 4 + 4 + 4 + 2) / t / 1e6:8.1f} GB/s")
    cs = torch.empty(D, device=dev)
    rs = torch.ones(64, device=dev)
    t = timeit(lambda: ops.layernorm_bwd(yb, xf, mean, rstd, g, dg, db, dres=xf, want_bf16=True, row_scale=rs, rows_per_group=257,
                                         colsum_out=cs))
    print(f"{'layernorm bwd (+res, +bf16 copy, +colsum)':44s} {t:8.3f} ms  {M * D * (2 + 4 + 4 + 4 + 2) / t / 1e6:8.1f} GB/s")
    t = timeit(lambda: ops.colsum(a))
    print(f"{'colsum [M,6144] bf16':44s} {t:8.3f} ms  {M * F * 2 / t / 1e6:8.1f} GB/s")
    B, H, S, d = 64, 16, 257, 88
    qkv = r(B, S, 3, H, d)
    do = r(B, S, H, d)
    o, lse = ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], d ** -0.5)
    t = timeit(lambda: ops.attention_fwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], d ** -0.5, out=o))
    print(f"{'attention fwd':44s} {t:8.3f} ms  {4.0 * B * H * S * S * d / t / 1e9:8.1f} TFLOP/s")
    dq = torch.empty_like(qkv)
    t = timeit(lambda: ops.attention_bwd(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], o, lse, do, d ** -0.5, dq=dq[:, :, 0], dk=dq[:, :, 1], dv=dq[:, :, 2]))
    print(f"{'attention bwd':44s} {t:8.3f} ms  {10.0 * B * H * S * S * d / t / 1e9:8.1f} TFLOP/s")


if __name__ == "__main__":
    main()
