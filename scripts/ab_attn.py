"""Micro-benchmark of the attention entry points at the omni step's shapes (CUDA events, 20 launches after 3 warm-ups).
A/B of two library builds on ONE box:  MICO_B200_LIB=/path/to/other.so python scripts/ab_attn.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mico_b200 import ops, _lib

dev = "cuda"
BF16 = torch.bfloat16
r = lambda *s: (torch.randn(*s, device=dev) * 0.5).to(BF16)


def timeit(f, n=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    print("lib:", _lib.LIB_PATH)
    # (a) ViT-g tower, bs 64
    B, H, S, d = 64, 16, 257, 88
    qkv, do = r(B, S, 3, H, d), r(B, S, H, d)
    dq = torch.empty_like(qkv)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.attention_fwd(q, k, v, d ** -0.5)
    fl = 4.0 * B * H * S * S * d
    t = timeit(lambda: ops.attention_fwd(q, k, v, d ** -0.5))
    print(f"tower 64x16x257x88      fwd {t:8.1f} us  {fl / t / 1e6:7.1f} TFLOP/s")
    t = timeit(lambda: ops.attention_bwd(q, k, v, o, lse, do, d ** -0.5, dq=dq[:, :, 0], dk=dq[:, :, 1], dv=dq[:, :, 2]))
    print(f"tower 64x16x257x88      bwd {t:8.1f} us  {2.5 * fl / t / 1e6:7.1f} TFLOP/s")
    # (b) fusion-encoder cross-attention: 256 sequences x 12 heads x 128 queries, 64 shared K/V entries of 2827 keys, dropout
    Bq, E, H, Sq, Sk, d = 256, 64, 12, 128, 2827, 64
    q, do = r(Bq, Sq, H, d), r(Bq, Sq, H, d)
    kv = r(E, Sk, 2, H, d)
    idx = (torch.arange(Bq, device=dev) % E).to(torch.int32)
    grp = ops.kv_groups(idx, E)
    o, lse = ops.attention_fwd(q, kv[:, :, 0], kv[:, :, 1], d ** -0.5, dropout=(0.1, 11), kv_index=idx)
    fl = 4.0 * Bq * H * Sq * Sk * d
    t = timeit(lambda: ops.attention_fwd(q, kv[:, :, 0], kv[:, :, 1], d ** -0.5, dropout=(0.1, 11), kv_index=idx))
    print(f"cross 256x12x128x2827x64 fwd {t:8.1f} us  {fl / t / 1e6:7.1f} TFLOP/s")
    t = timeit(lambda: ops.attention_bwd(q, kv[:, :, 0], kv[:, :, 1], o, lse, do, d ** -0.5, dropout=(0.1, 11), kv_index=idx, groups=grp))
    print(f"cross 256x12x128x2827x64 bwd {t:8.1f} us  {2.5 * fl / t / 1e6:7.1f} TFLOP/s")
    # (c) fusion-encoder self-attention, 3-D mask + dropout
    qkv = r(Bq, Sq, 3, H, d)
    mask = torch.zeros(Bq, Sq, Sq, device=dev).masked_fill_(torch.triu(torch.ones(Sq, Sq, device=dev), 1).bool(), -10000.0)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.attention_fwd(q, k, v, d ** -0.5, mask=mask, dropout=(0.1, 12))
    fl = 4.0 * Bq * H * Sq * Sq * d
    t = timeit(lambda: ops.attention_fwd(q, k, v, d ** -0.5, mask=mask, dropout=(0.1, 12)))
    print(f"self  256x12x128x128x64  fwd {t:8.1f} us  {fl / t / 1e6:7.1f} TFLOP/s")
    t = timeit(lambda: ops.attention_bwd(q, k, v, o, lse, do, d ** -0.5, mask=mask, dropout=(0.1, 12)))
    print(f"self  256x12x128x128x64  bwd {t:8.1f} us  {2.5 * fl / t / 1e6:7.1f} TFLOP/s")


if __name__ == "__main__":
    main()
