"""Same-box A/B of the tower weight-gradient GEMMs at the omni step's 197 376-token pass (qkv, proj, fc1, fc2) under
environment variants (MICO_AB_ENVS="NAME=VALUE:NAME=VALUE").  python scripts/ab_wgrad.py [child]"""
import os
import subprocess
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import torch
    from mico_b200 import ops
    M, D, F = 768 * 257, 1408, 6144
    r = lambda *s: (torch.randn(*s, device="cuda") * 0.05).to(torch.bfloat16)
    x, dqkv, dy, a = r(M, D), r(M, 3 * D), r(M, D), r(M, F)
    cases = [("qkv wgrad 4224x1408", dqkv, x), ("proj wgrad 1408x1408", dy, x), ("fc1 wgrad 6144x1408", a, x), ("fc2 wgrad 1408x6144", dy, a)]
    for name, g, act in cases:
        out = torch.empty(g.shape[1], act.shape[1], device="cuda")
        f = lambda: ops.gemm(g, act, a_mn=True, b_mn=True, out=out)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        ref = (g[:4096].float().t() @ act[:4096].float())
        chk = torch.empty_like(out)
        ops.gemm(g[:4096], act[:4096], a_mn=True, b_mn=True, out=chk)
        err = ((chk - ref).norm() / ref.norm()).item()
        print(f"{name}: {us:8.1f} us  {2.0 * M * g.shape[1] * act.shape[1] / us / 1e6:7.1f} TFLOP/s  (4096-row check rel {err:.1e})")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for v in os.environ.get("MICO_AB_ENVS", "A=0").split(":"):
            out = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **dict([v.split("=", 1)])), capture_output=True, text=True)
            print(v)
            print(out.stdout.strip() or out.stderr[-500:])
