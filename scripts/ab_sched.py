"""Same-box A/B of the tower's GEMM shapes at the omni step's 197 376-token pass (and the bs-64 ViT-g pass) with the balanced
work-unit schedule on / off (MICO_GEMM_SCHED=1/0; gemm.cu:GemmPlan).  Interleaved children so that both arms see the same
box temperature.  python scripts/ab_sched.py [child]"""
import os
import subprocess
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import torch
    from mico_b200 import ops
    D, F = 1408, 6144
    r = lambda *s: (torch.randn(*s, device="cuda") * 0.05).to(torch.bfloat16)
    for frames in (768, 64):
        M = frames * 257
        x, dqkv, dy, a = r(M, D), r(M, 3 * D), r(M, D), r(M, F)
        wqkv, wproj, w1, w2 = r(3 * D, D), r(D, D), r(F, D), r(D, F)
        res = torch.randn(M, D, device="cuda")
        o32 = torch.empty(M, D, device="cuda")
        o16 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
        bias = torch.zeros(D, device="cuda")
        cases = [
            ("fc2 fwd  +res  N1408 K6144", lambda: ops.gemm(a, w2, bias=bias, residual=res, out=o32), 2.0 * M * D * F),
            ("proj fwd +res  N1408 K1408", lambda: ops.gemm(x, wproj, bias=bias, residual=res, out=o32), 2.0 * M * D * D),
            ("fc1 dgrad      N1408 K6144", lambda: ops.gemm(a, w1, b_mn=True, out=o16), 2.0 * M * D * F),
            ("qkv dgrad      N1408 K4224", lambda: ops.gemm(dqkv, wqkv, b_mn=True, out=o16), 2.0 * M * D * 3 * D),
            ("proj dgrad     N1408 K1408", lambda: ops.gemm(dy, wproj, b_mn=True, out=o16), 2.0 * M * D * D),
            ("qkv wgrad 4224x1408", None, 2.0 * M * D * 3 * D), ("proj wgrad 1408x1408", None, 2.0 * M * D * D),
            ("fc1 wgrad 6144x1408", None, 2.0 * M * D * F), ("fc2 wgrad 1408x6144", None, 2.0 * M * D * F),
        ]
        wg = {"qkv wgrad 4224x1408": (dqkv, x), "proj wgrad 1408x1408": (dy, x), "fc1 wgrad 6144x1408": (a, x), "fc2 wgrad 1408x6144": (dy, a)}
        for name, f, flops in cases:
            if f is None:
                g, act = wg[name]
                out = torch.empty(g.shape[1], act.shape[1], device="cuda")
                f = (lambda g=g, act=act, out=out: ops.gemm(g, act, a_mn=True, b_mn=True, out=out))
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20 if frames == 768 else 60
            e0.record()
            for _ in range(n):
                f()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1e3
            print(f"frames {frames:4d} {name:28s} {us:8.1f} us  {flops / us / 1e6:7.1f} TFLOP/s", flush=True)
        del x, dqkv, dy, a, res, o32, o16
        torch.cuda.empty_cache()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        for v in ("1", "0", "1"):
            out = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, MICO_GEMM_SCHED=v), capture_output=True, text=True)
            print("MICO_GEMM_SCHED=" + v)
            print(out.stdout.strip() or out.stderr[-800:], flush=True)
