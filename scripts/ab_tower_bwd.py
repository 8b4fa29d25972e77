"""Same-box A/B of the tower attention backward (64 x 16 x 257 x 88): interleaved rounds, each variant in its own process.
Variants: the library builds named in MICO_AB_LIBS (colon-separated), or environment settings in MICO_AB_ENVS
("NAME=VALUE:NAME=VALUE", e.g. MICO_ATTN_TAIL_MERGED=1:MICO_ATTN_TAIL_MERGED=0).  python scripts/ab_tower_bwd.py [child]"""
import os
import subprocess
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child():
    import torch
    from mico_b200 import ops
    r = lambda *s: (torch.randn(*s, device="cuda") * 0.5).to(torch.bfloat16)
    B, H, S, d = 64, 16, 257, 88
    qkv, do = r(B, S, 3, H, d), r(B, S, H, d)
    dq = torch.empty_like(qkv)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o, lse = ops.attention_fwd(q, k, v, d ** -0.5)
    f = lambda: ops.attention_bwd(q, k, v, o, lse, do, d ** -0.5, dq=dq[:, :, 0], dk=dq[:, :, 1], dv=dq[:, :, 2])
    for _ in range(20):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        f()
    e1.record()
    torch.cuda.synchronize()
    print(f"{e0.elapsed_time(e1) / 200 * 1e3:.1f}")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child()
    else:
        if os.environ.get("MICO_AB_ENVS"):
            variants = [(v, dict([v.split("=", 1)])) for v in os.environ["MICO_AB_ENVS"].split(":")]
        else:
            variants = [(os.path.basename(l), dict(MICO_B200_LIB=l)) for l in os.environ["MICO_AB_LIBS"].split(":")]
        res = {n: [] for n, _ in variants}
        for _ in range(3):
            for n, env in variants:
                out = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
                res[n].append(out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-200:])
        for n, _ in variants:
            print(n, "us per call:", res[n])
