import os, sys
sys.path.insert(0, os.getcwd())
import torch
from mico_b200 import ops
r = lambda *s: (torch.randn(*s, device="cuda") * 0.5).to(torch.bfloat16)
B, H, S, d = 64, 16, 257, 88
qkv, do = r(B, S, 3, H, d), r(B, S, H, d)
dq = torch.empty_like(qkv)
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
o, lse = ops.attention_fwd(q, k, v, d ** -0.5)
for _ in range(3):
    ops.attention_bwd(q, k, v, o, lse, do, d ** -0.5, dq=dq[:, :, 0], dk=dq[:, :, 1], dv=dq[:, :, 2])
torch.cuda.synchronize()
