"""clock64 timeline of CTA 0 of the fused forward token kernel (bring-up): prints per global chunk the cycle stamps of
the MMA warp and of the first epilogue warp relative to the first stamp.
MMA events: 0 iteration start, 1 after xt_full, 2 after z_empty, 3 after wa_full, 4 G1 issued + committed,
            5 after h_full, 6 after wb_full/u_empty, 7 G2 issued + committed
EPI events: 0 chunk start, 1 after z_full, 2 Z loaded + z_empty arrive, 3 math done, 4 after h_empty, 5 after hs_empty,
            6 tile written + fence, 7 h_full arrive"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jittor_mlp_b200 as J  # noqa: E402,F401
from jittor_mlp_b200 import _lib as L, ops  # noqa: E402

B, N, C, Ds = 256, 196, 768, 784
bf = lambda *s: torch.randn(*s, device="cuda", dtype=torch.bfloat16) * 0.05
xhat, x = bf(B, N, C), bf(B, N, C)
w1, w2, b1, b2 = bf(Ds, N), bf(N, Ds), bf(Ds), bf(N)
for _ in range(3):
    ops.tokmix_fwd(xhat, x, w1, b1, w2, b2)
trace = torch.zeros(4, 64, 8, dtype=torch.int64, device="cuda")
L.lib().vmlp_tokmix_set_trace(trace.data_ptr())
ops.tokmix_fwd(xhat, x, w1, b1, w2, b2)
torch.cuda.synchronize()
L.lib().vmlp_tokmix_set_trace(0)
t = trace.cpu()
t0 = int(t[t > 0].min())
for role, name in ((0, "MMA"), (1, "EPI")):
    print(name)
    for g in range(40):
        row = [int(v) - t0 if v > 0 else -1 for v in t[role, g]]
        print(f"  g={g:2d} " + " ".join(f"{v:7d}" for v in row))
