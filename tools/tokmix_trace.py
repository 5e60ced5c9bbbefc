"""clock64 timeline of CTA 0 of the fused forward token kernel: per global chunk g the cycle stamps of the issuer warps,
the polling warp of the epilogue group that owns g, and the forwarding / store warp, relative to the first stamp.
ISSUE events: 0 G1 loop top, 1 after z_empty + wa_full, 2 G1 issued + committed; 4 G2 loop top, 5 after h_full + wb_full,
              6 G2 issued + committed
EPI events:   0 chunk start, 1 after z_full + h_free (+ group barrier), 2 Z loaded, "consumed" signalled, 3 GELU done and
              tile written, 4 "written" signalled
FWD events:   0 "H(g) written" received, 1 h_full forwarded, tile stored, h_free released
`python tools/tokmix_trace.py bwd`: the backward kernel --
ISSUE events: 0 G1+G2 loop top, 1 after zd_empty, 2 both issued + committed; 4 G3 top, 5 after dz_full, 6 G3 issued
EPI events:   0 chunk start, 1 after zd_full, 2 Z / dH loaded, 3 gelu' done, 4 after dz_empty + dzs_empty, 5 tile written,
              dz_full / dz_done signalled
FWD events:   (helper warp 2) 0 dz_done received, 1 column sums done + store issued, 2 buffer released"""
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
bwd = len(sys.argv) > 1 and sys.argv[1] == "bwd"
run = (lambda: ops.tokmix_bwd(xhat, x, w1, b1, w2)) if bwd else (lambda: ops.tokmix_fwd(xhat, x, w1, b1, w2, b2))
for _ in range(3):
    run()
trace = torch.zeros(4, 64, 8, dtype=torch.int64, device="cuda")
L.lib().vmlp_tokmix_set_trace(trace.data_ptr())
run()
torch.cuda.synchronize()
L.lib().vmlp_tokmix_set_trace(0)
t = trace.cpu()
t0 = int(t[t > 0].min())
for role, name in ((0, "ISSUE"), (1, "EPI"), (2, "FWD")):
    print(name)
    for g in range(40):
        row = [int(v) - t0 if v > 0 else -1 for v in t[role, g]]
        print(f"  g={g:2d} " + " ".join(f"{v:7d}" for v in row))
