"""Repeated launches of the fused forward / backward token kernels at the bench shape; on a failure prints the records the
bounded mbarrier waits left behind (which block / warp / barrier / parity timed out).
python tools/tokmix_stress.py [fwd|bwd] [iters] [save 0/1]"""
import atexit
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jittor_mlp_b200 as J  # noqa: E402,F401
from jittor_mlp_b200 import _lib as L, ops  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "fwd"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
save = int(sys.argv[3]) if len(sys.argv) > 3 else 1
B, N, C, Ds = 256, 196, 768, 784
bf = lambda *s: torch.randn(*s, device="cuda", dtype=torch.bfloat16) * 0.05
xhat, x = bf(B, N, C), bf(B, N, C)
w1, w2, b1, b2 = bf(Ds, N), bf(N, Ds), bf(Ds), bf(N)
done = [0]


def report():
    recs = L.debug_records()
    print(f"{what} save={save}: {done[0]} launches completed; {len(recs)} timeout records (block, warp, bar byte offset, parity):",
          sorted({(b, w, a % 1024, p) for b, w, a, p in recs})[:64], flush=True)


atexit.register(report)
for i in range(iters):
    if what == "fwd":
        ops.tokmix_fwd(xhat, x, w1, b1, w2, b2, save_hidden=bool(save))
    else:
        ops.tokmix_bwd(xhat, x, w1, b1, w2)
    if i % 4 == 3:
        torch.cuda.synchronize()
        done[0] = i + 1
torch.cuda.synchronize()
done[0] = iters
