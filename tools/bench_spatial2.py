"""Backward / reduction kernels of the shift family at the BASELINE config-4 stage-0 shapes (batch 256): CUDA-event
time vs the HBM floor (algorithmic bytes / measured copy bandwidth).  One line per kernel group."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jittor_mlp_b200 import fn, fn_s2, fn_spatial, ops  # noqa: E402

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
DEV = "cuda"
B = 256


def timeit(f, iters=20, warm=3):
    for _ in range(warm):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / iters


def bf(*s):
    return torch.randn(*s, device=DEV, dtype=torch.bfloat16)


def rec(name, secs, nbytes):
    print(f"{name:58s} {secs * 1e6:8.1f} us  {nbytes / 1e6:7.1f} MB  {nbytes / secs / 1e9:6.0f} GB/s  {nbytes / secs / 1e9 / PEAK:5.2f} of HBM", flush=True)


def bwd_time(make_out, x):
    y = make_out()
    dy = torch.randn_like(y)
    return timeit(lambda: torch.autograd.grad(y, x, dy, retain_graph=True))


# AS-MLP-T stage 0
x = bf(B, 56, 56, 96).requires_grad_(True); n = x.numel() * 2
w, b = bf(96).requires_grad_(True), bf(96).requires_grad_(True)
rec("as_mlp axial_shift fwd [256,56,56,96]", timeit(lambda: fn.axial_shift(x.detach(), 5, 3)), 2 * n)
rec("as_mlp axial_shift bwd", bwd_time(lambda: fn.axial_shift(x, 5, 3), x), 2 * n)
rec("as_mlp GroupNorm(1,C) fwd (stats + apply)", timeit(lambda: fn.group_norm1(x.detach(), w.detach(), b.detach(), 1e-5, False)), 3 * n)
rec("as_mlp GroupNorm(1,C) bwd (reduce + apply)", bwd_time(lambda: fn.group_norm1(x, w, b, 1e-5, False), x), 6 * n)
rec("as_mlp GroupNorm(1,C)+GELU bwd", bwd_time(lambda: fn.group_norm1(x, w, b, 1e-5, True), x), 6 * n)
a2 = x.detach().view(-1, 96)
out = torch.zeros(96, device=DEV)
rec("colsum [802816, 96]", timeit(lambda: ops.colsum_into(out, a2)), n)
a3 = bf(B * 56 * 56, 384)
out3 = torch.zeros(384, device=DEV)
rec("colsum [802816, 384]", timeit(lambda: ops.colsum_into(out3, a3)), a3.numel() * 2)
# S2-MLPv2 stage 0
C = 192
t = bf(B, 32, 32, 3 * C).requires_grad_(True); nt = t.numel() * 2
hat = bf(B, 3 * C).requires_grad_(True)
rec("s2v2 sum fwd", timeit(lambda: fn_s2.S2v2SumFn.apply(t.detach())), nt)
rec("s2v2 sum bwd (dt write)", bwd_time(lambda: fn_s2.S2v2SumFn.apply(t), t), nt)
rec("s2v2 combine fwd", timeit(lambda: fn_s2.S2v2CombineFn.apply(t.detach(), hat.detach())), nt + nt // 3)
rec("s2v2 combine bwd (reduce + dt)", bwd_time(lambda: fn_s2.S2v2CombineFn.apply(t, hat), t), nt + nt // 3 + nt // 3 + nt)
# Hire-MLP-T stage 0 LayerNorm [802816, 64], stage 1 [200704, 128]
for rows, Cc in ((B * 56 * 56, 64), (B * 28 * 28, 128), (B * 14 * 14, 320), (B * 196, 768)):
    xm = bf(rows, Cc); nm = xm.numel() * 2
    g, be = bf(Cc), bf(Cc)
    rec(f"layernorm fwd [{rows}, {Cc}]", timeit(lambda: ops.layernorm_fwd(xm, g, be)), 2 * nm)
    y, mean, rstd = ops.layernorm_fwd(xm, g, be)
    dy = bf(rows, Cc)
    rec(f"layernorm bwd [{rows}, {Cc}] (+add)", timeit(lambda: ops.layernorm_bwd(dy, xm, mean, rstd, g, add=dy)), 4 * nm)
# ConvMixer-768/32 (k = 7): depthwise stencil fwd (+bias+GELU, 2 outputs), dgrad + wgrad
xc = bf(B, 32, 32, 768).requires_grad_(True); nc = xc.numel() * 2
wd = (bf(768, 1, 7, 7) * 0.1).requires_grad_(True)
bd = bf(768).requires_grad_(True)
tdw = timeit(lambda: fn_spatial.DwConvGeluFn.apply(xc.detach(), wd.detach(), bd.detach()), iters=10)
rec(f"convmixer dwconv 7x7+bias+GELU fwd ({2 * xc.numel() * 49 / tdw / 1e12:.1f} TFLOP/s fp32)", tdw, 3 * nc)
yc = fn_spatial.DwConvGeluFn.apply(xc, wd, bd)
dyc = torch.randn_like(yc)
tbw = timeit(lambda: torch.autograd.grad(yc, (xc, wd, bd), dyc, retain_graph=True), iters=10)
rec(f"convmixer dwconv bwd (dgelu + dgrad + wgrad + dbias) ({4 * xc.numel() * 49 / tbw / 1e12:.1f} TFLOP/s fp32)", tbw, 6 * nc)
