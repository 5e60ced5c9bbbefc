"""HBM roofline of the spatial (shift-family) kernels at the BASELINE config-4 shapes, batch 256:
achieved GB/s = algorithmic bytes (passes x tensor bytes) / CUDA-event time, vs the measured copy
bandwidth in MEASURED_PEAKS.json.  Prints one JSON object (kept under profiles/)."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jittor_mlp_b200 as J  # noqa: E402
from jittor_mlp_b200 import _lib as L, fn, fn_s2, fn_spatial, ops  # noqa: E402

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
DEV = "cuda"


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


res = {}


def rec(name, secs, nbytes, note):
    res[name] = {"us": round(secs * 1e6, 1), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / secs / 1e9, 0),
                 "frac_of_measured_hbm": round(nbytes / secs / 1e9 / PEAK, 3), "note": note}


B = 256
with torch.no_grad():
    # AS-MLP-T stage 0: [256, 56, 56, 96]
    x = bf(B, 56, 56, 96); n = x.numel() * 2
    rec("as_mlp axial_shift W (zero pad), stage 0", timeit(lambda: fn.axial_shift(x, 5, 3)), 2 * n, "1R + 1W; reference: CuPy kernel, same bytes, scalar")
    rec("as_mlp axial_shift H (zero pad), stage 0", timeit(lambda: fn.axial_shift(x, 5, 2)), 2 * n, "1R + 1W")
    w, b = bf(96), bf(96)
    rec("as_mlp GroupNorm(1,C)+GELU fwd, stage 0", timeit(lambda: fn.group_norm1(x, w, b, 1e-5, True)), 3 * n, "stats 1R, apply 1R + 1W")
    # S2-MLPv2 stage 0: [256, 32, 32, 192], t = 3C
    C = 192
    t = bf(B, 32, 32, 3 * C); nt = t.numel() * 2
    hat = bf(B, 3 * C)
    rec("s2mlpv2 split-attention sum (shifted reads), stage 0", timeit(lambda: fn_s2.S2v2SumFn.apply(t)), nt, "1R of t[3C]; no stack / shifted copies")
    rec("s2mlpv2 split-attention combine (shifted reads), stage 0", timeit(lambda: fn_s2.S2v2CombineFn.apply(t, hat)), nt + nt // 3, "1R t[3C] + 1W [C]")
    xs = bf(B, 32, 32, C)
    rec("s2mlp spatial_shift (clamp), stage 0", timeit(lambda: fn.s2_shift(xs, 1)), 2 * xs.numel() * 2, "1R + 1W (S2-MLP v1 path)")
    # Hire-MLP-T stage 0: [256, 56, 56, 64], h = w = 4, step 2
    xh = bf(B, 56, 56, 64); nh = xh.numel() * 2
    zh, zw = fn_spatial.HireBuildFn.apply(xh, 4, 4, 2, 2)
    rec("hire region build (pad+roll+gather, both axes), stage 0", timeit(lambda: fn_spatial.HireBuildFn.apply(xh, 4, 4, 2, 2)),
        2 * nh + (zh.numel() + zw.numel()) * 2, "2R x + 1W zh + 1W zw (60/56 padding overhead included)")
    rec("hire combine (restore+roll back+crop+sum), stage 0", timeit(lambda: fn_spatial.HireCombineFn.apply(xh, zh, zw, 4, 4, 2, 2)),
        2 * nh + (zh.numel() + zw.numel()) * 2, "R base + R oh + R ow + W out")
    # ConvMixer-768 k7: [256, 32, 32, 768]
    xc = bf(B, 32, 32, 768); nc = xc.numel() * 2
    wd, bd = bf(768, 1, 7, 7) * 0.1, bf(768)
    tdw = timeit(lambda: fn_spatial.DwConvGeluFn.apply(xc, wd, bd), iters=5)
    rec("convmixer depthwise 7x7 + bias + GELU fwd", tdw, 3 * nc, f"1R + 2W; FP32-FMA bound: {2 * xc.numel() * 49 / tdw / 1e12:.1f} TFLOP/s")
    # row-wise
    xm = bf(B * 196, 768); nm = xm.numel() * 2
    g, be = bf(768), bf(768)
    rec("layernorm fwd [50176, 768]", timeit(lambda: ops.layernorm_fwd(xm, g, be)), 2 * nm, "1R + 1W")
print(json.dumps({"hbm_peak_GBps_measured": PEAK, "batch": B, "kernels": res}, indent=1))
