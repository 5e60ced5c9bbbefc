"""Stand-alone timing (CUDA events, back-to-back loop, operands >> L2) of the token-mixing half of one MixerBlock at the
Mixer-B/16 batch-256 shapes: the fused kernels (vmlp_tokmix_fwd / _bwd + the two weight-gradient GEMMs) against the unfused
GEMM sequence they replace.  python tools/bench_tokmix.py [B] [C]   -> one JSON line"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jittor_mlp_b200 as J  # noqa: E402,F401
from jittor_mlp_b200 import _lib as L, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
C = int(sys.argv[2]) if len(sys.argv) > 2 else 768
N, Ds = 196, 784
dev = "cuda"
bf = lambda *s: torch.randn(*s, device=dev, dtype=torch.bfloat16) * 0.05
xhat, x, du = bf(B, N, C), bf(B, N, C), bf(B, N, C)
w1, w2, b1, b2 = bf(Ds, N), bf(N, Ds), bf(Ds), bf(N)
Np = (N + 15) // 16 * 16
w1p, w1T = ops.tokmix_prepare(w1, pad=True, transpose=True)
_, w2T = ops.tokmix_prepare(w2, transpose=True, ldt=Np)
u, hT, dxh, dzT = torch.empty_like(x), bf(B, C, Ds), torch.empty_like(x), bf(B, C, Ds)
db1 = torch.zeros(Ds, device=dev)
gW1, gW2 = torch.zeros(Ds, N, device=dev), torch.zeros(N, Ds, device=dev)
Z, H, dZ = bf(B, Ds, C), bf(B, Ds, C), bf(B, Ds, C)
lib, sp = L.lib(), L.stream_ptr


def t(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / iters * 1e3, 1)      # microseconds


w1p_k = L.Operand(w1p.data_ptr(), Ds, N, Np, 0, 0)
w1p_mn = L.Operand(w1p.data_ptr(), Ds, N, Np, 0, 1)
cases = {
    "fused_fwd": lambda: L.check(lib.vmlp_tokmix_fwd(xhat.data_ptr(), x.data_ptr(), w1p.data_ptr(), Np, w2.data_ptr(), b1.data_ptr(),
                                                    b2.data_ptr(), u.data_ptr(), hT.data_ptr(), B, N, C, Ds, sp())),
    "fused_fwd_nosave": lambda: L.check(lib.vmlp_tokmix_fwd(xhat.data_ptr(), x.data_ptr(), w1p.data_ptr(), Np, w2.data_ptr(), b1.data_ptr(),
                                                           b2.data_ptr(), u.data_ptr(), 0, B, N, C, Ds, sp())),
    "fused_bwd": lambda: L.check(lib.vmlp_tokmix_bwd(xhat.data_ptr(), du.data_ptr(), w1p.data_ptr(), w2T.data_ptr(), Np, w1T.data_ptr(),
                                                    b1.data_ptr(), dxh.data_ptr(), dzT.data_ptr(), db1.data_ptr(), B, N, C, Ds, sp())),
    "wgrad1_new[Ds,N]<-dzT,xhat": lambda: ops.gemm(Ds, N, C, ops.operand(dzT, 1), ops.operand(xhat, 0), L.EPI_ATOMIC, batch=B, contract_batch=True, out_f32=gW1),
    "wgrad2_new[Ds,N]^T<-hT,du": lambda: ops.gemm(Ds, N, C, ops.operand(hT, 1), ops.operand(du, 0), L.EPI_ATOMIC, batch=B, contract_batch=True, out_f32=gW2, out_trans=True),
    "wgrad2_new_bn256": lambda: ops.gemm(Ds, N, C, ops.operand(hT, 1), ops.operand(du, 0), L.EPI_ATOMIC, batch=B, contract_batch=True, out_f32=gW2, out_trans=True, block_n=256),
    "unfused_fc1_gelu": lambda: ops.gemm(Ds, C, N, w1p_k, ops.operand(xhat, 1), L.EPI_GELU, batch=B, D=Z, D2=H, bias=b1, bias_mode=2),
    "unfused_fc2_resid": lambda: ops.gemm(N, C, Ds, ops.operand(w2, 0), ops.operand(H, 1), L.EPI_RESID, batch=B, D=u, bias=b2, bias_mode=2, aux=x),
    "unfused_dgrad2_dgelu": lambda: ops.gemm(Ds, C, N, ops.operand(w2, 1), ops.operand(du, 1), L.EPI_DGELU, batch=B, D=dZ, aux=Z, red_out=db1, red_mode=2),
    "unfused_dgrad1": lambda: ops.gemm(N, C, Ds, w1p_mn, ops.operand(dZ, 1), L.EPI_STORE, batch=B, D=dxh),
    "unfused_wgrad2": lambda: ops.gemm(N, Ds, C, ops.operand(du, 0), ops.operand(H, 0), L.EPI_ATOMIC, batch=B, contract_batch=True, out_f32=gW2),
    "unfused_wgrad1": lambda: ops.gemm(Ds, N, C, ops.operand(dZ, 0), ops.operand(xhat, 0), L.EPI_ATOMIC, batch=B, contract_batch=True, out_f32=gW1),
}
only = os.environ.get("TOKMIX_ONLY")
res = {k: t(f) for k, f in cases.items() if not only or k in only.split(",")}
unit = 2.0 * B * Ds * C * N          # FLOPs of one token GEMM
if not only:
    res["fused_total_us"] = round(res["fused_fwd"] + res["fused_bwd"] + res["wgrad1_new[Ds,N]<-dzT,xhat"] + res["wgrad2_new[Ds,N]^T<-hT,du"], 1)
    res["unfused_total_us"] = round(sum(v for k, v in res.items() if k.startswith("unfused_")), 1)
    res["fused_fwd_tflops_algorithmic"] = round(2 * unit / res["fused_fwd"] / 1e6, 1)
    res["fused_bwd_tflops_algorithmic"] = round(2 * unit / res["fused_bwd"] / 1e6, 1)
print(json.dumps({"B": B, "N": N, "C": C, "Ds": Ds, "us": res}))
