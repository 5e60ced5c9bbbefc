"""Token-mixing GEMM shapes (Mixer-B/16) under cta_group 1 vs 2: correctness against torch + CUDA-event timing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jittor_mlp_b200 import _lib as L, ops  # noqa: E402

N, C, Ds = 196, 768, 784
Np = (N + 7) // 8 * 8
dev = "cuda"


def bf(*s):
    return (torch.randn(*s, device=dev) * 0.05).bfloat16()


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def run(B, cg, check):
    torch.manual_seed(0)
    Xt, Ht, Zt = bf(B, N, C), bf(B, Ds, C), bf(B, Ds, C)
    W1p, W2t, b1t, b2t = bf(Ds, Np), bf(N, Ds), bf(Ds), bf(N)
    W1p[:, N:] = 0
    out, outZ, outH = torch.empty_like(Xt), torch.empty_like(Ht), torch.empty_like(Ht)
    gWt = torch.zeros(N, Ds, device=dev, dtype=torch.float32)
    w1k = L.Operand(W1p.data_ptr(), Ds, N, Np, 0, 0)
    w1mn = L.Operand(W1p.data_ptr(), Ds, N, Np, 0, 1)
    cases = {
        "fc1_gelu": lambda: ops.gemm(Ds, C, N, w1k, ops.operand(Xt, 1), L.EPI_GELU, batch=B, D=outZ, D2=outH, bias=b1t, bias_mode=2, cta_group=cg),
        "fc2_resid": lambda: ops.gemm(N, C, Ds, ops.operand(W2t, 0), ops.operand(Ht, 1), L.EPI_RESID, batch=B, D=out, bias=b2t, bias_mode=2, aux=Xt, cta_group=cg),
        "dgrad2_dgelu": lambda: ops.gemm(Ds, C, N, ops.operand(W2t, 1), ops.operand(Xt, 1), L.EPI_DGELU, batch=B, D=outH, aux=Zt, cta_group=cg),
        "dgrad1": lambda: ops.gemm(N, C, Ds, w1mn, ops.operand(Ht, 1), L.EPI_STORE, batch=B, D=out, cta_group=cg),
        "wgrad": lambda: ops.gemm(N, Ds, C, ops.operand(Xt, 0), ops.operand(Ht, 0), L.EPI_ATOMIC, batch=B, contract_batch=True, out_f32=gWt, cta_group=cg),
    }
    for k, fn in cases.items():
        line = f"B={B} cg={cg} {k:13s}"
        if check:
            gWt.zero_()
            fn()
            torch.cuda.synchronize()
            W1 = W1p[:, :N].float()
            if k == "fc1_gelu":
                z = torch.einsum("mn,bnc->bmc", W1, Xt.float()) + b1t.float()[None, :, None]
                line += f" err(H)={rel(outH, torch.nn.functional.gelu(z)):.2e}"
            elif k == "fc2_resid":
                r = torch.einsum("nm,bmc->bnc", W2t.float(), Ht.float()) + b2t.float()[None, :, None] + Xt.float()
                line += f" err={rel(out, r):.2e}"
            elif k == "dgrad2_dgelu":
                r = torch.einsum("nm,bnc->bmc", W2t.float(), Xt.float()) * Zt.float()
                line += f" err={rel(outH, r):.2e}"
            elif k == "dgrad1":
                r = torch.einsum("mn,bmc->bnc", W1, Ht.float())
                line += f" err={rel(out, r):.2e}"
            else:
                r = torch.einsum("bnc,bmc->nm", Xt.float(), Ht.float())
                line += f" err={rel(gWt, r):.2e}"
        else:
            line += f" {timeit(fn) * 1e3:8.1f} us"
        print(line, flush=True)


if __name__ == "__main__":
    for cg in (1, 2):
        run(3, cg, True)
    for cg in (1, 2):
        run(256, cg, False)
