"""The fused token-mixing kernels against fp32 torch, every (shape, direction) in its own forked process (a device-side
trap kills the CUDA context; forking before CUDA is initialised keeps the import cost out of each case):
    python tools/tokmix_check.py B N C Ds [fwd|bwd|both]      one case
    python tools/tokmix_check.py --all                         the bring-up list"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jittor_mlp_b200 as J  # noqa: E402,F401
from jittor_mlp_b200 import ops  # noqa: E402

ALL = [(2, 16, 64, 64), (2, 64, 128, 256), (3, 64, 128, 256), (2, 49, 200, 200), (5, 80, 256, 136), (2, 100, 128, 320),
       (2, 20, 128, 128), (2, 192, 128, 128), (2, 208, 128, 128), (2, 196, 128, 64), (2, 196, 128, 784),
       (4, 196, 768, 784), (2, 196, 1024, 784), (1, 240, 384, 1024), (256, 196, 768, 784)]
if sys.argv[1] == "--all":
    import signal
    import time
    bad = 0
    for shape in ALL:
        for what in ("fwd", "bwd"):
            pid = os.fork()
            if pid == 0:
                sys.argv = [sys.argv[0]] + [str(v) for v in shape] + [what]
                break
            t0, status = time.time(), None
            while time.time() - t0 < 120:
                done, st = os.waitpid(pid, os.WNOHANG)
                if done:
                    status = st
                    break
                time.sleep(0.2)
            if status is None:
                os.kill(pid, signal.SIGKILL)
                os.waitpid(pid, 0)
                print("TOKMIX", shape, what, "TIMEOUT", flush=True)
            if status != 0:
                bad += 1
                print("TOKMIX", shape, what, "exit status", status, flush=True)
        else:
            continue
        break
    else:
        print("TOKMIX bring-up:", "all green" if bad == 0 else f"{bad} cases failed", flush=True)
        sys.exit(1 if bad else 0)

B, N, C, Ds = map(int, sys.argv[1:5])
what = sys.argv[5] if len(sys.argv) > 5 else "both"
DEV = "cuda"


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).bfloat16()


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gelu(z):
    return 0.5 * z * (1 + torch.erf(z / math.sqrt(2.0)))


def dgelu(z):
    return 0.5 * (1 + torch.erf(z / math.sqrt(2.0))) + z * torch.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)


import atexit
import time
from jittor_mlp_b200 import _lib as L  # noqa: E402
_t0 = time.time()


def _report():
    recs = L.debug_records()
    if recs:
        print("TOKMIX", (B, N, C, Ds), what, f"{len(recs)} mbarrier timeouts after {time.time() - _t0:.1f}s; (block, warp, bar offset, parity):",
              sorted({(b, w, a % 1024 if a % 1024 < 512 else a, p) for b, w, a, p in recs})[:48], flush=True)


atexit.register(_report)
xhat, x = rnd(B, N, C, seed=1), rnd(B, N, C, seed=2)
w1, w2 = rnd(Ds, N, scale=N ** -0.5, seed=3), rnd(N, Ds, scale=Ds ** -0.5, seed=4)
b1, b2 = rnd(Ds, scale=0.5, seed=5), rnd(N, scale=0.5, seed=6)
du = rnd(B, N, C, seed=7)
z = torch.einsum("mn,bnc->bmc", w1.float(), xhat.float()) + b1.float()[None, :, None]
res = {}
if what in ("fwd", "both"):
    u, hT = ops.tokmix_fwd(xhat, x, w1, b1, w2, b2)
    torch.cuda.synchronize()
    h = gelu(z)
    ref = x.float() + torch.einsum("nm,bmc->bnc", w2.float(), h) + b2.float()[None, :, None]
    res["hT"] = rel(hT.float().transpose(1, 2), h)
    res["u"] = rel(u, ref)
    if res["hT"] > 5e-3:       # where is it wrong?  per 64-wide hidden chunk and per 32-channel quarter
        e = (hT.float().transpose(1, 2) - h)
        res["hT_err_by_chunk"] = [round(float(e[:, m:m + 64].norm() / h[:, m:m + 64].norm()), 4) for m in range(0, Ds, 64)]
        res["hT_err_by_cquarter"] = [round(float(e[:, :, c:c + 32].norm() / h[:, :, c:c + 32].norm()), 4) for c in range(0, min(C, 256), 32)]
    if res["u"] > 5e-3:
        e = (u.float() - ref)
        res["u_err_by_tok16"] = [round(float(e[:, n:n + 16].norm() / ref[:, n:n + 16].norm()), 4) for n in range(0, N, 16)]
        res["u_err_by_image"] = [round(float(e[b].norm() / ref[b].norm()), 4) for b in range(min(B, 8))]
if what in ("bwd", "both") and ops.tokmix_supported(B, N, C, Ds, backward=True):
    dxh, dzT, db1 = ops.tokmix_bwd(xhat, du, w1, b1, w2)
    torch.cuda.synchronize()
    dz = torch.einsum("nm,bnc->bmc", w2.float(), du.float()) * dgelu(z)
    ref_dx = torch.einsum("mn,bmc->bnc", w1.float(), dz)
    res["dzT"] = rel(dzT.float().transpose(1, 2), dz)
    res["dxh"] = rel(dxh, ref_dx)
    res["db1"] = rel(db1, dz.sum(dim=(0, 2)))
    if res["dzT"] > 5e-3:
        e = (dzT.float().transpose(1, 2) - dz)
        res["dzT_err_by_chunk"] = [round(float(e[:, m:m + 64].norm() / dz[:, m:m + 64].norm()), 4) for m in range(0, Ds, 64)]
print("TOKMIX", (B, N, C, Ds), what, f"{time.time() - _t0:.1f}s", {k: (round(v, 5) if isinstance(v, float) else v) for k, v in res.items()}, flush=True)
ok = all(v < 5e-3 for k, v in res.items() if isinstance(v, float))
sys.exit(0 if ok else 1)
