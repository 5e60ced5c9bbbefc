"""GPU bring-up probe: run each GEMM configuration in its own process (a device trap poisons the context) and
print one line per case with the relative error and a few sample values.  Not a test -- diagnostics for gpurun."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # name, M, N, K, a_major, b_major, batch
    ("kk_1tile", 128, 256, 64, 0, 0, 0),
    ("kk_k256", 128, 256, 256, 0, 0, 0),
    ("kk_multi", 512, 512, 512, 0, 0, 0),
    ("k_mn", 256, 512, 256, 0, 1, 0),
    ("mn_k", 256, 512, 256, 1, 0, 0),
    ("mn_mn", 256, 512, 256, 1, 1, 0),
    ("ragged", 200, 264, 200, 0, 0, 0),
    ("ragged_mn", 200, 264, 200, 1, 1, 0),
    ("bn128", 256, 128, 128, 0, 0, 0),
]


def run_case(name):
    import torch
    from jittor_mlp_b200 import _lib as L, ops
    c = {x[0]: x for x in CASES}[name]
    _, M, N, K, am, bm, _ = c
    g = torch.Generator().manual_seed(0)
    A = torch.randn((M, K) if am == 0 else (K, M), generator=g).cuda().bfloat16()
    B = torch.randn((N, K) if bm == 0 else (K, N), generator=g).cuda().bfloat16()
    D = torch.full((M, N), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.gemm(M, N, K, ops.operand(A, am), ops.operand(B, bm), L.EPI_STORE, D=D)
    torch.cuda.synchronize()
    Af = A.float() if am == 0 else A.float().t()
    Bf = B.float() if bm == 0 else B.float().t()
    ref = Af @ Bf.t()
    err = float((D.float() - ref).norm() / ref.norm())
    # per-quadrant error map helps to localise descriptor / swizzle mistakes
    qm, qn = max(M // 4, 1), max(N // 4, 1)
    emap = [[round(float((D.float()[i * qm:(i + 1) * qm, j * qn:(j + 1) * qn] - ref[i * qm:(i + 1) * qm, j * qn:(j + 1) * qn]).norm()
                         / ref[i * qm:(i + 1) * qm, j * qn:(j + 1) * qn].norm()), 3) for j in range(4)] for i in range(4)]
    print(json.dumps(dict(case=name, rel_err=err, d00=D[0, :4].float().tolist(), r00=ref[0, :4].tolist(), errmap=emap)))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
        sys.exit(0)
    for c in CASES:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), c[0]], capture_output=True, text=True, timeout=180)
            out = (r.stdout.strip().splitlines() or ["<no stdout>"])[-1]
            print(out if r.returncode == 0 else f"{c[0]}: rc={r.returncode} {out} :: {r.stderr.strip()[-600:]}")
        except subprocess.TimeoutExpired:
            print(f"{c[0]}: TIMEOUT")
        sys.stdout.flush()
