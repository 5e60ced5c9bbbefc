"""Row-wise kernels (LayerNorm fwd/bwd, reductions) against torch fp32 on the same bf16 inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from jittor_mlp_b200 import _lib as L, ops  # noqa: E402

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).bfloat16()


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("rows,C", [(1000, 768), (77, 64), (513, 1024), (64, 1536), (300, 72), (31, 3072)])
def test_layernorm_fwd_bwd(rows, C):
    x = rnd(rows, C, seed=1) * 2 + 0.5
    g, b = rnd(C, seed=2) * 0.2 + 1, rnd(C, seed=3) * 0.2
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    xf = x.float().requires_grad_(True)
    gf, bf = g.float().requires_grad_(True), b.float().requires_grad_(True)
    ref = F.layer_norm(xf, (C,), gf, bf, 1e-5)
    assert rel(y, ref) < 4e-3
    assert rel(mean, xf.mean(-1)) < 1e-5
    dy, add = rnd(rows, C, seed=4), rnd(rows, C, seed=5)
    ref.backward(dy.float())
    dx, dg, db = ops.layernorm_bwd(dy, x, mean, rstd, g, add)
    assert rel(dx, xf.grad + add.float()) < 4e-3
    assert rel(dg, gf.grad) < 2e-3
    assert rel(db, bf.grad) < 2e-3


def test_colsum_and_rowsum_and_cast():
    lib = L.lib()
    a = rnd(3000, 264, seed=1)
    out = torch.zeros(264, dtype=torch.float32, device=DEV)
    L.check(lib.vmlp_colsum(a.data_ptr(), 264, 0, 0, out.data_ptr(), 3000, 264, L.stream_ptr()))
    assert rel(out, a.float().sum(0)) < 1e-4
    t = rnd(7, 50, 128, seed=2)
    out2 = torch.zeros(50, dtype=torch.float32, device=DEV)
    L.check(lib.vmlp_rowsum_batched(t.data_ptr(), out2.data_ptr(), 7, 50, 128, L.stream_ptr()))
    assert rel(out2, t.float().sum((0, 2))) < 1e-4
    f = torch.randn(1001, device=DEV)
    assert torch.equal(ops.cast_f32_to_bf16(f), f.bfloat16())
