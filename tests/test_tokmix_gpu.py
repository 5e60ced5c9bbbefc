"""Fused token-mixing MLP kernels (vmlp_tokmix_fwd / _bwd, csrc/tokmix_sm100.cuh) against a torch fp32 evaluation of
models_pytorch/mlp_mixer.py:16-27,37 (FeedForward with Conv1d(k=1) over tokens + residual) on the same bf16 inputs.
Shapes cover Mixer-B/16 and L/16, an odd tile count (one dead CTA of the last pair), every token-axis tail form
(NT % 64 = 0 / 16 / 32 / 48), a ragged channel count and a hidden width whose last chunk is 16 wide.
Tolerance: bf16 storage of the hidden activation => rel-L2 <= 5e-3 on outputs (1e-2 is the north-star bound)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import jittor_mlp_b200 as J  # noqa: E402,F401
from jittor_mlp_b200 import ops  # noqa: E402

DEV = "cuda"
SHAPES = [  # B, N, C, Ds
    (4, 196, 768, 784),      # Mixer-B/16
    (2, 196, 1024, 784),     # Mixer-L/16
    (3, 64, 128, 256),       # smoke-test shape, 3 tiles -> dead CTA
    (2, 16, 64, 64),         # one k-step, one chunk
    (2, 49, 200, 200),       # NT 64 > N, ragged channels, last chunk 8 -> 16 wide
    (5, 80, 256, 136),       # NT 80: one full atom + SWIZZLE_32B tail
    (2, 100, 128, 320),      # NT 112: tail 48 (full-width tail atom, 3 k-steps)
    (2, 20, 128, 128),       # NT 32: tail only
    (1, 240, 384, 1024),     # largest supported token count / hidden width (forward only)
]


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


def make(B, N, C, Ds):
    xhat, x = rnd(B, N, C, seed=1), rnd(B, N, C, seed=2)
    w1, w2 = rnd(Ds, N, scale=N ** -0.5, seed=3), rnd(N, Ds, scale=Ds ** -0.5, seed=4)
    b1, b2 = rnd(Ds, scale=0.5, seed=5), rnd(N, scale=0.5, seed=6)
    return xhat, x, w1, b1, w2, b2


@pytest.mark.parametrize("B,N,C,Ds", SHAPES)
def test_forward(B, N, C, Ds):
    assert ops.tokmix_supported(B, N, C, Ds)
    xhat, x, w1, b1, w2, b2 = make(B, N, C, Ds)
    u, hT = ops.tokmix_fwd(xhat, x, w1, b1, w2, b2)
    torch.cuda.synchronize()
    z = torch.einsum("mn,bnc->bmc", w1.float(), xhat.float()) + b1.float()[None, :, None]
    h = gelu(z)
    ref = x.float() + torch.einsum("nm,bmc->bnc", w2.float(), h) + b2.float()[None, :, None]
    assert rel(hT.float().transpose(1, 2), h) < 5e-3, rel(hT.float().transpose(1, 2), h)
    assert rel(u, ref) < 5e-3, rel(u, ref)
    u2, none = ops.tokmix_fwd(xhat, x, w1, b1, w2, b2, save_hidden=False)
    assert none is None and torch.equal(u, u2)            # no atomics in forward: bit-reproducible


@pytest.mark.parametrize("B,N,C,Ds", [s for s in SHAPES if ops is not None and s[1] <= 208])
def test_backward(B, N, C, Ds):
    if not ops.tokmix_supported(B, N, C, Ds, backward=True):
        pytest.skip("backward tiles of this shape do not fit in shared memory")
    xhat, _, w1, b1, w2, _ = make(B, N, C, Ds)
    du = rnd(B, N, C, seed=7)
    dxh, dzT, db1 = ops.tokmix_bwd(xhat, du, w1, b1, w2)
    torch.cuda.synchronize()
    z = torch.einsum("mn,bnc->bmc", w1.float(), xhat.float()) + b1.float()[None, :, None]
    dh = torch.einsum("nm,bnc->bmc", w2.float(), du.float())
    dz = dh * dgelu(z)
    ref_dx = torch.einsum("mn,bmc->bnc", w1.float(), dz)
    assert rel(dzT.float().transpose(1, 2), dz) < 5e-3, rel(dzT.float().transpose(1, 2), dz)
    assert rel(dxh, ref_dx) < 5e-3, rel(dxh, ref_dx)
    assert rel(db1, dz.sum(dim=(0, 2))) < 5e-3, rel(db1, dz.sum(dim=(0, 2)))


def test_bench_shape_b256_slice():
    """The benchmarked launch (B = 256: 1536 tiles over 74 CTA pairs, ~10 items per pair): every image against the
    fp32 reference evaluated on the GPU."""
    B, N, C, Ds = 256, 196, 768, 784
    xhat, x, w1, b1, w2, b2 = make(B, N, C, Ds)
    u, hT = ops.tokmix_fwd(xhat, x, w1, b1, w2, b2)
    du = rnd(B, N, C, seed=7)
    dxh, dzT, db1 = ops.tokmix_bwd(xhat, du, w1, b1, w2)
    torch.cuda.synchronize()
    worst = 0.0
    dbsum = torch.zeros(Ds, device=DEV, dtype=torch.float64)
    for b0 in range(0, B, 32):
        sl = slice(b0, b0 + 32)
        z = torch.einsum("mn,bnc->bmc", w1.float(), xhat[sl].float()) + b1.float()[None, :, None]
        ref = x[sl].float() + torch.einsum("nm,bmc->bnc", w2.float(), gelu(z)) + b2.float()[None, :, None]
        dz = torch.einsum("nm,bnc->bmc", w2.float(), du[sl].float()) * dgelu(z)
        ref_dx = torch.einsum("mn,bmc->bnc", w1.float(), dz)
        dbsum += dz.double().sum(dim=(0, 2))
        worst = max(worst, rel(u[sl], ref), rel(dxh[sl], ref_dx), rel(dzT[sl].float().transpose(1, 2), dz),
                    rel(hT[sl].float().transpose(1, 2), gelu(z)))
    assert worst < 5e-3, worst
    assert rel(db1, dbsum) < 5e-3


def test_unsupported_shapes_are_reported():
    assert not ops.tokmix_supported(2, 257, 128, 256)         # token axis beyond one accumulator
    assert not ops.tokmix_supported(2, 196, 132, 256)         # channel pitch not 16-byte aligned
    assert not ops.tokmix_supported(2, 240, 128, 256, backward=True)   # two resident activation tiles + weight rings do not fit
    x = rnd(2, 257, 128)
    with pytest.raises(ValueError):
        ops.tokmix_fwd(x, x, rnd(256, 257), rnd(256), rnd(257, 256), rnd(257))
