"""vmlp_gemm_bf16 (tcgen05 GEMM through the C ABI) against torch fp32 matmul of the same bf16 inputs:
every operand major-ness, batching mode, ragged edge and epilogue the block code relies on."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

import jittor_mlp_b200 as J  # noqa: E402
from jittor_mlp_b200 import _lib as L, ops  # noqa: E402

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


def make_operand(rows_mn, K, major, batch, seed):
    """Returns (tensor, fp32 matrix [batch?, rows_mn, K])."""
    shape = (rows_mn, K) if major == 0 else (K, rows_mn)
    if batch:
        t = rnd(batch, *shape, seed=seed)
        f = t.float() if major == 0 else t.float().transpose(1, 2)
    else:
        t = rnd(*shape, seed=seed)
        f = t.float() if major == 0 else t.float().t()
    return t, f


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (384, 256, 768), (200, 264, 200), (784, 768, 196)])
@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_plain_store(M, N, K, a_major, b_major):
    if (a_major == 0 and K % 8) or (a_major == 1 and M % 8) or (b_major == 0 and K % 8) or (b_major == 1 and N % 8):
        pytest.skip("row pitch must be a multiple of 16 bytes")
    A, Af = make_operand(M, K, a_major, 0, 1)
    B, Bf = make_operand(N, K, b_major, 0, 2)
    D = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(M, N, K, ops.operand(A, a_major), ops.operand(B, b_major), L.EPI_STORE, D=D)
    torch.cuda.synchronize()
    ref = Af @ Bf.t()
    assert rel(D, ref) < 5e-3, rel(D, ref)


@pytest.mark.parametrize("block_n", [128, 256])
def test_block_n_and_bias_modes(block_n):
    M, N, K = 300, 384, 320
    A, Af = make_operand(M, K, 0, 0, 1)
    B, Bf = make_operand(N, K, 0, 0, 2)
    bc, br = rnd(N, seed=3), rnd(M, seed=4)
    for mode, bias, add in ((1, bc, bc.float()[None, :]), (2, br, br.float()[:, None])):
        D = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        ops.gemm(M, N, K, ops.operand(A, 0), ops.operand(B, 0), L.EPI_STORE, D=D, bias=bias, bias_mode=mode,
                 block_n=block_n)
        assert rel(D, Af @ Bf.t() + add) < 5e-3


def test_gelu_dual_output_and_dgelu_and_resid_and_mul():
    M, N, K = 520, 512, 192
    A, Af = make_operand(M, K, 0, 0, 1)
    B, Bf = make_operand(N, K, 0, 0, 2)
    A, Af = A * 0.1, Af * 0.1
    bias = rnd(N, seed=3)
    acc = (A.float() @ Bf.t())
    z_ref = acc + bias.float()
    Z = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    H = torch.zeros_like(Z)
    ops.gemm(M, N, K, ops.operand(A, 0), ops.operand(B, 0), L.EPI_GELU, D=Z, D2=H, bias=bias, bias_mode=1)
    assert rel(Z, dgelu(z_ref)) < 5e-3              # D  = gelu'(z): what backward multiplies by
    assert rel(H, gelu(z_ref)) < 5e-3               # D2 = gelu(z)
    H2 = torch.zeros_like(Z)
    ops.gemm(M, N, K, ops.operand(A, 0), ops.operand(B, 0), L.EPI_GELU_ONLY, D=H2, bias=bias, bias_mode=1)
    assert rel(H2, gelu(z_ref)) < 5e-3
    aux = rnd(M, N, seed=5)
    D = torch.zeros_like(Z)
    ops.gemm(M, N, K, ops.operand(A, 0), ops.operand(B, 0), L.EPI_DGELU, D=D, aux=aux)
    assert rel(D, acc * aux.float()) < 5e-3         # aux holds the saved gelu'(z)
    cs = rnd(N, seed=6)
    ops.gemm(M, N, K, ops.operand(A, 0), ops.operand(B, 0), L.EPI_RESID, D=D, aux=aux, bias=bias, bias_mode=1, colscale=cs)
    assert rel(D, z_ref * cs.float()[None, :] + aux.float()) < 5e-3
    ops.gemm(M, N, K, ops.operand(A, 0), ops.operand(B, 0), L.EPI_MUL, D=D, aux=aux, bias=bias, bias_mode=1)
    assert rel(D, z_ref * aux.float()) < 5e-3


def test_batched_token_mixing_shapes():
    """Z[b] [Ds, C] = W [Ds, N] (shared, padded pitch) * X[b] [N, C] (MN-major) -- the Mixer token GEMM at N=196."""
    Bn, N, C, Ds = 3, 196, 192, 784
    W = rnd(Ds, N, seed=1) * 0.1
    Wp = torch.zeros(Ds, 200, dtype=torch.bfloat16, device=DEV)
    Wp[:, :N] = W
    X = rnd(Bn, N, C, seed=2)
    Z = torch.zeros(Bn, Ds, C, dtype=torch.bfloat16, device=DEV)
    a = L.Operand(Wp.data_ptr(), Ds, N, 200, 0, 0)
    ops.gemm(Ds, C, N, a, ops.operand(X, 1), L.EPI_STORE, batch=Bn, D=Z)
    ref = torch.einsum("mn,bnc->bmc", W.float(), X.float())
    assert rel(Z, ref) < 5e-3
    # transposed use of the same padded weight: dX[b] [N, C] = W^T [N, Ds] * Z[b] [Ds, C]
    dX = torch.zeros(Bn, N, C, dtype=torch.bfloat16, device=DEV)
    a = L.Operand(Wp.data_ptr(), Ds, N, 200, 0, 1)
    ops.gemm(N, C, Ds, a, ops.operand(Z, 1), L.EPI_STORE, batch=Bn, D=dX)
    assert rel(dX, torch.einsum("mn,bmc->bnc", W.float(), Z.float())) < 5e-3


@pytest.mark.parametrize("split_k", [0, 1, 3, 7])
def test_atomic_split_k_weight_gradient(split_k):
    """dW [Dout, Din] += dY^T [Dout, R] * X [R, Din]: both operands MN-major, contraction over R rows."""
    R, Dout, Din = 2000, 264, 200
    dY, X = rnd(R, Dout, seed=1), rnd(R, Din, seed=2)
    out = torch.zeros(Dout, Din, dtype=torch.float32, device=DEV)
    ops.gemm(Dout, Din, R, ops.operand(dY, 1), ops.operand(X, 1), L.EPI_ATOMIC, out_f32=out, split_k=split_k)
    assert rel(out, dY.float().t() @ X.float()) < 2e-3


def test_atomic_contract_over_batch():
    """dW [N, Ds] += sum_b dU[b] [N, C] * H[b]^T: K-major operands, contraction over (batch, channels)."""
    Bn, N, C, Ds = 5, 196, 128, 392
    dU, H = rnd(Bn, N, C, seed=1), rnd(Bn, Ds, C, seed=2)
    out = torch.zeros(N, Ds, dtype=torch.float32, device=DEV)
    ops.gemm(N, Ds, C, ops.operand(dU, 0), ops.operand(H, 0), L.EPI_ATOMIC, batch=Bn, contract_batch=True, out_f32=out)
    assert rel(out, torch.einsum("bnc,bmc->nm", dU.float(), H.float())) < 2e-3


def test_large_persistent_many_tiles():
    M, N, K = 128 * 40, 256 * 9, 512           # 360 tiles > 148 SMs: exercises the persistent loop + TMEM double buffer
    A, Af = make_operand(M, K, 0, 0, 1)
    B, Bf = make_operand(N, K, 1, 0, 2)
    D = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(M, N, K, ops.operand(A, 0), ops.operand(B, 1), L.EPI_STORE, D=D)
    assert rel(D, Af @ Bf.t()) < 5e-3


def test_bad_arguments_raise():
    A = rnd(128, 60)
    with pytest.raises(ValueError):      # row pitch 120 B is not a multiple of 16
        ops.gemm(128, 128, 60, ops.operand(A, 0), ops.operand(A, 0), L.EPI_STORE, D=torch.zeros(128, 128, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(TypeError):
        ops.gemm(128, 128, 64, ops.operand(rnd(128, 64), 0), ops.operand(rnd(128, 64), 0), L.EPI_STORE,
                 D=torch.zeros(128, 128, device=DEV))


# ------------------------------------------------------------------------------------------------ cta_group::2 (CTA pair)
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 512, 256), (1024, 768, 768), (600, 520, 200), (256 * 9, 256 * 5, 512)])
@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_cta_pair_plain_store(M, N, K, a_major, b_major):
    A, Af = make_operand(M, K, a_major, 0, 1)
    B, Bf = make_operand(N, K, b_major, 0, 2)
    D = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(M, N, K, ops.operand(A, a_major), ops.operand(B, b_major), L.EPI_STORE, D=D, cta_group=2)
    torch.cuda.synchronize()
    assert rel(D, Af @ Bf.t()) < 5e-3


def test_cta_pair_epilogues_and_split_k():
    M, N, K = 1024 + 64, 512, 320
    A, Af = make_operand(M, K, 0, 0, 1)
    B, Bf = make_operand(N, K, 0, 0, 2)
    A, Af = A * 0.1, Af * 0.1
    bias, aux = rnd(N, seed=3), rnd(M, N, seed=5)
    acc = A.float() @ Bf.t()
    Z, H = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV), torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
    ops.gemm(M, N, K, ops.operand(A, 0), ops.operand(B, 0), L.EPI_GELU, D=Z, D2=H, bias=bias, bias_mode=1, cta_group=2)
    assert rel(Z, dgelu(acc + bias.float())) < 5e-3 and rel(H, gelu(acc + bias.float())) < 5e-3
    D = torch.zeros_like(Z)
    ops.gemm(M, N, K, ops.operand(A, 0), ops.operand(B, 0), L.EPI_DGELU, D=D, aux=aux, cta_group=2)
    assert rel(D, acc * aux.float()) < 5e-3
    ops.gemm(M, N, K, ops.operand(A, 0), ops.operand(B, 0), L.EPI_RESID, D=D, aux=aux, bias=bias, bias_mode=1, cta_group=2)
    assert rel(D, acc + bias.float() + aux.float()) < 5e-3
    R, Dout, Din = 3000, 512, 264
    dY, X = rnd(R, Dout, seed=1), rnd(R, Din, seed=2)
    for sk in (0, 1, 5):
        out = torch.zeros(Dout, Din, dtype=torch.float32, device=DEV)
        ops.gemm(Dout, Din, R, ops.operand(dY, 1), ops.operand(X, 1), L.EPI_ATOMIC, out_f32=out, split_k=sk, cta_group=2)
        assert rel(out, dY.float().t() @ X.float()) < 2e-3


@pytest.mark.parametrize("cg", [1, 2])
def test_fused_bias_gradient_reductions(cg):
    """red_mode 1/2: column / row sums of the stored (bf16) output, accumulated in fp32 by the epilogue."""
    M, N, K = 1000, 520, 256
    A, Af = make_operand(M, K, 0, 0, 1)
    B, Bf = make_operand(N, K, 0, 0, 2)
    aux = rnd(M, N, seed=5)
    for mode in (1, 2):
        D = torch.zeros(M, N, dtype=torch.bfloat16, device=DEV)
        red = torch.zeros(N if mode == 1 else M, dtype=torch.float32, device=DEV)
        ops.gemm(M, N, K, ops.operand(A, 0), ops.operand(B, 0), L.EPI_DGELU, D=D, aux=aux, red_out=red, red_mode=mode, cta_group=cg)
        assert rel(D, (Af @ Bf.t()) * aux.float()) < 5e-3
        assert rel(red, D.float().sum(0 if mode == 1 else 1)) < 1e-4      # exact sums of what was stored
    # batched (token-mixing shape): rows summed over batch and columns
    Bn, Nt, C, Ds = 3, 40, 64, 136
    W = rnd(Nt, Ds, seed=7)
    dU, G = rnd(Bn, Nt, C, seed=8), rnd(Bn, Ds, C, seed=9)
    dZ = torch.zeros(Bn, Ds, C, dtype=torch.bfloat16, device=DEV)
    red = torch.zeros(Ds, dtype=torch.float32, device=DEV)
    ops.gemm(Ds, C, Nt, ops.operand(W, 1), ops.operand(dU, 1), L.EPI_DGELU, batch=Bn, D=dZ, aux=G, red_out=red, red_mode=2, cta_group=1)
    assert rel(red, dZ.float().sum((0, 2))) < 1e-4


@pytest.mark.parametrize("Nt,Ds", [(196, 784), (144, 136), (208, 128)])
def test_token_weight_gradient_tile_208_and_transposed_output(Nt, Ds):
    """The token weight gradients of the fused Mixer block: out [Ds, Nt] with Nt in (128, 208] takes the 208-column
    tile (one N tile instead of a 256-wide one), A is the transposed hidden tensor [B, C, Ds] (MN-major), contraction
    over (batch, channels); out_trans adds the result transposed into a [Nt, Ds] gradient."""
    Bn, C = 6, 192
    HT, dU = rnd(Bn, C, Ds, seed=1), rnd(Bn, Nt, C, seed=2)
    ref = torch.einsum("bcm,bnc->mn", HT.float(), dU.float())
    out = torch.zeros(Ds, Nt, dtype=torch.float32, device=DEV)
    ops.gemm(Ds, Nt, C, ops.operand(HT, 1), ops.operand(dU, 0), L.EPI_ATOMIC, batch=Bn, contract_batch=True, out_f32=out)
    assert rel(out, ref) < 2e-3
    out_t = torch.zeros(Nt, Ds, dtype=torch.float32, device=DEV)
    ops.gemm(Ds, Nt, C, ops.operand(HT, 1), ops.operand(dU, 0), L.EPI_ATOMIC, batch=Bn, contract_batch=True, out_f32=out_t,
             out_trans=True)
    assert rel(out_t, ref.t()) < 2e-3
    out256 = torch.zeros(Ds, Nt, dtype=torch.float32, device=DEV)
    ops.gemm(Ds, Nt, C, ops.operand(HT, 1), ops.operand(dU, 0), L.EPI_ATOMIC, batch=Bn, contract_batch=True, out_f32=out256,
             block_n=256)
    assert rel(out256, ref) < 2e-3
