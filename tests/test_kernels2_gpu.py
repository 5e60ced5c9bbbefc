"""Second-round kernel paths against torch fp32 on the same bf16 inputs: fused column-sum pairs, strided column sums,
row-packed LayerNorm, the double-buffered depthwise stencil with several tiles per block, and the single-node S2v2 split
attention against its two-node composition."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from jittor_mlp_b200 import _lib as L, fn, fn_s2, fn_spatial, ops  # noqa: E402

DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).bfloat16()


def rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.mark.parametrize("rows,C", [(20000, 96), (4099, 384), (513, 768), (70000, 64), (9, 2048)])
def test_colsum2_and_strided_colsum(rows, C):
    lib = L.lib()
    a, b = rnd(rows, C, seed=1), rnd(rows, C, seed=2)
    s1 = torch.zeros(C, dtype=torch.float32, device=DEV)
    s2 = torch.zeros_like(s1)
    L.check(lib.vmlp_colsum2(a.data_ptr(), b.data_ptr(), s1.data_ptr(), s2.data_ptr(), rows, C, L.stream_ptr()))
    assert rel(s1, a.float().sum(0)) < 1e-4
    assert rel(s2, (a.float() * b.float()).sum(0)) < 1e-4
    q1, q2 = torch.zeros_like(s1), torch.zeros_like(s1)          # b aliasing a: sum and sum of squares (BatchNorm stats)
    L.check(lib.vmlp_colsum2(a.data_ptr(), a.data_ptr(), q1.data_ptr(), q2.data_ptr(), rows, C, L.stream_ptr()))
    assert rel(q2, a.float().square().sum(0)) < 1e-4
    # strided views: the right half of a [rows, 2C] tensor (gMLP's chunk), with and without the product operand
    wide, wide2 = rnd(rows, 2 * C, seed=3), rnd(rows, 2 * C, seed=4)
    v, v2 = wide[:, C:], wide2[:, C:]
    o = torch.zeros_like(s1)
    L.check(lib.vmlp_colsum(v.data_ptr(), 2 * C, 0, 0, o.data_ptr(), rows, C, L.stream_ptr()))
    assert rel(o, v.float().sum(0)) < 1e-4
    o2 = torch.zeros_like(s1)
    L.check(lib.vmlp_colsum(v.data_ptr(), 2 * C, v2.data_ptr(), 2 * C, o2.data_ptr(), rows, C, L.stream_ptr()))
    assert rel(o2, (v.float() * v2.float()).sum(0)) < 1e-4


@pytest.mark.parametrize("rows,C", [(5003, 128), (4099, 64), (12345, 32), (2051, 104)])
def test_layernorm_row_packing(rows, C):
    """C <= 128: 2 or 4 rows share a warp; ragged row counts exercise the partially filled last warp."""
    x = rnd(rows, C, seed=1) * 2 + 0.5
    g, b = rnd(C, seed=2) * 0.2 + 1, rnd(C, seed=3) * 0.2
    y, mean, rstd = ops.layernorm_fwd(x, g, b)
    xf = x.float().requires_grad_(True)
    gf, bf_ = g.float().requires_grad_(True), b.float().requires_grad_(True)
    ref = F.layer_norm(xf, (C,), gf, bf_, 1e-5)
    assert rel(y, ref) < 4e-3
    assert rel(mean, xf.mean(-1)) < 1e-5
    assert rel(rstd, (xf.var(-1, unbiased=False) + 1e-5).rsqrt()) < 1e-4
    dy, add = rnd(rows, C, seed=4), rnd(rows, C, seed=5)
    ref.backward(dy.float())
    dx, dg, db = ops.layernorm_bwd(dy, x, mean, rstd, g, add)
    assert rel(dx, xf.grad + add.float()) < 4e-3
    assert rel(dg, gf.grad) < 2e-3
    assert rel(db, bf_.grad) < 2e-3


@pytest.mark.parametrize("K,C,B,H,W", [(7, 64, 160, 32, 32), (9, 72, 96, 24, 20), (3, 128, 64, 40, 33)])
def test_depthwise_many_tiles_per_block(K, C, B, H, W):
    """Enough tiles that every persistent block walks several of them: the TMA double buffer, the mbarrier parities
    and the cross-tile weight-gradient accumulators are exercised (the small fixtures give each block one tile)."""
    x = rnd(B, H, W, C, seed=0)
    w = rnd(C, 1, K, K, seed=1, scale=0.2)
    b = rnd(C, seed=2, scale=0.2)
    xr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    wr, br = w.float().requires_grad_(True), b.float().requires_grad_(True)
    ref = F.gelu(F.conv2d(xr, wr, br, padding=K // 2, groups=C))
    dy = rnd(B, H, W, C, seed=3)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    xg, wg, bg = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    out = fn_spatial.DwConvGeluFn.apply(xg, wg, bg)
    out.backward(dy)
    assert rel(out.float().permute(0, 3, 1, 2), ref) < 6e-3
    assert rel(xg.grad.float().permute(0, 3, 1, 2), xr.grad) < 1e-2
    assert rel(wg.grad, wr.grad) < 1e-2
    assert rel(bg.grad, br.grad) < 1e-2


@pytest.mark.parametrize("B,H,W,C", [(3, 8, 8, 192), (2, 5, 7, 72)])
def test_s2v2_split_attention_single_node_matches_two_nodes(B, H, W, C):
    t = rnd(B, H, W, 3 * C, seed=0)
    w1, w2 = rnd(C, C, seed=1, scale=0.05), rnd(3 * C, C, seed=2, scale=0.05)
    do = rnd(B, H, W, C, seed=3)

    def two_nodes(t_, w1_, w2_):
        a = fn_s2.S2v2SumFn.apply(t_)
        hat = fn.linear(fn.linear_gelu(a, w1_, None), w2_, None)
        return fn_s2.S2v2CombineFn.apply(t_, hat)

    outs = []
    for f in (two_nodes, fn_s2.S2v2SplitAttentionFn.apply):
        tt, a1, a2 = (v.clone().requires_grad_(True) for v in (t, w1, w2))
        o = f(tt, a1, a2)
        o.backward(do)
        outs.append((o, tt.grad, a1.grad, a2.grad))
    (o0, g0, u0, v0), (o1, g1, u1, v1) = outs
    # the single node feeds the pooled vector to the first Linear as hi + lo (~16 mantissa bits) where the two-node form
    # rounds it to bf16: outputs agree to bf16 noise, not bit for bit
    assert rel(o1, o0) < 5e-3
    assert rel(g1, g0) < 1.2e-2        # one bf16 rounding of the summed gradient instead of two roundings + an add
    assert rel(u1, u0) < 2e-2 and rel(v1, v0) < 2e-2    # tiny-MLP weight gradients see the pooled vector's extra bits directly
