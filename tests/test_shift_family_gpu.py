"""Shift family on the GPU: S2-MLP v1/v2, AS-MLP (and later Hire-MLP, ConvMixer) against the reference golden vectors,
plus operator-level checks of the spatial kernels against the oracle restatement."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import jittor_mlp_b200 as J  # noqa: E402
from jittor_mlp_b200 import fn, fn_s2, fn_spatial  # noqa: E402
from oracle import models, restate  # noqa: E402

DEV = "cuda"
TOL = 1e-2


def bf(t):
    return t.to(DEV).bfloat16()


def bf16_floor(cls, kw, sd, x):
    """Noise floor of a fixture: the REFERENCE ALGORITHM ITSELF (the oracle restatement) evaluated by ATen on the CPU with
    bf16 storage of every tensor, against its own fp32 result -- what no bf16 implementation of these modules can beat
    by much.  Returns rel-L2 errors {"out", "dx", parameter name: ...}.  Used only where a check exceeds the north-star
    bound, to tell arithmetic errors from the conditioning of a small randomised fixture."""
    xr = x.clone().requires_grad_(True)
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    ref = models.forward(cls, kw, sdg, xr)
    ref.square().mean().backward()
    xb = x.bfloat16().requires_grad_(True)
    sdb = {k: (v.bfloat16().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    out = models.forward(cls, kw, sdb, xb)
    out.float().square().mean().backward()
    fl = {"out": restate.rel_l2(out.float(), ref), "dx": restate.rel_l2(xb.grad.float(), xr.grad)}
    for k in sd:
        if sdg[k].grad is not None and sdb[k].grad is not None:
            fl[k] = restate.rel_l2(sdb[k].grad.float(), sdg[k].grad)
    return fl


class Bound:
    """err < tol, or -- for fixtures whose bf16 noise floor is above tol -- err < factor * floor (measured lazily, printed)."""

    def __init__(self, cls, kw, sd, x):
        self.args, self.fl = (cls, kw, sd, x), None

    def check(self, key, err, tol, factor, what):
        if err < tol:
            return
        if self.fl is None:
            self.fl = bf16_floor(*self.args)
        floor = self.fl.get(key, 0.0)
        print(f"{what}: rel-L2 {err:.4f} > {tol}; bf16 floor of the reference algorithm on this fixture {floor:.4f}")
        assert err < factor * floor, (what, err, floor)


def run_model(model, x):
    model = model.to(DEV).bfloat16().train()
    xg = bf(x).requires_grad_(True)
    out = model(xg)
    out.float().square().mean().backward()
    return out, xg.grad, {k: p.grad for k, p in model.named_parameters()}


@pytest.mark.parametrize("name", ["s2v1_tiny", "s2v2_tiny", "asmlp_tiny", "hire_tiny", "convmixer_tiny"])
def test_against_reference_golden(golden, name):
    fx = golden(name)
    m = getattr(J, fx["cls"])(**fx["kwargs"])
    m.load_state_dict(fx["state_dict"], strict=True)
    out, dx, grads = run_model(m, fx["x"])
    # (s2v2_tiny was regenerated in round 2 with O(1) split-attention logits -- oracle/gen_golden.py -- and now holds the
    # same bounds as the other fixtures, gradients included)
    bound = Bound(fx["cls"], fx["kwargs"], fx["state_dict"], fx["x"])
    # 1e-2 (north star); a tiny randomised fixture whose own bf16 floor is close to that gets 2x its floor (s2v2_tiny:
    # measured 1.03e-2 against a floor of 0.61e-2 -- two independent bf16 roundings of the same computation)
    bound.check("out", restate.rel_l2(out.cpu(), fx["out"]), TOL, 2.0, name + " forward")
    assert restate.rel_l2(dx.cpu(), fx["dx"]) < 3 * TOL
    scale = float(fx["dx"].abs().max() + 1)
    ours, refs = [], []
    for k, g in fx["grads"].items():
        if g is None:
            assert grads[k] is None, k
            continue
        err = restate.rel_l2(grads[k].cpu(), g)
        # bound: 3e-2 per tensor; tensors whose gradient is numerically zero at this fixture's scale (sums over < 100
        # rows that cancel) are held to an absolute bound instead
        assert err < 3 * TOL or float((grads[k].cpu().float() - g).abs().max()) < 1e-4 * scale, (k, err)
        ours.append(grads[k].cpu().float().flatten()); refs.append(g.flatten())
    assert restate.rel_l2(torch.cat(ours), torch.cat(refs)) < 2 * TOL      # all parameter gradients together


@pytest.mark.parametrize("C,H,W", [(96, 14, 14), (24, 5, 7), (64, 8, 8)])
@pytest.mark.parametrize("kind", ["as2", "as3", "s2p1", "s2p2"])
def test_shift_forward_and_adjoint(C, H, W, kind):
    """The gather and its adjoint: bit-exact against the index-arithmetic restatement (pure data movement)."""
    x = torch.randn(3, H, W, C, generator=torch.Generator().manual_seed(0)).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    if kind.startswith("as"):
        dim = int(kind[-1])
        ref = restate.shift_tokens(xr, restate.as_groups(C, 5, dim), False)
        xg = bf(x).requires_grad_(True)
        out = fn.axial_shift(xg, 5, dim)
    else:
        plan = int(kind[-1])
        ref = restate.shift_tokens(xr, restate.s2_groups(C, plan), True)
        xg = bf(x).requires_grad_(True)
        out = fn.s2_shift(xg, plan)
    assert torch.equal(out.float().cpu(), ref.detach())
    dy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(1)).bfloat16().float()
    ref.backward(dy)
    out.backward(bf(dy))
    # the adjoint sums at most two bf16 values per element: compare with one bf16 rounding of slack
    assert restate.rel_l2(xg.grad.cpu(), xr.grad) < 4e-3


def test_group_norm1_fwd_bwd():
    B, H, W, C = 4, 14, 14, 96
    x = (torch.randn(B, H, W, C, generator=torch.Generator().manual_seed(0)) * 2 + 0.3).bfloat16().float()
    w = (torch.randn(C, generator=torch.Generator().manual_seed(1)) * 0.2 + 1).bfloat16().float()
    b = (torch.randn(C, generator=torch.Generator().manual_seed(2)) * 0.2).bfloat16().float()
    for gelu in (False, True):
        xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        ref = restate.group_norm1(xr, wr, br)
        if gelu:
            ref = restate.gelu(ref)
        dy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3)).bfloat16().float()
        ref.backward(dy)
        xg, wg, bg = bf(x).requires_grad_(True), bf(w).requires_grad_(True), bf(b).requires_grad_(True)
        out = fn.group_norm1(xg, wg, bg, 1e-5, gelu)
        out.backward(bf(dy))
        assert restate.rel_l2(out.cpu(), ref) < 5e-3
        assert restate.rel_l2(xg.grad.cpu(), xr.grad) < TOL
        assert restate.rel_l2(wg.grad.cpu(), wr.grad) < TOL
        assert restate.rel_l2(bg.grad.cpu(), br.grad) < TOL


def test_s2v2_split_attention_ops():
    B, H, W, C = 3, 8, 6, 48
    t = torch.randn(B, H, W, 3 * C, generator=torch.Generator().manual_seed(0)).bfloat16().float()
    hat = torch.randn(B, 3 * C, generator=torch.Generator().manual_seed(1)).bfloat16().float()
    tr, hr = t.clone().requires_grad_(True), hat.clone().requires_grad_(True)
    xs = [restate.shift_tokens(tr[..., :C], restate.s2_groups(C, 1), True),
          restate.shift_tokens(tr[..., C:2 * C], restate.s2_groups(C, 2), True), tr[..., 2 * C:]]
    a_ref = (xs[0] + xs[1] + xs[2]).sum((1, 2))
    bar = torch.softmax(hr.reshape(B, 3, C), 1)
    o_ref = sum(bar[:, k, None, None, :] * xs[k] for k in range(3))
    da = torch.randn(a_ref.shape, generator=torch.Generator().manual_seed(2)).bfloat16().float()
    do = torch.randn(o_ref.shape, generator=torch.Generator().manual_seed(3)).bfloat16().float()
    (a_ref * da).sum().backward(retain_graph=True)
    dt_a = tr.grad.clone(); tr.grad = None
    (o_ref * do).sum().backward()
    tg, hg = bf(t).requires_grad_(True), bf(hat).requires_grad_(True)
    a = fn_s2.S2v2SumFn.apply(tg)
    a.backward(bf(da))
    assert restate.rel_l2(a.cpu(), a_ref) < 5e-3
    assert restate.rel_l2(tg.grad.cpu(), dt_a) < 5e-3
    tg.grad = None
    o = fn_s2.S2v2CombineFn.apply(tg, hg)
    o.backward(bf(do))
    assert restate.rel_l2(o.cpu(), o_ref) < 5e-3
    assert restate.rel_l2(tg.grad.cpu(), tr.grad) < TOL
    assert restate.rel_l2(hg.grad.cpu(), hr.grad) < TOL


@pytest.mark.parametrize("K,C,H,W", [(3, 64, 9, 12), (5, 32, 8, 8), (7, 96, 16, 16), (9, 72, 10, 7)])
def test_depthwise_conv_gelu_fwd_bwd(K, C, H, W):
    import torch.nn.functional as F
    B = 3
    x = torch.randn(B, C, H, W, generator=torch.Generator().manual_seed(0)).bfloat16().float()
    w = (torch.randn(C, 1, K, K, generator=torch.Generator().manual_seed(1)) * 0.2).bfloat16().float()
    b = (torch.randn(C, generator=torch.Generator().manual_seed(2)) * 0.2).bfloat16().float()
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = restate.gelu(F.conv2d(xr, wr, br, padding=K // 2, groups=C))
    dy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3)).bfloat16().float()
    ref.backward(dy)
    xg = bf(x.permute(0, 2, 3, 1).contiguous()).requires_grad_(True)
    wg, bg = bf(w).requires_grad_(True), bf(b).requires_grad_(True)
    out = fn_spatial.DwConvGeluFn.apply(xg, wg, bg)
    out.backward(bf(dy.permute(0, 2, 3, 1).contiguous()))
    assert restate.rel_l2(out.cpu().permute(0, 3, 1, 2), ref) < 6e-3
    assert restate.rel_l2(xg.grad.cpu().permute(0, 3, 1, 2), xr.grad) < TOL
    assert restate.rel_l2(wg.grad.cpu(), wr.grad) < TOL
    assert restate.rel_l2(bg.grad.cpu(), br.grad) < TOL


def test_batch_norm_train_fwd_bwd_and_running_stats():
    B, H, W, C = 4, 8, 8, 64
    a = (torch.randn(B, H, W, C, generator=torch.Generator().manual_seed(0)) * 1.5 + 0.4).bfloat16().float()
    res = torch.randn(B, H, W, C, generator=torch.Generator().manual_seed(1)).bfloat16().float()
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_((torch.rand(C) + 0.5).bfloat16().float()); bn.bias.copy_((torch.randn(C) * 0.1).bfloat16().float())
    ar = a.clone().requires_grad_(True)
    ref = bn(ar.permute(0, 3, 1, 2)).permute(0, 2, 3, 1) + res
    dy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2)).bfloat16().float()
    ref.backward(dy)
    ag = bf(a).requires_grad_(True)
    g, bt = bf(bn.weight.detach()).requires_grad_(True), bf(bn.bias.detach()).requires_grad_(True)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    out = fn_spatial.BatchNormFn.apply(ag, g, bt, rm, rv, 0.1, 1e-5, bf(res))
    out.backward(bf(dy))
    assert restate.rel_l2(out.cpu(), ref) < 6e-3
    assert restate.rel_l2(ag.grad.cpu(), ar.grad) < TOL
    assert restate.rel_l2(g.grad.cpu(), bn.weight.grad) < TOL
    assert restate.rel_l2(bt.grad.cpu(), bn.bias.grad) < TOL
    assert restate.rel_l2(rm.cpu(), bn.running_mean) < 1e-3      # momentum update with the batch mean
    assert restate.rel_l2(rv.cpu(), bn.running_var) < 1e-3       # ... and the UNBIASED batch variance


@pytest.mark.parametrize("H,W,h,w,step", [(10, 8, 4, 4, 2), (7, 7, 2, 2, 1), (14, 14, 3, 3, 0), (5, 4, 3, 3, 1)])
def test_hire_region_ops_are_exact_data_movement(H, W, h, w, step):
    """build: gathered rows equal the reference pad/roll/rearrange pipeline; combine: inverse; adjoints via autograd."""
    import torch.nn.functional as F
    from einops import rearrange
    B, C = 2, 16
    x = torch.randn(B, H, W, C, generator=torch.Generator().manual_seed(0)).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    xn = xr.permute(0, 3, 1, 2)
    xp = F.pad(xn, (0, w - W % w, 0, h - H % h), "circular")
    zh_ref = rearrange(torch.roll(xp, step, 2), "b c (h g) w -> b g w (h c)", h=h)[:, :, :W]     # [i][c] feature order
    zw_ref = rearrange(torch.roll(xp, step, 3), "b c h (w g) -> b h g (w c)", w=w)[:, :H]
    xg = bf(x).requires_grad_(True)
    zh, zw = fn_spatial.HireBuildFn.apply(xg, h, w, step, step)
    assert torch.equal(zh.float().cpu(), zh_ref.detach()) and torch.equal(zw.float().cpu(), zw_ref.detach())
    d1 = torch.randn(zh_ref.shape, generator=torch.Generator().manual_seed(1)).bfloat16().float()
    d2 = torch.randn(zw_ref.shape, generator=torch.Generator().manual_seed(2)).bfloat16().float()
    ((zh_ref * d1).sum() + (zw_ref * d2).sum()).backward()
    torch.autograd.backward([zh, zw], [bf(d1), bf(d2)])
    assert restate.rel_l2(xg.grad.cpu(), xr.grad) < 5e-3


@pytest.mark.parametrize("cls,kw,xshape", [
    ("S2MLPv2", dict(image_size=56, patch_size=[7, 2], d_model=[192, 384], depth=[1, 1], expansion_factor=[3, 3], num_classes=16), (4, 3, 56, 56)),
    ("AS_MLP", dict(img_size=64, patch_size=4, embed_dim=96, depths=[1, 1], shift_size=5, num_classes=16, drop_path_rate=0.), (4, 3, 64, 64)),
    ("S2MLPv1", dict(image_size=64, patch_size=[16], d_model=[384], depth=[2], expansion_factor=[4], num_classes=16), (4, 3, 64, 64)),
    ("HireMLP", dict(d_model=[64, 128], h=[4, 3], w=[4, 3], cross_region_step=[2, 2], depth=[2, 2], num_classes=16), (4, 3, 64, 64)),
    ("ConvMixer", dict(dim=256, depth=2, kernel_size=7, patch_size=7, n_classes=16), (8, 3, 56, 56)),
])
def test_config4_channel_widths_against_oracle(cls, kw, xshape):
    """Real channel widths of BASELINE config 4 (C 96/192/384) at small spatial size: forward, input gradient and every
    parameter gradient."""
    torch.manual_seed(0)
    m = getattr(J, cls)(**kw)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.03 * torch.randn_like(p))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.randn(*xshape, generator=torch.Generator().manual_seed(1))
    xr = x.clone().requires_grad_(True)
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    ref = models.forward(cls, kw, sdg, xr)
    ref.square().mean().backward()
    out, dx, grads = run_model(m, x)
    bound = Bound(cls, kw, sd, x)
    # north star: 1e-2 on forward outputs.  Hire-MLP at these widths: measured 1.08e-2 where the reference algorithm with
    # bf16 storage is itself 1.45e-2 away from fp32 (6.7e-3 for the unperturbed Hire-MLP-T)
    bound.check("out", restate.rel_l2(out.cpu(), ref), TOL, 1.0, cls + " forward")
    assert restate.rel_l2(dx.cpu(), xr.grad) < 3 * TOL
    scale = float(xr.grad.abs().max() + 1)
    ours, refs = [], []
    for k, g in grads.items():
        rg = sdg[k].grad
        if g is None or rg is None:                       # never-used parameters on both sides
            assert g is None and (rg is None or float(rg.abs().max()) == 0), k
            continue
        err = restate.rel_l2(g.cpu(), rg)
        if float((g.cpu().float() - rg).abs().max()) >= 1e-4 * scale:
            # 3e-2 per tensor; S2-MLPv2's split-attention weights sit behind a softmax over token-SUM statistics: measured
            # 3.1e-2 where the bf16 reference algorithm shows 3.4e-2
            bound.check(k, err, 3 * TOL, 1.0, f"{cls} grad {k}")
        ours.append(g.cpu().float().flatten()); refs.append(rg.flatten())
    assert restate.rel_l2(torch.cat(ours), torch.cat(refs)) < 2 * TOL
