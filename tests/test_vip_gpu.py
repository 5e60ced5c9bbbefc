"""Vision Permutator and the fused optimizer step (row f4) on the GPU, against the reference golden
vectors, the oracle restatement and torch.optim."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import jittor_mlp_b200 as J  # noqa: E402
from jittor_mlp_b200 import fn_vip  # noqa: E402
from oracle import models, restate  # noqa: E402
from test_shift_family_gpu import Bound, bf, run_model  # noqa: E402

DEV = "cuda"
TOL = 1e-2


@pytest.mark.parametrize("B,H,W,c,S", [(2, 4, 6, 4, 8), (3, 14, 28, 16, 16), (1, 7, 5, 3, 24)])
def test_permute5_is_the_einops_rearrangement_bit_exact(B, H, W, c, S):
    """`b h w (c s) -> b w c (h s)` and `-> b h c (w s)` (vip.py:68,73), their inverses into a channel slot of a wider
    buffer, and the accumulating form: pure data movement => bit-exact."""
    C = c * S
    x = torch.randn(B, H, W, C, generator=torch.Generator().manual_seed(0)).bfloat16()
    xg = x.to(DEV)
    sh, sw = fn_vip._specs(B, H, W, C, S, C)
    th = torch.empty(B * W * c, H * S, dtype=torch.bfloat16, device=DEV)
    fn_vip.permute5(xg, th, sh[0], sh[1], sh[2])
    ref_h = x.view(B, H, W, c, S).permute(0, 2, 3, 1, 4).reshape(B * W * c, H * S)
    assert torch.equal(th.cpu(), ref_h)
    tw = torch.empty(B * H * c, W * S, dtype=torch.bfloat16, device=DEV)
    fn_vip.permute5(xg, tw, sw[0], sw[1], sw[2])
    ref_w = x.view(B, H, W, c, S).permute(0, 1, 3, 2, 4).reshape(B * H * c, W * S)
    assert torch.equal(tw.cpu(), ref_w)
    # inverse into slots 0 / 1 of a [B, H, W, 3C] buffer
    wide = torch.zeros(B, H, W, 3 * C, dtype=torch.bfloat16, device=DEV)
    oh, ow = fn_vip._specs(B, H, W, C, S, 3 * C)
    fn_vip.permute5(th, wide, oh[0], oh[2], oh[1])
    fn_vip.permute5(tw, wide.view(-1)[C:], ow[0], ow[2], ow[1])
    assert torch.equal(wide[..., :C].cpu(), x) and torch.equal(wide[..., C:2 * C].cpu(), x)
    assert float(wide[..., 2 * C:].abs().max()) == 0.0
    # accumulate: out += in, fp32 add rounded once
    acc = xg.clone()
    fn_vip.permute5(th, acc, sh[0], sh[2], sh[1], accumulate=True)
    assert torch.equal(acc.cpu(), (x.float() * 2).bfloat16())


@pytest.mark.parametrize("name", ["vip_tiny", "vip_sum_tiny"])
def test_vip_against_reference_golden(golden, name):
    fx = golden(name)
    m = J.ViP(**fx["kwargs"])
    m.load_state_dict(fx["state_dict"], strict=True)
    out, dx, grads = run_model(m, fx["x"])
    bound = Bound(fx["cls"], fx["kwargs"], fx["state_dict"], fx["x"])
    bound.check("out", restate.rel_l2(out.cpu(), fx["out"]), TOL, 2.0, name + " forward")
    bound.check("dx", restate.rel_l2(dx.cpu(), fx["dx"]), 3 * TOL, 2.0, name + " dx")
    scale = float(fx["dx"].abs().max() + 1)
    ours, refs = [], []
    for k, g in fx["grads"].items():
        if g is None:
            assert grads[k] is None, k
            continue
        err = restate.rel_l2(grads[k].cpu(), g)
        if not (err < 3 * TOL or float((grads[k].cpu().float() - g).abs().max()) < 1e-4 * scale):
            bound.check(k, err, 3 * TOL, 2.0, name + " " + k)
        ours.append(grads[k].cpu().float().flatten()); refs.append(g.flatten())
    assert restate.rel_l2(torch.cat(ours), torch.cat(refs)) < 2 * TOL


def test_vip_compare_py_width_against_oracle():
    """compare.py:90-99's ViP geometry (224 px, patch (16, 8) -> 14 x 28 positions, d_model 256, 16 segments), depth 2,
    4 images: forward, dx and every parameter gradient against the fp32 oracle."""
    kw = dict(image_size=(224, 224), patch_size=(16, 8), d_model=256, depth=2, segments=16, num_classes=10, weighted=True)
    torch.manual_seed(0)
    m = J.ViP(**kw)
    with torch.no_grad():
        for k, p in m.named_parameters():
            if "split_attention" in k:
                p.mul_(0.2)                      # pooled logits O(1), see oracle/gen_golden.py
    sd = {k: v.detach().clone().float() for k, v in m.state_dict().items()}
    x = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = x.clone().requires_grad_(True)
    ref = models.forward("ViP", kw, sdg, xr)
    ref.square().mean().backward()
    out, dx, grads = run_model(m, x)
    bound = Bound("ViP", kw, sd, x)
    bound.check("out", restate.rel_l2(out.cpu(), ref.detach()), TOL, 2.0, "ViP forward")
    bound.check("dx", restate.rel_l2(dx.cpu(), xr.grad), 3 * TOL, 2.0, "ViP dx")
    ours = torch.cat([grads[k].cpu().float().flatten() for k in sd if sdg[k].grad is not None])
    refs = torch.cat([sdg[k].grad.flatten() for k in sd if sdg[k].grad is not None])
    assert restate.rel_l2(ours, refs) < 2 * TOL
    for k in sd:
        if sdg[k].grad is not None and sdg[k].grad.numel() > 256:
            bound.check(k, restate.rel_l2(grads[k].cpu(), sdg[k].grad), 3 * TOL, 2.0, "ViP " + k)


@pytest.mark.parametrize("kind", ["adamw", "sgd"])
def test_fused_optimizer_matches_torch_optim_on_fp32_master(kind):
    """Four steps over tensors of awkward sizes (unaligned views into one flat buffer, a run longer than one chunk):
    fp32 master weights equal torch.optim's fp32 result, the bf16 parameters are their rounding."""
    g0 = torch.Generator().manual_seed(7)
    sizes = [(70001,), (33, 7), (5,), (128, 64), (1,)]
    flat = torch.zeros(sum(torch.Size(s).numel() for s in sizes) + 8, dtype=torch.bfloat16, device=DEV)
    ps, off = [], 1                                         # offset 1: the first view is only 2-byte aligned
    for s in sizes:
        n = torch.Size(s).numel()
        v = flat[off:off + n].view(s)
        v.copy_(torch.randn(s, generator=g0))
        ps.append(torch.nn.Parameter(v))
        off += n
    refs = [torch.nn.Parameter(p.detach().float().clone()) for p in ps]
    if kind == "adamw":
        mine = J.FusedAdamW(ps, lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
        ref = torch.optim.AdamW(refs, lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    else:
        mine = J.FusedSGD(ps, lr=1e-2, momentum=0.9, weight_decay=1e-3)
        ref = torch.optim.SGD(refs, lr=1e-2, momentum=0.9, weight_decay=1e-3)
    n0 = J._lib.lib().vmlp_launch_count()
    for step in range(4):
        for p, r in zip(ps, refs):
            g = torch.randn(p.shape, generator=g0).bfloat16()
            p.grad = g.to(DEV)
            r.grad = g.float().to(DEV)
        if step == 2:
            ps[2].grad = None                               # a parameter without a gradient is skipped, like torch
            refs[2].grad = None
        mine.step()
        ref.step()
    assert J._lib.lib().vmlp_launch_count() - n0 == 4       # one launch per step
    for p, r in zip(ps, refs):
        mw = mine.state[p]["master"]
        assert float((mw - r.detach()).abs().max()) < 2e-6 * float(r.detach().abs().max() + 1), p.shape
        assert torch.equal(p.detach(), mw.bfloat16())
    # state interchange with the stock optimizer
    back = type(ref)(refs, lr=1.0)
    back.load_state_dict(mine.state_dict())
    key = "exp_avg" if kind == "adamw" else "momentum_buffer"
    assert float((back.state[refs[0]][key] - ref.state[refs[0]][key]).abs().max()) < 1e-6


def test_training_step_with_fused_adamw_reduces_the_loss():
    """Mixer block path + fused optimizer: a few steps on one batch must lower the loss (end-to-end wiring)."""
    torch.manual_seed(0)
    m = J.MLPMixerForImageClassification(d_model=64, depth=2, image_size=32, patch_size=8, num_classes=10).to(DEV).bfloat16()
    opt = J.FusedAdamW(m.parameters(), lr=2e-3, weight_decay=0.0)
    x = bf(torch.randn(16, 3, 32, 32))
    y = torch.randint(0, 10, (16,), device=DEV)
    losses = []
    for _ in range(8):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(m(x).float(), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.7 * losses[0], losses


@pytest.mark.parametrize("shape", [(3, 196, 768), (2, 7, 5, 24), (5, 14, 28, 256), (1, 1, 8)])
def test_token_mean_head_kernels(shape):
    """Position mean of the classification heads (mlp_mixer.py:75, hire_mlp.py:219): fp32 accumulation, rounded once."""
    from jittor_mlp_b200 import fn
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(3)).bfloat16()
    xg = x.to(DEV).requires_grad_(True)
    out = fn.TokenMeanFn.apply(xg)
    ref = x.float().reshape(shape[0], -1, shape[-1]).mean(1)
    assert float((out.float().cpu() - ref).abs().max()) <= 2.0 ** -8 * float(ref.abs().max())     # one bf16 rounding
    g = torch.randn(shape[0], shape[-1], generator=torch.Generator().manual_seed(4)).bfloat16()
    out.backward(g.to(DEV))
    inv = torch.tensor(1.0 / (x.numel() // (shape[0] * shape[-1])), dtype=torch.float32)
    dref = (g.float() * inv).bfloat16().float().reshape(shape[0], *([1] * (len(shape) - 2)), shape[-1]).expand(*shape)
    assert torch.equal(xg.grad.float().cpu(), dref)


# ------------------------------------------------------------------------------------------------ SparseMLP (row f3)
@pytest.mark.parametrize("Bt,N,C,Mo", [(6, 56, 96, 56), (4, 7, 64, 7), (3, 16, 256, 16), (5, 28, 40, 28)])
def test_token_linear_against_fp32(Bt, N, C, Mo):
    """Linear along the token axis of [Bt, N, C] (sparse_mlp.py:66-71): forward, dgrad, wgrad, bias gradient."""
    from jittor_mlp_b200 import fn
    g = torch.Generator().manual_seed(11)
    x = torch.randn(Bt, N, C, generator=g).bfloat16()
    w = (torch.randn(Mo, N, generator=g) * N ** -0.5).bfloat16()
    b = torch.randn(Mo, generator=g).bfloat16()
    dy = torch.randn(Bt, Mo, C, generator=g).bfloat16()
    xr, wr, br = (t.float().requires_grad_(True) for t in (x, w, b))
    ref = torch.einsum("mn,bnc->bmc", wr, xr) + br[None, :, None]
    ref.backward(dy.float())
    xg, wg, bg = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    out = fn.TokenLinearFn.apply(xg, wg, bg)
    out.backward(dy.to(DEV))
    assert restate.rel_l2(out.cpu(), ref.detach()) < 5e-3
    assert restate.rel_l2(xg.grad.cpu(), xr.grad) < 5e-3
    assert restate.rel_l2(wg.grad.cpu(), wr.grad) < 5e-3
    assert restate.rel_l2(bg.grad.cpu(), br.grad) < 5e-3


def test_concat_channels_and_plain_depthwise_conv():
    from jittor_mlp_b200 import fn, fn_spatial
    g = torch.Generator().manual_seed(12)
    parts = [torch.randn(2, 5, 7, c, generator=g).bfloat16() for c in (16, 32, 16)]
    pg = [t.to(DEV).requires_grad_(True) for t in parts]
    out = fn.ConcatChannelsFn.apply(*pg)
    assert torch.equal(out.cpu(), torch.cat(parts, -1))
    dout = torch.randn(2, 5, 7, 64, generator=g).bfloat16()
    out.backward(dout.to(DEV))
    for t, lo in zip(pg, (0, 16, 48)):
        assert torch.equal(t.grad.cpu(), dout[..., lo:lo + t.shape[-1]])
    # depthwise 3x3 + bias, no activation, against F.conv2d in fp32
    x = torch.randn(3, 9, 12, 48, generator=g).bfloat16()
    w = (torch.randn(48, 1, 3, 3, generator=g) * 0.3).bfloat16()
    b = torch.randn(48, generator=g).bfloat16()
    dy = torch.randn(3, 9, 12, 48, generator=g).bfloat16()
    xr, wr, br = (t.float().requires_grad_(True) for t in (x, w, b))
    ref = torch.nn.functional.conv2d(xr.permute(0, 3, 1, 2), wr, br, padding=1, groups=48).permute(0, 2, 3, 1)
    ref.backward(dy.float())
    xg, wg, bg = (t.to(DEV).requires_grad_(True) for t in (x, w, b))
    y = fn_spatial.DwConvFn.apply(xg, wg, bg)
    y.backward(dy.to(DEV))
    assert restate.rel_l2(y.cpu(), ref.detach()) < 5e-3
    assert restate.rel_l2(xg.grad.cpu(), xr.grad) < 5e-3
    assert restate.rel_l2(wg.grad.cpu(), wr.grad) < 5e-3
    assert restate.rel_l2(bg.grad.cpu(), br.grad) < 5e-3


def test_sparsemlp_against_reference_golden(golden):
    fx = golden("sparsemlp_tiny")
    m = J.SparseMLP(**fx["kwargs"])
    m.load_state_dict(fx["state_dict"], strict=True)
    out, dx, grads = run_model(m, fx["x"])
    bound = Bound(fx["cls"], fx["kwargs"], fx["state_dict"], fx["x"])
    bound.check("out", restate.rel_l2(out.cpu(), fx["out"]), TOL, 2.0, "sparsemlp forward")
    bound.check("dx", restate.rel_l2(dx.cpu(), fx["dx"]), 3 * TOL, 2.0, "sparsemlp dx")
    scale = float(fx["dx"].abs().max() + 1)
    ours, refs = [], []
    for k, g in fx["grads"].items():
        if g is None:
            assert grads[k] is None, k
            continue
        err = restate.rel_l2(grads[k].cpu(), g)
        if not (err < 3 * TOL or float((grads[k].cpu().float() - g).abs().max()) < 1e-4 * scale):
            bound.check(k, err, 3 * TOL, 2.0, "sparsemlp " + k)
        ours.append(grads[k].cpu().float().flatten()); refs.append(g.flatten())
    assert restate.rel_l2(torch.cat(ours), torch.cat(refs)) < 2 * TOL
