"""MLP-Mixer drop-in parity on the GPU: the fused block path (C ABI) against
  (1) the committed golden vectors produced by the real reference modules (fp32, CPU), and
  (2) the oracle restatement + autograd on the same seeded inputs at larger sizes.
Tolerance: bf16 storage / fp32 accumulate => rel-L2 <= 1e-2 (north_star), written per assert."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import jittor_mlp_b200 as J  # noqa: E402
from oracle import restate  # noqa: E402

DEV = "cuda"
TOL = 1e-2


def run_model(model, x):
    model = model.to(DEV).bfloat16().train()
    xg = x.to(DEV).bfloat16().requires_grad_(True)
    out = model(xg)
    out.float().square().mean().backward()
    return out, xg.grad, {k: p.grad for k, p in model.named_parameters()}


@pytest.mark.parametrize("name", ["mixer_tiny", "mixer_ragged"])
def test_against_reference_golden(golden, name):
    fx = golden(name)
    m = getattr(J, fx["cls"])(**fx["kwargs"])
    m.load_state_dict(fx["state_dict"], strict=True)
    out, dx, grads = run_model(m, fx["x"])
    assert restate.rel_l2(out.cpu(), fx["out"]) < TOL
    assert restate.rel_l2(dx.cpu(), fx["dx"]) < 2 * TOL
    for k, g in fx["grads"].items():
        err = restate.rel_l2(grads[k].cpu(), g)
        assert err < 3 * TOL or float((grads[k].cpu().float() - g).abs().max()) < 1e-4 * float(fx["dx"].abs().max() + 1), (k, err)


@pytest.mark.parametrize("C,batch", [(768, 4), (1024, 2)])
def test_block_against_oracle_b16_shapes(C, batch):
    """One block at the Mixer-B/16 (N 196, C 768, Ds 784, Dc 3072) and L/16 (C 1024, Dc 4096) shapes; fwd and all
    gradients (the output-bias gradients come out of the fused LayerNorm-backward pass)."""
    torch.manual_seed(0)
    m = J.MLPMixer(196, C, 1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.05 * torch.randn_like(p))
    sd = {k: v.detach().clone().bfloat16().float().requires_grad_(True) for k, v in m.state_dict().items()}
    x = torch.randn(batch, 196, C, generator=torch.Generator().manual_seed(1)).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    ref = restate.mixer_block(sd, "model.0.", xr)
    dy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2)).bfloat16().float()
    ref.backward(dy)
    m = m.to(DEV).bfloat16()
    xg = x.to(DEV).bfloat16().requires_grad_(True)
    y = m(xg)
    y.backward(dy.to(DEV).bfloat16())
    assert restate.rel_l2(y.cpu(), ref) < TOL
    assert restate.rel_l2(xg.grad.cpu(), xr.grad) < TOL
    for k, p in m.named_parameters():
        err = restate.rel_l2(p.grad.cpu(), sd[k].grad)
        assert err < 2 * TOL, (k, err)


def test_whole_model_against_oracle_and_determinism():
    torch.manual_seed(0)
    kw = dict(d_model=256, depth=3, image_size=64, patch_size=8, num_classes=100)
    m = J.MLPMixerForImageClassification(**kw)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.05 * torch.randn_like(p))
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.randn(8, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    ref = restate.mixer_forward(sd, x, kw["depth"])
    m = m.to(DEV).bfloat16().eval()
    with torch.no_grad():
        o1 = m(x.to(DEV).bfloat16())
        o2 = m(x.to(DEV).bfloat16())
    assert torch.equal(o1, o2)                     # forward has no atomics: bit-reproducible
    assert restate.rel_l2(o1.cpu(), ref) < TOL
    assert restate.compare_py_metric(o1.cpu(), ref) < 2e-2


def test_unsupported_inputs_raise():
    m = J.MLPMixerForImageClassification(d_model=64, depth=1, image_size=32, patch_size=8).to(DEV)
    with pytest.raises(TypeError):          # fp32 parameters: no silent fallback
        m(torch.randn(1, 3, 32, 32, device=DEV))


def test_cuda_graph_step_matches_eager():
    """GraphedStep replays the same kernels: loss and gradients equal the eager step (up to fp32-atomic ordering)."""
    torch.manual_seed(0)
    kw = dict(d_model=128, depth=2, image_size=64, patch_size=8, num_classes=10)
    m = J.MLPMixerForImageClassification(**kw).to(DEV).bfloat16().train()
    x = torch.randn(8, 3, 64, 64, device=DEV).bfloat16()
    loss_fn = lambda o: o.float().square().mean()
    m.zero_grad(set_to_none=True)
    l0 = loss_fn(m(x))
    l0.backward()
    ref = {k: p.grad.clone() for k, p in m.named_parameters()}
    l0 = float(l0)          # drop the eager autograd graph: its AccumulateGrad nodes are bound to the default stream
    gs = J.GraphedStep(m, x, loss_fn)
    x2 = torch.randn(8, 3, 64, 64, device=DEV).bfloat16()
    gs.run(x2)                                   # different input through the same graph
    l1 = gs.run(x)
    torch.cuda.synchronize()
    assert abs(float(l1) - l0) < 1e-3 * abs(l0)
    for k, p in m.named_parameters():
        assert restate.rel_l2(p.grad.cpu(), ref[k].cpu()) < 2e-3, k
