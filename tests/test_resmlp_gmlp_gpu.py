"""ResMLP and gMLP drop-in parity on the GPU against the reference golden vectors and the oracle restatement."""
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


@pytest.mark.parametrize("name", ["resmlp_tiny", "gmlp_tiny"])
def test_against_reference_golden(golden, name):
    fx = golden(name)
    m = getattr(J, fx["cls"])(**fx["kwargs"])
    m.load_state_dict(fx["state_dict"], strict=True)
    out, dx, grads = run_model(m, fx["x"])
    assert restate.rel_l2(out.cpu(), fx["out"]) < TOL
    assert restate.rel_l2(dx.cpu(), fx["dx"]) < 2 * TOL
    scale = float(fx["dx"].abs().max() + 1)
    for k, g in fx["grads"].items():
        if g is None:
            assert grads[k] is None, k              # never-used parameters keep grad None
            continue
        err = restate.rel_l2(grads[k].cpu(), g)
        assert err < 3 * TOL or float((grads[k].cpu().float() - g).abs().max()) < 1e-4 * scale, (k, err)


@pytest.mark.parametrize("which", ["resmlp", "gmlp"])
def test_block_against_oracle_config3_shapes(which):
    """One block at the BASELINE config-3 shapes (ResMLP-24: C 384; gMLP-S: C 256, F 1536; N 196), batch 4."""
    torch.manual_seed(0)
    if which == "resmlp":
        m, fn, pre = J.ResMLP(196, 384, 1, 4), restate.resmlp_block, "model.0."
        C = 384
    else:
        m, fn, pre = J.gMLP(256, 1536, 196, 1), restate.gmlp_block, "model.0."
        C = 256
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.05 * torch.randn_like(p))
    sd = {k: v.detach().clone().bfloat16().float().requires_grad_(True) for k, v in m.state_dict().items()}
    x = torch.randn(4, 196, C, generator=torch.Generator().manual_seed(1)).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    ref = fn(sd, pre, xr)
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
