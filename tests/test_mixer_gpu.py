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


def _grad_check(named_grads, ref_grads, scale, tol=3 * TOL):
    """Every parameter gradient within `tol` (rel-L2), tensors that are numerically zero at this scale by an absolute
    bound; and all of them together within 2e-2."""
    ours, refs = [], []
    for k, g in named_grads.items():
        rg = ref_grads[k]
        err = restate.rel_l2(g.cpu(), rg.cpu())
        assert err < tol or float((g.cpu().float() - rg.cpu()).abs().max()) < 1e-4 * scale, (k, err)
        ours.append(g.cpu().float().flatten()); refs.append(rg.cpu().float().flatten())
    assert restate.rel_l2(torch.cat(ours), torch.cat(refs)) < 2 * TOL


def test_whole_mixer_b16_as_benchmarked():
    """BASELINE config 2 exactly as bench.py runs it (d_model 768, depth 12, 224 px, patch 16, 1000 classes) on 4 images
    against the fp32 oracle on the CPU: output <= 1e-2, input gradient and all 150 parameter gradients <= 3e-2
    (12 blocks of bf16 error accumulation, 196-token GEMMs through the fused token kernels)."""
    torch.manual_seed(0)
    kw = dict(d_model=768, depth=12, image_size=224, patch_size=16, num_classes=1000)
    m = J.MLPMixerForImageClassification(**kw)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.01 * torch.randn_like(p))
            p.copy_(p.bfloat16().float())                  # both sides start from the same bf16-representable weights
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    x = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(1)).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    ref = restate.mixer_forward(sd, xr, kw["depth"])
    ref.square().mean().backward()
    out, dx, grads = run_model(m, x)
    assert restate.rel_l2(out.cpu(), ref) < TOL, restate.rel_l2(out.cpu(), ref)
    assert restate.compare_py_metric(out.cpu(), ref.detach()) < 1e-2       # the reference's own parity metric (compare.py:179-186)
    assert restate.rel_l2(dx.cpu(), xr.grad) < 3 * TOL
    _grad_check(grads, {k: v.grad for k, v in sd.items()}, float(xr.grad.abs().max() + 1))


def test_block_at_the_benchmarked_batch_256():
    """One MixerBlock at the bench shape (B 256, N 196, C 768, Ds 784, Dc 3072: 5376-tile token kernels, 392-tile channel
    GEMMs, split-K weight gradients over 50176 rows) against the oracle restatement evaluated in fp32 ON THE GPU (the
    oracle is plain torch; TF32 off) -- y, dx and every parameter gradient."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(0)
    B, N, C = 256, 196, 768
    m = J.MLPMixer(N, C, 1)
    with torch.no_grad():
        for p in m.parameters():
            p.add_(0.02 * torch.randn_like(p))
    sd = {k: v.detach().clone().bfloat16().float().to(DEV).requires_grad_(True) for k, v in m.state_dict().items()}
    x = torch.randn(B, N, C, generator=torch.Generator().manual_seed(1)).bfloat16()
    dy = torch.randn(B, N, C, generator=torch.Generator().manual_seed(2)).bfloat16()
    xr = x.float().to(DEV).requires_grad_(True)
    ref = restate.mixer_block(sd, "model.0.", xr)
    ref.backward(dy.float().to(DEV))
    m = m.to(DEV).bfloat16()
    xg = x.to(DEV).requires_grad_(True)
    y = m(xg)
    y.backward(dy.to(DEV))
    torch.cuda.synchronize()
    assert restate.rel_l2(y.float(), ref.detach()) < TOL
    assert restate.rel_l2(xg.grad.float(), xr.grad) < TOL
    for b0 in (0, 100, 255):                                # per-image: no sample is left behind by the tile schedule
        assert restate.rel_l2(y[b0].float(), ref[b0].detach()) < TOL and restate.rel_l2(xg.grad[b0].float(), xr.grad[b0]) < TOL
    _grad_check({k: p.grad for k, p in m.named_parameters()}, {k: v.grad for k, v in sd.items()},
                float(xr.grad.abs().max() + 1), tol=2 * TOL)


def test_fused_and_unfused_token_paths_agree():
    """VMLP_TOKMIX=0 (read at library load) selects the unfused GEMM sequence: run it in a subprocess on the same seeded
    block and compare with the fused path of this process."""
    import os
    import subprocess
    import sys
    import tempfile
    code = r'''
import sys, torch
sys.path.insert(0, %r)
import jittor_mlp_b200 as J
torch.manual_seed(0)
m = J.MLPMixer(196, 256, 2).to("cuda").bfloat16()
x = torch.randn(6, 196, 256, generator=torch.Generator().manual_seed(1)).to("cuda").bfloat16().requires_grad_(True)
y = m(x)
y.backward(torch.ones_like(y))
torch.save({"y": y.cpu(), "dx": x.grad.cpu(), "g": {k: p.grad.cpu() for k, p in m.named_parameters()}}, sys.argv[1])
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("1", "0"):
        with tempfile.NamedTemporaryFile(suffix=".pt") as f:
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env={**os.environ, "VMLP_TOKMIX": flag}, timeout=600)
            outs.append(torch.load(f.name))
    a, b = outs
    assert restate.rel_l2(a["y"], b["y"].float()) < 5e-3 and restate.rel_l2(a["dx"], b["dx"].float()) < 5e-3
    for k in a["g"]:
        assert restate.rel_l2(a["g"][k], b["g"][k].float()) < 1e-2, k
