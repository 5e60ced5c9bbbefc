"""The oracle restatements (oracle/restate.py) against (a) the committed golden vectors produced by the real
reference modules and (b), when /root/reference is present, the live reference modules."""
import pytest
import torch

from oracle import models, ref_loader, restate

GOLDEN = ["mixer_tiny", "mixer_ragged", "resmlp_tiny", "gmlp_tiny", "s2v1_tiny", "s2v2_tiny", "asmlp_tiny", "hire_tiny",
          "convmixer_tiny", "vip_tiny", "vip_sum_tiny", "sparsemlp_tiny"]


@pytest.mark.parametrize("name", GOLDEN)
def test_restatement_matches_golden(golden, name):
    fx = golden(name)
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in fx["state_dict"].items()}
    x = fx["x"].clone().requires_grad_(True)
    out = models.forward(fx["cls"], fx["kwargs"], sd, x)
    assert restate.rel_l2(out, fx["out"]) < 1e-5            # fp32 tolerance (north_star: 1e-4)
    assert restate.compare_py_metric(out, fx["out"]) < 1e-4  # the reference's own metric (compare.py:179-186)
    out.square().mean().backward()
    assert restate.rel_l2(x.grad, fx["dx"]) < 1e-4
    for k, g in fx["grads"].items():
        if g is None:
            assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k   # unused params
        else:
            # absolute floor: some gradients are mathematically zero (a per-token constant is removed by the
            # following LayerNorm), leaving only fp32 noise on both sides
            assert restate.rel_l2(sd[k].grad, g) < 1e-4 or float((sd[k].grad - g).abs().max()) < 1e-7, k


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("mod,cls,kw,fn", [
    ("mlp_mixer", "MLPMixerForImageClassification", dict(d_model=128, depth=2, image_size=64, patch_size=16), restate.mixer_forward),
    ("res_mlp", "ResMLPForImageClassification", dict(d_model=96, depth=3, image_size=64, patch_size=16), restate.resmlp_forward),
    ("g_mlp", "gMLPForImageClassification", dict(d_model=64, d_ffn=96, depth=2, image_size=64, patch_size=16), restate.gmlp_forward),
])
def test_restatement_matches_live_reference(mod, cls, kw, fn):
    torch.manual_seed(3)
    model = getattr(ref_loader.load(mod), cls)(**kw)
    ref_loader.randomize_(model, 0.1, seed=4)
    x = torch.randn(2, 3, 64, 64)
    with torch.no_grad():
        ref = model(x)
        out = fn(model.state_dict(), x, kw["depth"])
    assert restate.rel_l2(out, ref) < 1e-5


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_config1_mixer_s16_cpu_plumbing():
    """BASELINE config 1: MLP-Mixer-S/16 forward, 1x3x224x224, reference on CPU vs the restatement."""
    torch.manual_seed(0)
    model = ref_loader.load("mlp_mixer").MLPMixerForImageClassification(d_model=512, depth=8).eval()
    x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        assert restate.rel_l2(restate.mixer_forward(model.state_dict(), x, 8), model(x)) < 1e-4


def test_aten_fast_path_of_the_port_equals_the_elementary_restatement(golden):
    """bench.py times the port with USE_ATEN=True; it must be the same function."""
    fx = golden("mixer_tiny")
    a = restate.mixer_forward(fx["state_dict"], fx["x"], fx["kwargs"]["depth"])
    restate.USE_ATEN = True
    try:
        b = restate.mixer_forward(fx["state_dict"], fx["x"], fx["kwargs"]["depth"])
    finally:
        restate.USE_ATEN = False
    assert restate.rel_l2(b, a) < 1e-5


def test_optimizer_restatements_match_torch_optim():
    """the reference has no optimizer, so the oracle is pinned against torch.optim itself."""
    g0 = torch.Generator().manual_seed(5)
    w0 = torch.randn(37, 11, generator=g0)
    grads = [torch.randn(37, 11, generator=g0) for _ in range(4)]
    p = torch.nn.Parameter(w0.clone())
    opt = torch.optim.AdamW([p], lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    w, m, v = w0.clone(), torch.zeros_like(w0), torch.zeros_like(w0)
    for t, g in enumerate(grads, 1):
        p.grad = g.clone()
        opt.step()
        w, m, v = restate.adamw_step(w, g, m, v, t, 3e-3, 0.9, 0.95, 1e-8, 0.05)
        assert float((w - p.detach()).abs().max()) < 1e-6
    p = torch.nn.Parameter(w0.clone())
    opt = torch.optim.SGD([p], lr=1e-2, momentum=0.9, weight_decay=1e-3)
    w, buf = w0.clone(), torch.zeros_like(w0)
    for t, g in enumerate(grads, 1):
        p.grad = g.clone()
        opt.step()
        w, buf = restate.sgd_step(w, g, buf, t, 1e-2, 0.9, 1e-3)
        assert float((w - p.detach()).abs().max()) < 1e-6
