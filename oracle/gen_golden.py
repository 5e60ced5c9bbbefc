"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules (authoring container only).

TEST INFRASTRUCTURE.  Usage: python -m oracle.gen_golden   (needs /root/reference; see oracle/ref_loader.py)
Each fixture holds: ctor kwargs, a randomised fp32 state_dict (every parameter perturbed),
a seeded input, the reference forward output (train() mode, fp32, CPU) and the autograd gradients of
loss = out.square().mean() w.r.t. the input and every parameter.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (module, class, kwargs, input shape)
    "mixer_tiny": ("mlp_mixer", "MLPMixerForImageClassification",
                   dict(d_model=64, depth=2, image_size=32, patch_size=8, num_classes=10), (2, 3, 32, 32)),
    "mixer_ragged": ("mlp_mixer", "MLPMixerForImageClassification",
                     dict(d_model=72, depth=1, image_size=(48, 40), patch_size=8, num_classes=7, expansion_factor=4),
                     (3, 3, 48, 40)),
    "resmlp_tiny": ("res_mlp", "ResMLPForImageClassification",
                    dict(d_model=64, depth=2, image_size=32, patch_size=8, num_classes=10), (2, 3, 32, 32)),
    "s2v1_tiny": ("s2_mlp_v1", "S2MLPv1", dict(image_size=32, patch_size=[4, 2], d_model=[32, 64], depth=[1, 2],
                                               expansion_factor=[2, 3], num_classes=10), (2, 3, 32, 32)),
    "s2v2_tiny": ("s2_mlp_v2", "S2MLPv2", dict(image_size=(32, 24), patch_size=[4, 2], d_model=[24, 64], depth=[1, 2],
                                               expansion_factor=[2, 3], num_classes=10), (2, 3, 32, 24),
                  # SplitAttention pools by a token SUM (s2_mlp_v2.py:44); with every weight perturbed by 0.1 its softmax
                  # logits reach ~50 and the block output flips with one bf16 ulp of the pooled vector.  Scaled weights
                  # keep the logits O(1) like a trained model's, so the fixture tests the arithmetic, not the conditioning.
                  {"split_attention": 0.15}),
    "asmlp_tiny": ("as_mlp", "AS_MLP", dict(img_size=32, patch_size=4, embed_dim=24, depths=[1, 2], shift_size=5,
                                            num_classes=10, drop_path_rate=0.), (2, 3, 32, 32)),
    "hire_tiny": ("hire_mlp", "HireMLP", dict(patch_size=4, d_model=[16, 32], h=[4, 3], w=[4, 3], cross_region_step=[2, 1],
                                             depth=[2, 2], num_classes=10), (2, 3, 40, 32)),
    "convmixer_tiny": ("conv_mixer", "ConvMixer", dict(dim=32, depth=2, kernel_size=5, patch_size=4, n_classes=10),
                       (4, 3, 32, 32)),
    "vip_tiny": ("vip", "ViP", dict(image_size=(32, 48), patch_size=(8, 8), d_model=32, depth=2, segments=8, num_classes=10,
                                    expansion_factor=3, weighted=True), (2, 3, 32, 48), {"split_attention": 0.15}),
    "vip_sum_tiny": ("vip", "ViP", dict(image_size=(32, 32), patch_size=(8, 4), d_model=48, depth=1, segments=16,
                                        num_classes=10, weighted=False), (3, 3, 32, 32)),
    "sparsemlp_tiny": ("sparse_mlp", "SparseMLP", dict(image_size=(32, 64), patch_size=4, d_model=16, depth=[1, 2], num_classes=10,
                                                        expansion_factor=2), (3, 3, 32, 64)),
    "gmlp_tiny": ("g_mlp", "gMLPForImageClassification",
                  dict(image_size=32, patch_size=8, num_classes=10, d_model=64, d_ffn=128, depth=2), (2, 3, 32, 32)),
}


def make(name):
    mod, cls, kwargs, xshape = CASES[name][:4]
    torch.manual_seed(0)
    model = getattr(R.load(mod), cls)(**kwargs)
    R.randomize_(model, 0.1, seed=1)
    for key, scale in (CASES[name][4] if len(CASES[name]) > 4 else {}).items():
        with torch.no_grad():
            for k, p in model.named_parameters():
                if key in k:
                    p.mul_(scale)
    model.train()
    x = torch.randn(*xshape, generator=torch.Generator().manual_seed(2)).requires_grad_(True)
    out = model(x)
    loss = out.square().mean()
    loss.backward()
    fx = dict(module=mod, cls=cls, kwargs=kwargs,
              state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
              x=x.detach().clone(), out=out.detach().clone(), dx=x.grad.clone(),
              grads={k: (p.grad.clone() if p.grad is not None else None) for k, p in model.named_parameters()})
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name, tuple(out.shape), f"{os.path.getsize(os.path.join(OUT, name + '.pt')) / 1024:.0f} KiB")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for n in (sys.argv[1:] or CASES):
        make(n)
