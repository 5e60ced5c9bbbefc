"""Import the UNMODIFIED reference modules from /root/reference/models_pytorch (authoring container only).

TEST INFRASTRUCTURE.  `import models_pytorch` fails as shipped: utils/__init__.py:2 pulls in
shift_cuda -> `import cupy` (shift_cuda.py:9), as_mlp.py:5 needs timm, __init__.py:21 imports a broken module.
This loader therefore:
  1. stubs `cupy` (only cupy._util.memoize is touched at import time, shift_cuda.py:23) and
     `timm.models.layers` (DropPath, to_2tuple, trunc_normal_, as_mlp.py:5);
  2. registers a synthetic package object so relative imports resolve without running __init__.py;
  3. routes Shift.forward to the reference's own CPU formulation torch_shift (shift_cuda.py:195-205);
  4. replaces the S2-MLP in-place overlapping-slice shifts (s2_mlp_v1.py:19-25, s2_mlp_v2.py:15-29; undefined
     behaviour) by clone-semantics versions = the intended function.
Nothing here is reachable on the GPU box (no /root/reference there): used to pin oracle.restate and to
generate tests/golden/*.pt (oracle/gen_golden.py).
"""
import importlib
import os
import sys
import types

import torch
from torch import nn

REF_ROOTS = ["/root/reference"]


def available():
    return any(os.path.isdir(os.path.join(r, "models_pytorch")) for r in REF_ROOTS)


def _root():
    for r in REF_ROOTS:
        if os.path.isdir(os.path.join(r, "models_pytorch")):
            return r
    raise FileNotFoundError("reference checkout not found (expected /root/reference)")


class _DropPath(nn.Module):
    """timm.models.layers.DropPath semantics: per-sample Bernoulli keep with 1/keep scaling in train()."""

    def __init__(self, drop_prob=0.):
        super().__init__()
        self.drop_prob = float(drop_prob or 0.)

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


def _install_stubs():
    if "cupy" not in sys.modules:
        cupy = types.ModuleType("cupy")
        util = types.ModuleType("cupy._util")
        util.memoize = lambda **kw: (lambda f: f)
        cupy._util = util
        cupy.cuda = types.ModuleType("cupy.cuda")
        sys.modules["cupy"], sys.modules["cupy._util"], sys.modules["cupy.cuda"] = cupy, util, cupy.cuda
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm")
        models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = _DropPath
        layers.to_2tuple = lambda v: v if isinstance(v, tuple) else (v, v)
        layers.trunc_normal_ = nn.init.trunc_normal_
        timm.models, models.layers = models, layers
        sys.modules["timm"], sys.modules["timm.models"], sys.modules["timm.models.layers"] = timm, models, layers


_loaded = {}


def _clone_shift(x, plan):
    """Intended semantics of the S2-MLP spatial shifts: out[i] = in[clamp(i - delta)] per channel quarter."""
    c = x.shape[-1]
    src = x.clone()
    q = [(0, c // 4), (c // 4, c // 2), (c // 2, c * 3 // 4), (c * 3 // 4, c)]
    for (lo, hi), (axis, delta) in zip(q, plan):
        n = x.shape[axis]
        dst = [slice(None)] * 4
        s = [slice(None)] * 4
        dst[3] = s[3] = slice(lo, hi)
        if delta > 0:
            dst[axis], s[axis] = slice(1, None), slice(0, n - 1)
        else:
            dst[axis], s[axis] = slice(0, n - 1), slice(1, None)
        x[tuple(dst)] = src[tuple(s)]
    return x


PLAN1 = [(1, +1), (1, -1), (2, +1), (2, -1)]
PLAN2 = [(2, +1), (2, -1), (1, +1), (1, -1)]


def load(name):
    """Return the reference module models_pytorch.<name> (e.g. 'mlp_mixer')."""
    if name in _loaded:
        return _loaded[name]
    _install_stubs()
    root = _root()
    if "models_pytorch" not in sys.modules:
        pkg = types.ModuleType("models_pytorch")
        pkg.__path__ = [os.path.join(root, "models_pytorch")]
        sys.modules["models_pytorch"] = pkg
    mod = importlib.import_module(f"models_pytorch.{name}")
    if name == "as_mlp":
        sc = importlib.import_module("models_pytorch.utils.shift_cuda")
        sc.Shift.forward = lambda self, x: x if self.kernel_size == 1 else sc.torch_shift(x, self.kernel_size, self.dim)
    if name == "s2_mlp_v1":
        mod.Spatial_Shift.forward = lambda self, x: _clone_shift(x, PLAN1)
    if name == "s2_mlp_v2":
        mod.spatial_shift1 = lambda x: _clone_shift(x, PLAN1)
        mod.spatial_shift2 = lambda x: _clone_shift(x, PLAN2)
    _loaded[name] = mod
    return mod


def randomize_(model, scale=0.1, seed=0):
    """Perturb EVERY parameter (default ResMLP layer-scale 1e-5 would hide block errors)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(scale * torch.randn(p.shape, generator=g, dtype=p.dtype))
    return model
