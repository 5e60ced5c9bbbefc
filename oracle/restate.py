"""CPU fp32 restatement of the reference's per-block algorithms (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function works from a ``state_dict`` with the reference's key names and uses only elementary torch
ops (einsum / layer_norm / gelu), written from the formulas of SURVEY.md Appendix A -- not by calling the
reference modules.  Pinned against the real reference modules by tests/test_oracle.py (in the authoring
container, through oracle.ref_loader) and against the committed golden vectors in tests/golden/ (everywhere).
Autograd through these functions is the gradient oracle.
"""
import math

import torch
import torch.nn.functional as F


# bench.py's CPU-baseline legs flip this to time the SAME ATen primitives the reference modules dispatch to
# (conv1d / layer_norm / gelu); the parity tests keep the elementary formulas and check both agree.
USE_ATEN = False


def gelu(z):
    """Exact erf GELU = nn.GELU() default (mlp_mixer.py:21)."""
    if USE_ATEN:
        return F.gelu(z)
    return 0.5 * z * (1.0 + torch.erf(z / math.sqrt(2.0)))


def layer_norm(x, w, b, eps=1e-5):
    """nn.LayerNorm over the last axis, biased variance (mlp_mixer.py:10)."""
    if USE_ATEN:
        return F.layer_norm(x, (x.shape[-1],), w, b, eps)
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def token_mix(w, x, bias):
    """nn.Conv1d(N_in, N_out, kernel_size=1) applied to [B, N_in, C] (mlp_mixer.py:34,37): contraction over tokens.
    w is the Conv1d weight [N_out, N_in, 1]."""
    if USE_ATEN:
        return F.conv1d(x, w, bias)
    return torch.einsum("mn,bnc->bmc", w[:, :, 0], x) + bias[None, :, None]


# ----------------------------------------------------------------------------------------------- MLP-Mixer
def mixer_block(sd, pre, x):
    """One MLPMixer.model[i] (mlp_mixer.py:36-39); x [B, N, C]; `pre` = 'model.{i}.'.

    token half  (mlp_mixer.py:12-13,19-25,37): Conv1d(k=1) over the token axis == W[m, n] contraction over n.
    channel half (mlp_mixer.py:38): Linear over channels.
    """
    xh = layer_norm(x, sd[pre + "0.norm.weight"], sd[pre + "0.norm.bias"])
    z1 = token_mix(sd[pre + "0.fn.net.0.weight"], xh, sd[pre + "0.fn.net.0.bias"])      # [B, Ds, C]
    u = x + token_mix(sd[pre + "0.fn.net.3.weight"], gelu(z1), sd[pre + "0.fn.net.3.bias"])
    uh = layer_norm(u, sd[pre + "1.norm.weight"], sd[pre + "1.norm.bias"])
    z2 = uh @ sd[pre + "1.fn.net.0.weight"].t() + sd[pre + "1.fn.net.0.bias"]
    return u + gelu(z2) @ sd[pre + "1.fn.net.3.weight"].t() + sd[pre + "1.fn.net.3.bias"]


def patchify(sd, x, key="patcher.0"):
    """Stem Conv2d(k = s = patch) then permute(0,2,3,1).view(B, -1, C) (mlp_mixer.py:58-60,68-71)."""
    w = sd[key + ".weight"]
    p = F.conv2d(x, w, sd[key + ".bias"], stride=w.shape[-1])
    return p.permute(0, 2, 3, 1).reshape(p.shape[0], -1, p.shape[1])


def mixer_forward(sd, x, depth):
    """MLPMixerForImageClassification.forward (mlp_mixer.py:67-76)."""
    t = patchify(sd, x)
    for i in range(depth):
        t = mixer_block(sd, f"model.{i}.", t)
    t = layer_norm(t, sd["active.weight"], sd["active.bias"])
    return t.mean(1) @ sd["mlp_head.0.weight"].t() + sd["mlp_head.0.bias"]


# ----------------------------------------------------------------------------------------------- ResMLP
def resmlp_block(sd, pre, x):
    """MLPblock.forward (res_mlp.py:52-57): the residual is taken AFTER the pre-affine."""
    a = x * sd[pre + "pre_affine.alpha"] + sd[pre + "pre_affine.beta"]
    t = a + sd[pre + "gamma_1"] * token_mix(sd[pre + "token_mix.weight"], a, sd[pre + "token_mix.bias"])
    u = t * sd[pre + "post_affine.alpha"] + sd[pre + "post_affine.beta"]
    z = u @ sd[pre + "ff.net.0.weight"].t() + sd[pre + "ff.net.0.bias"]
    return u + sd[pre + "gamma_2"] * (gelu(z) @ sd[pre + "ff.net.3.weight"].t() + sd[pre + "ff.net.3.bias"])


def resmlp_forward(sd, x, depth):
    """ResMLPForImageClassification.forward (res_mlp.py:91-99); `affine` is never applied (res_mlp.py:86)."""
    t = patchify(sd, x)
    for i in range(depth):
        t = resmlp_block(sd, f"model.{i}.", t)
    return t.mean(1) @ sd["mlp_head.0.weight"].t() + sd["mlp_head.0.bias"]


# ----------------------------------------------------------------------------------------------- gMLP
def gmlp_block(sd, pre, x):
    """gMLPBlock.forward with SpatialGatingUnit (g_mlp.py:17-22,32-39)."""
    z = gelu(layer_norm(x, sd[pre + "norm.weight"], sd[pre + "norm.bias"]) @ sd[pre + "channel_proj1.weight"].t()
             + sd[pre + "channel_proj1.bias"])
    u, v = z.chunk(2, dim=-1)
    v = layer_norm(v, sd[pre + "sgu.norm.weight"], sd[pre + "sgu.norm.bias"])
    v = token_mix(sd[pre + "sgu.spatial_proj.weight"], v, sd[pre + "sgu.spatial_proj.bias"])
    return (u * v) @ sd[pre + "channel_proj2.weight"].t() + sd[pre + "channel_proj2.bias"] + x


def gmlp_forward(sd, x, depth):
    """gMLPForImageClassification.forward (g_mlp.py:72-81)."""
    t = patchify(sd, x)
    for i in range(depth):
        t = gmlp_block(sd, f"model.{i}.", t)
    return t.mean(1) @ sd["mlp_head.0.weight"].t() + sd["mlp_head.0.bias"]


# ----------------------------------------------------------------------------------------------- metrics
def rel_l2(a, b):
    """||a - b|| / ||b|| in fp64 (the parity metric: SURVEY.md §4)."""
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def compare_py_metric(a, b):
    """The reference's own parity metric, mean(|(x+1)-(y+1)| / |y+1|) (compare.py:179-186)."""
    a, b = a.detach().double(), b.detach().double()
    return float(((a + 1) - (b + 1)).abs().div((b + 1).abs()).mean())
