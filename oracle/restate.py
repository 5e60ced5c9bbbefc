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


# ----------------------------------------------------------------------------------------------- shift family helpers
def linear(x, w, b=None):
    """x [..., Cin] @ W^T (+ b); W may be a Linear weight [out, in] or a 1x1 Conv2d weight [out, in, 1, 1]."""
    y = x @ w.reshape(w.shape[0], -1).t()
    return y if b is None else y + b


def shift_tokens(x, groups, clamp):
    """x [B, H, W, C]; groups = [(c_lo, c_hi, dh, dw)]: out[h, w] = x[h + dh, w + dw] per channel group, zero outside
    (AS-MLP Shift kernel, shift_cuda.py:44-72) or clamped to the edge (S2-MLP intended semantics).
    Written with explicit index tensors (no roll / no in-place slice copies)."""
    B, H, W, C = x.shape
    out = torch.zeros_like(x)
    hh = torch.arange(H)
    ww = torch.arange(W)
    for lo, hi, dh, dw in groups:
        hs, ws = hh + dh, ww + dw
        if clamp:
            g = x[:, hs.clamp(0, H - 1)][:, :, ws.clamp(0, W - 1)][..., lo:hi]
        else:
            g = x[:, hs.clamp(0, H - 1)][:, :, ws.clamp(0, W - 1)][..., lo:hi]
            ok = ((hs >= 0) & (hs < H))[:, None] & ((ws >= 0) & (ws < W))[None, :]
            g = g * ok[None, :, :, None].to(x.dtype)
        out = torch.cat([out[..., :lo], g, out[..., hi:]], -1)
    return out


def s2_groups(C, plan):
    """Channel quarters and offsets of spatial_shift1 / spatial_shift2 (s2_mlp_v2.py:15-29):
    `x[:,1:] = x[:,:-1]` is out[i] = in[i-1]."""
    q = [0, C // 4, C // 2, C * 3 // 4, C]
    offs = [(-1, 0), (1, 0), (0, -1), (0, 1)] if plan == 1 else [(0, -1), (0, 1), (-1, 0), (1, 0)]
    return [(q[i], q[i + 1], offs[i][0], offs[i][1]) for i in range(4)]


def as_groups(C, S, dim):
    """Shift(kernel_size=S, dim): group g = c // ceil(C/S) reads position + (S//2 - g) (shift_cuda.py:44-72)."""
    cs = -(-C // S)
    out = []
    for g in range(S):
        if g * cs >= C:
            break
        s = S // 2 - g
        out.append((g * cs, min((g + 1) * cs, C), s, 0) if dim == 2 else (g * cs, min((g + 1) * cs, C), 0, s))
    return out


def group_norm1(x, w, b, eps=1e-5):
    """nn.GroupNorm(1, C) on channels-last x [B, H, W, C]: statistics over the whole sample (as_mlp.py:343-344)."""
    mu = x.mean((1, 2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean((1, 2, 3), keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


# ----------------------------------------------------------------------------------------------- S2-MLP v1 / v2
def s2v1_forward(sd, x, depths, patch_sizes):
    """S2MLPv1.forward (s2_mlp_v1.py:88-93, block :32-46), clamp-shift semantics."""
    t = x
    for s, (depth, ps) in enumerate(zip(depths, patch_sizes)):
        t = F.conv2d(t, sd[f"stages.{s}.0.weight"], sd[f"stages.{s}.0.bias"], stride=ps).permute(0, 2, 3, 1)
        C = t.shape[-1]
        for i in range(depth):
            p = f"stages.{s}.1.model.{i}."
            h = gelu(linear(layer_norm(t, sd[p + "0.norm.weight"], sd[p + "0.norm.bias"]), sd[p + "0.fn.0.weight"], sd[p + "0.fn.0.bias"]))
            t = t + linear(shift_tokens(h, s2_groups(C, 1), True), sd[p + "0.fn.3.weight"], sd[p + "0.fn.3.bias"])
            h = gelu(linear(layer_norm(t, sd[p + "1.norm.weight"], sd[p + "1.norm.bias"]), sd[p + "1.fn.0.weight"], sd[p + "1.fn.0.bias"]))
            t = t + linear(h, sd[p + "1.fn.3.weight"], sd[p + "1.fn.3.bias"])
        t = t.permute(0, 3, 1, 2)
    return linear(t.mean((2, 3)), sd["mlp_head.1.weight"], sd["mlp_head.1.bias"])


def s2v2_forward(sd, x, depths, patch_sizes):
    """S2MLPv2.forward (s2_mlp_v2.py:129-132; S2Attention :60-69; SplitAttention :41-51)."""
    t = x
    for s, (depth, ps) in enumerate(zip(depths, patch_sizes)):
        t = F.conv2d(t, sd[f"stages.{s}.0.weight"], sd[f"stages.{s}.0.bias"], stride=ps).permute(0, 2, 3, 1)
        C = t.shape[-1]
        for i in range(depth):
            p = f"stages.{s}.1.model.{i}."
            u = linear(layer_norm(t, sd[p + "0.norm.weight"], sd[p + "0.norm.bias"]), sd[p + "0.fn.mlp1.weight"], sd[p + "0.fn.mlp1.bias"])
            xs = [shift_tokens(u[..., :C], s2_groups(C, 1), True), shift_tokens(u[..., C:2 * C], s2_groups(C, 2), True), u[..., 2 * C:]]
            a = (xs[0] + xs[1] + xs[2]).sum((1, 2))                                   # [B, C]
            hat = linear(gelu(linear(a, sd[p + "0.fn.split_attention.mlp1.weight"])), sd[p + "0.fn.split_attention.mlp2.weight"])
            bar = torch.softmax(hat.reshape(-1, 3, C), 1)                             # [B, 3, C]
            o = sum(bar[:, k, None, None, :] * xs[k] for k in range(3))
            t = t + linear(o, sd[p + "0.fn.mlp2.weight"], sd[p + "0.fn.mlp2.bias"])
            h = gelu(linear(layer_norm(t, sd[p + "1.norm.weight"], sd[p + "1.norm.bias"]), sd[p + "1.fn.0.weight"], sd[p + "1.fn.0.bias"]))
            t = t + linear(h, sd[p + "1.fn.3.weight"], sd[p + "1.fn.3.bias"])
        t = t.permute(0, 3, 1, 2)
    return linear(t.mean((2, 3)), sd["mlp_head.1.weight"], sd["mlp_head.1.bias"])


# ----------------------------------------------------------------------------------------------- AS-MLP
def asmlp_block(sd, p, t, S):
    """AxialShiftedBlock.forward (as_mlp.py:149-162) with AxialShift (as_mlp.py:55-95) on channels-last t; DropPath = identity."""
    C = t.shape[-1]
    n = group_norm1(t, sd[p + "norm1.weight"], sd[p + "norm1.bias"])
    a = p + "axial_shift."
    v = gelu(group_norm1(linear(n, sd[a + "conv1.weight"], sd.get(a + "conv1.bias")), sd[a + "norm1.weight"], sd[a + "norm1.bias"]))
    lr = gelu(linear(shift_tokens(v, as_groups(C, S, 3), False), sd[a + "conv2_1.weight"], sd.get(a + "conv2_1.bias")))
    td = gelu(linear(shift_tokens(v, as_groups(C, S, 2), False), sd[a + "conv2_2.weight"], sd.get(a + "conv2_2.bias")))
    t = t + linear(group_norm1(lr + td, sd[a + "norm2.weight"], sd[a + "norm2.bias"]), sd[a + "conv3.weight"], sd.get(a + "conv3.bias"))
    n = group_norm1(t, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    return t + linear(gelu(linear(n, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


def asmlp_forward(sd, x, depths, patch_size=4, shift_size=5):
    """AS_MLP.forward (as_mlp.py:428-443) with drop_path = 0 / eval."""
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch_size).permute(0, 2, 3, 1)
    if "patch_embed.norm.weight" in sd:
        t = group_norm1(t, sd["patch_embed.norm.weight"], sd["patch_embed.norm.bias"])
    for li, depth in enumerate(depths):
        for i in range(depth):
            t = asmlp_block(sd, f"layers.{li}.blocks.{i}.", t, shift_size)
        if li < len(depths) - 1:                                    # PatchMerging (as_mlp.py:197-216)
            t = torch.cat([t[:, 0::2, 0::2], t[:, 1::2, 0::2], t[:, 0::2, 1::2], t[:, 1::2, 1::2]], -1)
            d = f"layers.{li}.downsample."
            t = linear(group_norm1(t, sd[d + "norm.weight"], sd[d + "norm.bias"]), sd[d + "reduction.weight"])
    t = group_norm1(t, sd["norm.weight"], sd["norm.bias"])
    return linear(t.mean((1, 2)), sd["head.weight"], sd["head.bias"])


# ----------------------------------------------------------------------------------------------- Hire-MLP
def hire_block(sd, p, t, h, w, step):
    """HireMLPBlock.forward (hire_mlp.py:130-152) on channels-last t = LN(x) [B, H, W, C], index formulation of
    SURVEY.md Appendix A6: circular pad (a full extra region when divisible), roll, strided region gather, bottleneck
    1x1-conv MLP, inverse gather, roll back, crop."""
    B, H, W, C = t.shape
    Hp, Wp = H + (h - H % h), W + (w - W % w)
    xp = t[:, torch.arange(Hp) % H][:, :, torch.arange(Wp) % W]                 # circular pad right / bottom

    def branch(xp, n, L, axis, key):
        G = L // n
        src = (torch.arange(L) - step) % L                                       # roll(+step): rolled[r] = x[r - step]
        xr = xp.index_select(axis, src)
        # region gather: position r = i*G + g -> feature index (c, i) at spatial g
        shape = list(xr.shape)
        xr = xr.reshape(shape[:axis] + [n, G] + shape[axis + 1:])               # [.., i, g, ..]
        xr = xr.movedim(axis, -1)                                                # [..., g, .., C, i]
        z = xr.reshape(list(xr.shape[:-2]) + [C * n])                            # feature index c*n + i
        hdn = gelu(linear(z, sd[key + "net.0.weight"], sd[key + "net.0.bias"]))
        o = linear(hdn, sd[key + "net.2.weight"], sd[key + "net.2.bias"])
        o = o.reshape(list(o.shape[:-1]) + [C, n]).movedim(-1, axis)             # back to [.., i, g, .., C]
        shape2 = list(o.shape)
        o = o.reshape(shape2[:axis] + [L] + shape2[axis + 2:])
        return o.index_select(axis, (torch.arange(L) + step) % L)                # roll(-step)

    xh = branch(xp, h, Hp, 1, p + "proj_h.")
    xw = branch(xp, w, Wp, 2, p + "proj_w.")
    xc = linear(xp, sd[p + "proj_c.weight"], sd[p + "proj_c.bias"])
    return (xc + xh + xw)[:, :H, :W]


def hire_forward(sd, x, kw):
    """HireMLP.forward (hire_mlp.py:223-229); kw = constructor kwargs."""
    d_model = kw.get("d_model", [64, 128, 320, 512]); hs = kw.get("h", [4, 3, 3, 2]); ws = kw.get("w", [4, 3, 3, 2])
    steps = kw.get("cross_region_step", [2, 2, 1, 1]); interval = kw.get("cross_region_interval", 2)
    depth = kw.get("depth", [4, 6, 24, 3]); ps = kw.get("patch_size", 4)
    t = F.conv2d(x, sd["patcher.reduction.0.weight"], sd["patcher.reduction.0.bias"], stride=ps, padding=3)
    t = t.permute(0, 2, 3, 1)
    if "patcher.reduction.1.1.weight" in sd:
        t = layer_norm(t, sd["patcher.reduction.1.1.weight"], sd["patcher.reduction.1.1.bias"])
    for s in range(len(depth)):
        for i in range(depth[s]):
            p = f"layers.{s}.model.{i}."
            step = steps[s] if ((i + 1) % interval == 0) else 0
            n = layer_norm(t, sd[p + "0.norm.weight"], sd[p + "0.norm.bias"])
            t = t + hire_block(sd, p + "0.fn.0.", n, hs[s], ws[s], step)
            n = layer_norm(t, sd[p + "1.norm.weight"], sd[p + "1.norm.bias"])
            t = t + linear(gelu(linear(n, sd[p + "1.fn.0.weight"], sd[p + "1.fn.0.bias"])), sd[p + "1.fn.3.weight"], sd[p + "1.fn.3.bias"])
        if s + 1 < len(depth):
            m = f"layers.{s}.patch_merge.1.reduction.0."
            t = F.conv2d(t.permute(0, 3, 1, 2), sd[m + "weight"], sd[m + "bias"], stride=2, padding=1).permute(0, 2, 3, 1)
    t = layer_norm(t, sd["mlp_head.0.weight"], sd["mlp_head.0.bias"])
    return linear(t.mean((1, 2)), sd["mlp_head.2.weight"], sd["mlp_head.2.bias"])


# ----------------------------------------------------------------------------------------------- ConvMixer
def batch_norm_train(x, w, b, eps=1e-5):
    """nn.BatchNorm2d in train(): biased batch variance over (B, H, W) per channel (conv_mixer.py:20)."""
    mu = x.mean((0, 2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean((0, 2, 3), keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w[None, :, None, None] + b[None, :, None, None]


def convmixer_forward(sd, x, kw):
    """ConvMixer.forward in train() (conv_mixer.py:41-45); NCHW, batch statistics."""
    depth, k, ps = kw["depth"], kw.get("kernel_size", 9), kw.get("patch_size", 7)
    t = F.conv2d(x, sd["embedding.0.weight"], sd["embedding.0.bias"], stride=ps, padding=ps // 2)
    t = batch_norm_train(gelu(t), sd["embedding.2.weight"], sd["embedding.2.bias"])
    for i in range(depth):
        p = f"blocks.{i}."
        dw = F.conv2d(t, sd[p + "0.fn.0.weight"], sd[p + "0.fn.0.bias"], padding=k // 2, groups=t.shape[1])
        t = batch_norm_train(gelu(dw), sd[p + "0.fn.2.weight"], sd[p + "0.fn.2.bias"]) + t
        pw = F.conv2d(t, sd[p + "1.weight"], sd[p + "1.bias"])
        t = batch_norm_train(gelu(pw), sd[p + "3.weight"], sd[p + "3.bias"])
    return linear(t.mean((2, 3)), sd["classifier.2.weight"], sd["classifier.2.bias"])


# ----------------------------------------------------------------------------------------------- ViP
def vip_forward(sd, x, kw):
    """ViP.forward (vip.py:166-171); WeightedPermutator / Permutator blocks (vip.py:59-128), ParallelWeightedSum
    (vip.py:24-35), SplitAttention (vip.py:37-57).  The einops rearrangements are written as explicit reshapes:
    `b h w (c s) -> b w c (h s)` = [B,H,W,c,S] -> permute(0,2,3,1,4) -> [B,W,c,H*S]."""
    ps = kw.get("patch_size", 16)
    ps = (ps, ps) if isinstance(ps, int) else tuple(ps)
    S, depth, weighted = kw.get("segments", 14), kw.get("depth", 30), kw.get("weighted", True)
    t = F.conv2d(x, sd["patcher.0.weight"], sd["patcher.0.bias"], stride=ps).permute(0, 2, 3, 1)
    B, H, W, C = t.shape
    c = C // S
    for i in range(depth):
        p = f"blocks.model.{i}."
        q = p + "0.fn.0."
        u = layer_norm(t, sd[p + "0.norm.weight"], sd[p + "0.norm.bias"])
        u5 = u.reshape(B, H, W, c, S)
        xh = linear(u5.permute(0, 2, 3, 1, 4).reshape(B, W, c, H * S), sd[q + "fns.0.1.weight"], sd[q + "fns.0.1.bias"])
        xh = xh.reshape(B, W, c, H, S).permute(0, 3, 1, 2, 4).reshape(B, H, W, C)
        xw = linear(u5.permute(0, 1, 3, 2, 4).reshape(B, H, c, W * S), sd[q + "fns.1.1.weight"], sd[q + "fns.1.1.bias"])
        xw = xw.reshape(B, H, c, W, S).permute(0, 1, 3, 2, 4).reshape(B, H, W, C)
        xc = linear(u, sd[q + "fns.2.weight"], sd[q + "fns.2.bias"])
        if weighted:
            a = (xh + xw + xc).sum((1, 2))                                            # [B, C]
            hat = linear(gelu(linear(a, sd[q + "split_attention.mlp1.weight"])), sd[q + "split_attention.mlp2.weight"])
            bar = torch.softmax(hat.reshape(B, 3, C), 1)
            o = bar[:, 0, None, None, :] * xh + bar[:, 1, None, None, :] * xw + bar[:, 2, None, None, :] * xc
        else:
            o = xh + xw + xc
        t = t + linear(o, sd[p + "0.fn.1.weight"], sd[p + "0.fn.1.bias"])
        h = gelu(linear(layer_norm(t, sd[p + "1.norm.weight"], sd[p + "1.norm.bias"]), sd[p + "1.fn.0.weight"], sd[p + "1.fn.0.bias"]))
        t = t + linear(h, sd[p + "1.fn.3.weight"], sd[p + "1.fn.3.bias"])
    t = layer_norm(t, sd["mlp_head.0.weight"], sd["mlp_head.0.bias"])
    return linear(t.mean((1, 2)), sd["mlp_head.2.weight"], sd["mlp_head.2.bias"])


# ----------------------------------------------------------------------------------------------- optimizer step (f4)
def adamw_step(w, g, m, v, t, lr, b1, b2, eps, wd):
    """One torch.optim.AdamW step on fp32 tensors (decoupled weight decay, bias-corrected moments); t = step count
    AFTER this step.  The reference has no optimizer (compare.py:141-145): this restates the library's documented
    algorithm, and tests pin it against torch.optim.AdamW itself."""
    w = w * (1 - lr * wd)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    w = w - (lr / (1 - b1 ** t)) * m / (v.sqrt() / (1 - b2 ** t) ** 0.5 + eps)
    return w, m, v


def sgd_step(w, g, buf, t, lr, momentum, wd):
    """One torch.optim.SGD(momentum, dampening=0) step; the buffer starts as the first gradient."""
    g = g + wd * w
    buf = g.clone() if t == 1 else momentum * buf + g
    return w - lr * buf, buf


# ----------------------------------------------------------------------------------------------- SparseMLP (row f3)
def batch_norm_train_nchw_rows(x, w, b, eps=1e-5):
    """nn.BatchNorm2d in train() on channels-last rows [..., C]: batch statistics over every leading axis (biased
    variance), sparse_mlp.py:91,96 (norm = nn.BatchNorm2d)."""
    flat = x.reshape(-1, x.shape[-1])
    mean = flat.mean(0)
    var = ((flat - mean) ** 2).mean(0)
    return (x - mean) / torch.sqrt(var + eps) * w + b


def sparsemlp_forward(sd, x, kw):
    """SparseMLP.forward (sparse_mlp.py:158-165); sMLPStage :77-115; sMLPBlock :60-75; PatchMerging :17-50.  Channels-last
    restatement: proj_h contracts the H axis, proj_w the W axis, `fuse` is a 1x1 conv over the concatenated channels."""
    ps = kw.get("patch_size", 4)
    ps = (ps, ps) if isinstance(ps, int) else tuple(ps)
    depth = kw.get("depth", [2, 10, 24, 2])
    t = F.conv2d(x, sd["patcher.0.weight"], sd["patcher.0.bias"], stride=ps).permute(0, 2, 3, 1)
    if kw.get("patcher_norm", False):
        t = layer_norm(t, sd["patcher.1.1.weight"], sd["patcher.1.1.bias"])
    for s, d in enumerate(depth):
        for i in range(d):
            p = f"layers.{s}.model.{i}."
            C = t.shape[-1]
            u = batch_norm_train_nchw_rows(t, sd[p + "0.norm.weight"], sd[p + "0.norm.bias"])
            u = F.conv2d(u.permute(0, 3, 1, 2), sd[p + "0.fn.0.weight"], sd[p + "0.fn.0.bias"], padding=1, groups=C).permute(0, 2, 3, 1)
            t = t + u
            u = batch_norm_train_nchw_rows(t, sd[p + "1.norm.weight"], sd[p + "1.norm.bias"])
            q = p + "1.fn.0."
            xh = torch.einsum("gh,bhwc->bgwc", sd[q + "proj_h.weight"], u) + sd[q + "proj_h.bias"][None, :, None, None]
            xw = torch.einsum("vw,bhwc->bhvc", sd[q + "proj_w.weight"], u) + sd[q + "proj_w.bias"][None, None, :, None]
            cat = torch.cat([xh, xw, u], -1)
            t = t + linear(cat, sd[q + "fuse.weight"].reshape(C, 3 * C), sd[q + "fuse.bias"])
            h = gelu(linear(layer_norm(t, sd[p + "3.norm.weight"], sd[p + "3.norm.bias"]), sd[p + "3.fn.0.weight"], sd[p + "3.fn.0.bias"]))
            t = t + linear(h, sd[p + "3.fn.3.weight"], sd[p + "3.fn.3.bias"])
        if s + 1 < len(depth):
            q = f"layers.{s}.patch_merge.1."
            t = torch.cat([t[:, 0::2, 0::2], t[:, 1::2, 0::2], t[:, 0::2, 1::2], t[:, 1::2, 1::2]], -1)
            t = linear(layer_norm(t, sd[q + "norm.weight"], sd[q + "norm.bias"]), sd[q + "reduction.weight"])
    t = layer_norm(t, sd["mlp_head.1.weight"], sd["mlp_head.1.bias"])
    return linear(t.mean((1, 2)), sd["mlp_head.3.weight"], sd["mlp_head.3.bias"])
