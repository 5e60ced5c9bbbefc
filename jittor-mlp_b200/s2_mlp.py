"""S2-MLP v1 and v2 with the block bodies on the sm_100a path.

Drop-ins for /root/reference/models_pytorch/s2_mlp_v1.py and s2_mlp_v2.py (same classes, constructor signatures,
defaults, state_dict keys).  The spatial shifts implement the INTENDED clamp-to-edge semantics: the reference's
in-place overlapping slice copies (s2_mlp_v1.py:21-24, s2_mlp_v2.py:17-28) are undefined behaviour and not
repeatable; its autograd backward and its Jittor twin both define the clone semantics used here.
"""
from einops.layers.torch import Reduce
import torch
from torch import nn

from . import fn, fn_s2
from .utils import pair


class PreNormResidual(nn.Module):
    """Parameter container (s2_mlp_v1.py:6-13 / s2_mlp_v2.py:6-13)."""

    def __init__(self, dim, fn_):
        super().__init__()
        self.fn = fn_
        self.norm = nn.LayerNorm(dim)


class Spatial_Shift(nn.Module):
    def forward(self, x):
        return fn.s2_shift(x, 1)


def _ff(seq, x, res):
    """Sequential(Linear, GELU, Dropout, Linear, Dropout) (s2_mlp_v1.py:40-45)."""
    return fn.mlp(x, seq[0].weight, seq[0].bias, seq[3].weight, seq[3].bias, res)


def _check_dropout(p):
    if p != 0.:
        raise ValueError("the fused blocks implement dropout = 0 only (the reference default)")


# ------------------------------------------------------------------------------------------------- v1
class S2BlockV1(nn.Module):
    def __init__(self, d_model, depth, expansion_factor=4, dropout=0.):
        super().__init__()
        _check_dropout(dropout)
        self.model = nn.Sequential(
            *[nn.Sequential(
                PreNormResidual(d_model, nn.Sequential(
                    nn.Linear(d_model, d_model), nn.GELU(), Spatial_Shift(), nn.Linear(d_model, d_model))),
                PreNormResidual(d_model, nn.Sequential(
                    nn.Linear(d_model, d_model * expansion_factor), nn.GELU(), nn.Dropout(dropout),
                    nn.Linear(d_model * expansion_factor, d_model), nn.Dropout(dropout)))
            ) for _ in range(depth)])

    def forward(self, x):
        x = x.permute(0, 2, 3, 1).contiguous()                         # NCHW -> channels-last rows
        for blk in self.model:
            a, b = blk[0], blk[1]
            t = fn.linear_gelu(fn.layer_norm(x, a.norm.weight, a.norm.bias, a.norm.eps), a.fn[0].weight, a.fn[0].bias)
            x = fn.linear(fn.s2_shift(t, 1), a.fn[3].weight, a.fn[3].bias, x)
            x = _ff(b.fn, fn.layer_norm(x, b.norm.weight, b.norm.bias, b.norm.eps), x)
        return x.permute(0, 3, 1, 2)


class S2MLPv1(nn.Module):
    def __init__(self, image_size=224, patch_size=[7, 2], in_channels=3, num_classes=1000, d_model=[192, 384],
                 depth=[4, 14], expansion_factor=[3, 3]):
        image_size = pair(image_size)
        oldps = [1, 1]
        for ps in patch_size:
            ps = pair(ps)
            assert (image_size[0] % (ps[0] * oldps[0])) == 0, 'image must be divisible by patch size'
            assert (image_size[1] % (ps[1] * oldps[1])) == 0, 'image must be divisible by patch size'
            oldps[0] = oldps[0] * ps[0]
            oldps[1] = oldps[1] * ps[1]
        assert (len(patch_size) == len(depth) == len(d_model) == len(expansion_factor)), \
            'patch_size/depth/d_model/expansion_factor must be a list'
        super().__init__()
        self.stage = len(patch_size)
        self.stages = nn.Sequential(
            *[nn.Sequential(
                nn.Conv2d(in_channels if i == 0 else d_model[i - 1], d_model[i], kernel_size=patch_size[i],
                          stride=patch_size[i]),
                S2BlockV1(d_model[i], depth[i], expansion_factor[i], dropout=0.)
            ) for i in range(self.stage)])
        self.mlp_head = nn.Sequential(Reduce('b c h w -> b c', 'mean'), nn.Linear(d_model[-1], num_classes))

    def forward(self, x):
        # channels-last input: the stage convs (cuDNN) run their NHWC kernels and the block-side permutes are views
        t = self.stages(x.contiguous(memory_format=torch.channels_last))      # NCHW view of channels-last rows
        return fn.head(t.permute(0, 2, 3, 1), self.mlp_head[1])


def S2MLPv1_deep(num_classes: int = 1000, **kwargs):
    return S2MLPv1(image_size=224, patch_size=[16], d_model=[384], depth=[36], num_classes=num_classes,
                   expansion_factor=[4], **kwargs)


def S2MLPv1_wide(num_classes: int = 1000, **kwargs):
    return S2MLPv1(image_size=224, patch_size=[16], d_model=[768], depth=[12], num_classes=num_classes,
                   expansion_factor=[4], **kwargs)


# ------------------------------------------------------------------------------------------------- v2
class SplitAttention(nn.Module):
    """Parameter container (s2_mlp_v2.py:31-39)."""

    def __init__(self, channel=512, k=3):
        super().__init__()
        if k != 3:
            raise ValueError("split attention is implemented for k = 3 (the only value the reference uses)")
        self.channel = channel
        self.k = k
        self.mlp1 = nn.Linear(channel, channel, bias=False)
        self.gelu = nn.GELU()
        self.mlp2 = nn.Linear(channel, channel * k, bias=False)
        self.softmax = nn.Softmax(1)


class S2Attention(nn.Module):
    def __init__(self, channels=512):
        super().__init__()
        self.mlp1 = nn.Linear(channels, channels * 3)
        self.mlp2 = nn.Linear(channels, channels)
        self.split_attention = SplitAttention(channels)

    def attend(self, xn, res):
        """mlp2(split_attention(shifts(mlp1(xn)))) + res  (s2_mlp_v2.py:60-69)."""
        sa = self.split_attention
        t = fn.linear(xn, self.mlp1.weight, self.mlp1.bias)                       # [B, H, W, 3C]
        s = fn_s2.S2v2SplitAttentionFn.apply(t, sa.mlp1.weight, sa.mlp2.weight)    # sum -> tiny MLP -> softmax-combine
        return fn.linear(s, self.mlp2.weight, self.mlp2.bias, res)


class S2BlockV2(nn.Module):
    def __init__(self, d_model, depth, expansion_factor=4, dropout=0.):
        super().__init__()
        _check_dropout(dropout)
        self.model = nn.Sequential(
            *[nn.Sequential(
                PreNormResidual(d_model, S2Attention(d_model)),
                PreNormResidual(d_model, nn.Sequential(
                    nn.Linear(d_model, d_model * expansion_factor), nn.GELU(), nn.Dropout(dropout),
                    nn.Linear(d_model * expansion_factor, d_model), nn.Dropout(dropout)))
            ) for _ in range(depth)])

    def forward(self, x):
        x = x.permute(0, 2, 3, 1).contiguous()
        for blk in self.model:
            a, b = blk[0], blk[1]
            x = a.fn.attend(fn.layer_norm(x, a.norm.weight, a.norm.bias, a.norm.eps), x)
            x = _ff(b.fn, fn.layer_norm(x, b.norm.weight, b.norm.bias, b.norm.eps), x)
        return x.permute(0, 3, 1, 2)


class S2MLPv2(nn.Module):
    def __init__(self, image_size=224, patch_size=[7, 2], in_channels=3, num_classes=1000, d_model=[192, 384],
                 depth=[4, 14], expansion_factor=[3, 3]):
        image_size = pair(image_size)
        oldps = [1, 1]
        for ps in patch_size:
            ps = pair(ps)
            assert (image_size[0] % (ps[0] * oldps[0])) == 0, 'image must be divisible by patch size'
            assert (image_size[1] % (ps[1] * oldps[1])) == 0, 'image must be divisible by patch size'
            oldps[0] = oldps[0] * ps[0]
            oldps[1] = oldps[1] * ps[1]
        assert (len(patch_size) == len(depth) == len(d_model) == len(expansion_factor)), \
            'patch_size/depth/d_model/expansion_factor must be a list'
        super().__init__()
        self.stage = len(patch_size)
        self.stages = nn.Sequential(
            *[nn.Sequential(
                nn.Conv2d(in_channels if i == 0 else d_model[i - 1], d_model[i], kernel_size=patch_size[i],
                          stride=patch_size[i]),
                S2BlockV2(d_model[i], depth[i], expansion_factor[i], dropout=0.)
            ) for i in range(self.stage)])
        self.mlp_head = nn.Sequential(Reduce('b c h w -> b c', 'mean'), nn.Linear(d_model[-1], num_classes))

    def forward(self, x):
        # channels-last input: the stage convs (cuDNN) run their NHWC kernels and the block-side permutes are views
        t = self.stages(x.contiguous(memory_format=torch.channels_last))      # NCHW view of channels-last rows
        return fn.head(t.permute(0, 2, 3, 1), self.mlp_head[1])
