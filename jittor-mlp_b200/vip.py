"""Vision Permutator (ViP) with the block bodies on the sm_100a path.

Drop-in for /root/reference/models_pytorch/vip.py (same classes, constructor signatures, defaults, state_dict keys):
the einops `Rearrange` layers of the reference carry no parameters, so `nn.Identity` place-holders keep the Sequential
indices (`fns.0.1.weight`, ...) and the rearrangements run inside fn_vip.VipBranchesFn as strided copies.
"""
from einops.layers.torch import Reduce
from torch import nn

from . import fn, fn_s2, fn_vip
from .s2_mlp import PreNormResidual, SplitAttention, _check_dropout, _ff
from .utils import pair


class ParallelSum(nn.Module):
    """Parameter container (vip.py:16-22)."""

    def __init__(self, *fns):
        super().__init__()
        self.fns = nn.ModuleList(fns)


class ParallelWeightedSum(nn.Module):
    """Parameter container (vip.py:24-35)."""

    def __init__(self, sa, *fns):
        super().__init__()
        self.fns = nn.ModuleList(fns)
        self.split_attention = sa


def _branches(height, width, d_model, segments):
    return (nn.Sequential(nn.Identity(), nn.Linear(height * segments, height * segments), nn.Identity()),
            nn.Sequential(nn.Identity(), nn.Linear(width * segments, width * segments), nn.Identity()),
            nn.Linear(d_model, d_model))


def _channel_mlp(d_model, expansion_factor, dropout):
    return nn.Sequential(nn.Linear(d_model, d_model * expansion_factor), nn.GELU(), nn.Dropout(dropout),
                         nn.Linear(d_model * expansion_factor, d_model), nn.Dropout(dropout))


class _PermutatorBase(nn.Module):
    weighted = True

    def __init__(self, height, width, d_model, depth, segments, expansion_factor=4, dropout=0.):
        super().__init__()
        _check_dropout(dropout)
        self.segments = segments

        def mixer():
            fns = _branches(height, width, d_model, segments)
            if self.weighted:
                return ParallelWeightedSum(SplitAttention(d_model, k=3), *fns)
            return ParallelSum(*fns)

        self.model = nn.Sequential(
            *[nn.Sequential(
                PreNormResidual(d_model, nn.Sequential(mixer(), nn.Linear(d_model, d_model))),
                PreNormResidual(d_model, _channel_mlp(d_model, expansion_factor, dropout))
            ) for _ in range(depth)])

    def forward(self, x):
        """x: [B, H, W, C] channels-last (vip.py:94-95 / 127-128)."""
        x = x.contiguous()
        for blk in self.model:
            a, b = blk[0], blk[1]
            par, proj = a.fn[0], a.fn[1]
            xn = fn.layer_norm(x, a.norm.weight, a.norm.bias, a.norm.eps)
            lh, lw, lc = par.fns[0][1], par.fns[1][1], par.fns[2]
            t = fn_vip.VipBranchesFn.apply(xn, lh.weight, lh.bias, lw.weight, lw.bias, lc.weight, lc.bias, self.segments,
                                           self.weighted)
            if self.weighted:
                sa = par.split_attention
                t = fn_s2.S2v2SplitAttentionFn.apply(t, sa.mlp1.weight, sa.mlp2.weight, 1)
            x = fn.linear(t, proj.weight, proj.bias, x)
            x = _ff(b.fn, fn.layer_norm(x, b.norm.weight, b.norm.bias, b.norm.eps), x)
        return x


class WeightedPermutator(_PermutatorBase):
    weighted = True


class Permutator(_PermutatorBase):
    weighted = False


class ViP(nn.Module):
    def __init__(self, image_size=224, patch_size=16, in_channels=3, num_classes=1000, d_model=256, depth=30,
                 segments=14, expansion_factor=4, weighted=True):
        image_size = pair(image_size)
        patch_size = pair(patch_size)
        assert (image_size[0] % patch_size[0]) == 0, 'image must be divisible by patch size'
        assert (image_size[1] % patch_size[1]) == 0, 'image must be divisible by patch size'
        assert (d_model % segments) == 0, 'dimension must be divisible by the number of segments'
        height = image_size[0] // patch_size[0]
        width = image_size[1] // patch_size[1]
        super().__init__()
        self.patcher = nn.Sequential(nn.Conv2d(in_channels, d_model, kernel_size=patch_size, stride=patch_size))
        cls = WeightedPermutator if weighted else Permutator
        self.blocks = cls(height, width, d_model, depth, segments, expansion_factor, dropout=0.)
        self.mlp_head = nn.Sequential(nn.LayerNorm(d_model), Reduce('b h w c -> b c', 'mean'),
                                      nn.Linear(d_model, num_classes))

    def forward(self, x):
        import torch
        patches = self.patcher[0](x.contiguous(memory_format=torch.channels_last))   # cuDNN NHWC; the permute is a view
        emb = self.blocks(patches.permute(0, 2, 3, 1))
        ln = self.mlp_head[0]
        emb = fn.layer_norm(emb, ln.weight, ln.bias, ln.eps)
        return fn.head(emb, self.mlp_head[2])
