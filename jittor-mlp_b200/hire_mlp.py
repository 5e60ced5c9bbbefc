"""Hire-MLP with the block bodies on the sm_100a path.

Drop-in for /root/reference/models_pytorch/hire_mlp.py (same classes, constructor signatures, defaults, state_dict
keys).  Quirks preserved: a FULL extra region is padded when H is already divisible (hire_mlp.py:134-136),
the region gather is strided (hire_mlp.py:58,69), the last stage builds an unused ``patch_merge`` (hire_mlp.py:159-163).
"""
from einops.layers.torch import Rearrange, Reduce
import torch
from torch import nn

from . import fn, fn_spatial
from .utils import pair


class PreNormResidual(nn.Module):
    def __init__(self, dim, fn_, norm=nn.LayerNorm):
        super().__init__()
        self.fn = fn_
        self.norm = norm(dim)


class PatchEmbedding(nn.Module):
    """Stem / stage-transition conv (hire_mlp.py:17-31); stays on cuDNN."""

    def __init__(self, dim_in, dim_out, kernel_size, stride, padding, norm_layer=False):
        super().__init__()
        self.reduction = nn.Sequential(
            nn.Conv2d(dim_in, dim_out, kernel_size=kernel_size, stride=stride, padding=padding),
            nn.Identity() if (not norm_layer) else nn.Sequential(
                Rearrange('b c h w -> b h w c'), nn.LayerNorm(dim_out), Rearrange('b h w c -> b c h w')))

    def forward(self, x):
        return self.reduction(x)


class FeedForward(nn.Module):
    """Parameter container (hire_mlp.py:33-42): Conv2d 1x1 -> GELU -> Conv2d 1x1."""

    def __init__(self, dim_in, hidden_dim, dim_out):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(dim_in, hidden_dim, kernel_size=1), nn.GELU(),
                                 nn.Conv2d(hidden_dim, dim_out, kernel_size=1))

    def run_regions(self, z, n, C):
        """The bottleneck MLP on gathered rows whose feature axis is ordered [region index i][channel c]; the reference
        orders it (c i) (hire_mlp.py:58,69), so the 1x1-conv weights are permuted (tiny, differentiable)."""
        w1, b1, w2, b2 = self.net[0].weight, self.net[0].bias, self.net[2].weight, self.net[2].bias
        hid = w1.shape[0]
        w1p = w1.reshape(hid, C, n).permute(0, 2, 1).reshape(hid, n * C)
        w2p = w2.reshape(C, n, hid).permute(1, 0, 2).reshape(n * C, hid)
        b2p = b2.reshape(C, n).t().reshape(n * C)
        return fn.mlp(z, w1p.contiguous(), b1, w2p.contiguous(), b2p.contiguous(), None)


class HireMLPBlock(nn.Module):
    def __init__(self, h, w, d_model, cross_region_step=1, cross_region_id=0, cross_region_interval=2,
                 padding_type='circular'):
        super().__init__()
        assert (padding_type in ['constant', 'reflect', 'replicate', 'circular'])
        if padding_type != 'circular':
            raise ValueError("only padding_type='circular' (the reference default) is implemented")
        self.padding_type = padding_type
        self.w = w
        self.h = h
        self.cross_region = (cross_region_id % cross_region_interval == 0)
        self.step = cross_region_step if self.cross_region else 0
        self.proj_h = FeedForward(h * d_model, d_model // 2, h * d_model)
        self.proj_w = FeedForward(w * d_model, d_model // 2, w * d_model)
        self.proj_c = nn.Conv2d(d_model, d_model, kernel_size=1)

    def run(self, xn, res):
        """proj_c(x) + restore(proj_h(regions_H(x))) + restore(proj_w(regions_W(x))) + res  (hire_mlp.py:130-152)."""
        C = xn.shape[-1]
        zh, zw = fn_spatial.HireBuildFn.apply(xn, self.h, self.w, self.step, self.step)
        oh = self.proj_h.run_regions(zh, self.h, C)
        ow = self.proj_w.run_regions(zw, self.w, C)
        base = fn.linear(xn, self.proj_c.weight, self.proj_c.bias, res)
        return fn_spatial.HireCombineFn.apply(base, oh, ow, self.h, self.w, self.step, self.step)


class HireMLPStage(nn.Module):
    def __init__(self, h, w, d_model_in, d_model_out, depth, cross_region_step, cross_region_interval, expansion_factor=2,
                 dropout=0., pooling=False, padding_type='circular'):
        super().__init__()
        if dropout != 0.:
            raise ValueError("the fused blocks implement dropout = 0 only (the reference default)")
        self.pooling = pooling
        self.patch_merge = nn.Sequential(
            Rearrange('b h w c -> b c h w'),
            PatchEmbedding(d_model_in, d_model_out, kernel_size=3, stride=2, padding=1, norm_layer=False),
            Rearrange('b c h w -> b h w c'))
        self.model = nn.Sequential(
            *[nn.Sequential(
                PreNormResidual(d_model_in, nn.Sequential(
                    HireMLPBlock(h, w, d_model_in, cross_region_step=cross_region_step, cross_region_id=i_depth + 1,
                                 cross_region_interval=cross_region_interval, padding_type=padding_type)),
                    norm=nn.LayerNorm),
                PreNormResidual(d_model_in, nn.Sequential(
                    nn.Linear(d_model_in, d_model_in * expansion_factor), nn.GELU(), nn.Dropout(dropout),
                    nn.Linear(d_model_in * expansion_factor, d_model_in), nn.Dropout(dropout)), norm=nn.LayerNorm),
            ) for i_depth in range(depth)])

    def forward(self, x):                                   # [B, H, W, C]
        x = x.contiguous()
        for blk in self.model:
            a, b = blk[0], blk[1]
            x = a.fn[0].run(fn.layer_norm(x, a.norm.weight, a.norm.bias, a.norm.eps), x)
            x = fn.mlp(fn.layer_norm(x, b.norm.weight, b.norm.bias, b.norm.eps), b.fn[0].weight, b.fn[0].bias,
                       b.fn[3].weight, b.fn[3].bias, x)
        if self.pooling:
            x = self.patch_merge(x)
        return x


class HireMLP(nn.Module):
    def __init__(self, patch_size=4, in_channels=3, num_classes=1000, d_model=[64, 128, 320, 512], h=[4, 3, 3, 2],
                 w=[4, 3, 3, 2], cross_region_step=[2, 2, 1, 1], cross_region_interval=2, depth=[4, 6, 24, 3],
                 expansion_factor=2, patcher_norm=False, padding_type='circular'):
        patch_size = pair(patch_size)
        super().__init__()
        self.patcher = PatchEmbedding(dim_in=in_channels, dim_out=d_model[0], kernel_size=7, stride=patch_size, padding=3,
                                      norm_layer=patcher_norm)
        self.layers = nn.ModuleList()
        for i_layer in range(len(depth)):
            self.layers.append(HireMLPStage(
                h[i_layer], w[i_layer], d_model[i_layer],
                d_model_out=d_model[i_layer + 1] if (i_layer + 1 < len(depth)) else d_model[-1], depth=depth[i_layer],
                cross_region_step=cross_region_step[i_layer], cross_region_interval=cross_region_interval,
                expansion_factor=expansion_factor, pooling=((i_layer + 1) < len(depth)), padding_type=padding_type))
        self.mlp_head = nn.Sequential(nn.LayerNorm(d_model[-1]), Reduce('b h w c -> b c', 'mean'),
                                      nn.Linear(d_model[-1], num_classes))

    def forward(self, x):
        embedding = self.patcher(x.contiguous(memory_format=torch.channels_last))   # cuDNN NHWC kernel, no layout passes
        embedding = embedding.permute(0, 2, 3, 1)
        for layer in self.layers:
            embedding = layer(embedding)
        ln = self.mlp_head[0]                               # LayerNorm -> position mean -> Linear (hire_mlp.py:217-221)
        embedding = fn.layer_norm(embedding.contiguous(), ln.weight, ln.bias, ln.eps)
        return fn.head(embedding, self.mlp_head[2])
