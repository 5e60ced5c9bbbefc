"""SparseMLP (sMLP) with the block bodies on the sm_100a path -- SURVEY.md row f3 (first of its three models).

Drop-in for /root/reference/models_pytorch/sparse_mlp.py (same classes, constructor signatures, defaults, state_dict
keys; the parameterless einops `Rearrange` layers become `nn.Identity` place-holders so the Sequential indices stay).
Activations are channels-last rows [B, H, W, C] throughout: the reference's NCHW <-> NHWC rearrangements disappear,
BatchNorm2d works on the rows (fn_spatial.BatchNormFn), the depthwise 3x3 conv is the TMA-halo stencil without activation,
the axial Linears along H and W are token-axis GEMMs (fn.TokenLinearFn), `cat` + 1x1 `fuse` is one strided-copy
concatenation + one K-major GEMM with the residual in its epilogue.
"""
import torch
from torch import nn

from . import fn, fn_spatial
from .conv_mixer import _bn
from .s2_mlp import _check_dropout, _ff
from .utils import pair


class PreNormResidual(nn.Module):
    """Parameter container (sparse_mlp.py:8-15)."""

    def __init__(self, dim, fn_, norm=nn.LayerNorm):
        super().__init__()
        self.fn = fn_
        self.norm = norm(dim)


class PatchMerging(nn.Module):
    """sparse_mlp.py:17-58: 2x2 space-to-depth (x0, x1, x2, x3 order), LayerNorm(4C), Linear(4C -> 2C, no bias)."""

    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward(self, x):                                   # [B, H, W, C] -> [B, H/2, W/2, 2C]
        B, H, W, C = x.shape
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        x = fn.ConcatChannelsFn.apply(x[:, 0::2, 0::2].contiguous(), x[:, 1::2, 0::2].contiguous(),
                                      x[:, 0::2, 1::2].contiguous(), x[:, 1::2, 1::2].contiguous())
        x = fn.layer_norm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        return fn.linear(x, self.reduction.weight, None)


class sMLPBlock(nn.Module):
    def __init__(self, h=224, w=224, d_model=3):
        super().__init__()
        self.proj_h = nn.Linear(h, h)
        self.proj_w = nn.Linear(w, w)
        self.fuse = nn.Conv2d(3 * d_model, d_model, kernel_size=1)

    def run(self, xn, res):
        """fuse(cat[proj_h along H, proj_w along W, identity]) + res (sparse_mlp.py:67-74); xn: [B, H, W, C]."""
        B, H, W, C = xn.shape
        x_h = fn.TokenLinearFn.apply(xn.view(B, H, W * C), self.proj_h.weight, self.proj_h.bias).view(B, H, W, C)
        x_w = fn.TokenLinearFn.apply(xn.view(B * H, W, C), self.proj_w.weight, self.proj_w.bias).view(B, H, W, C)
        t = fn.ConcatChannelsFn.apply(x_h, x_w, xn)
        return fn.linear(t, self.fuse.weight, self.fuse.bias, res)


class sMLPStage(nn.Module):
    def __init__(self, height, width, d_model, depth, expansion_factor=2, dropout=0., pooling=False):
        super().__init__()
        _check_dropout(dropout)
        self.pooling = pooling
        self.patch_merge = nn.Sequential(nn.Identity(), PatchMerging((height, width), d_model), nn.Identity())
        self.model = nn.Sequential(
            *[nn.Sequential(
                PreNormResidual(d_model, nn.Sequential(
                    nn.Conv2d(d_model, d_model, kernel_size=3, padding=1, groups=d_model)), norm=nn.BatchNorm2d),
                PreNormResidual(d_model, nn.Sequential(sMLPBlock(height, width, d_model)), norm=nn.BatchNorm2d),
                nn.Identity(),
                PreNormResidual(d_model, nn.Sequential(
                    nn.Linear(d_model, d_model * expansion_factor), nn.GELU(), nn.Dropout(dropout),
                    nn.Linear(d_model * expansion_factor, d_model), nn.Dropout(dropout)), norm=nn.LayerNorm),
                nn.Identity(),
            ) for _ in range(depth)])

    def forward(self, x):                                   # [B, H, W, C] channels-last
        for blk in self.model:
            a, b, c = blk[0], blk[1], blk[3]
            conv = a.fn[0]
            x = fn_spatial.DwConvFn.apply(_bn(x, a.norm), conv.weight, conv.bias) + x      # sparse_mlp.py:88-91
            x = b.fn[0].run(_bn(x, b.norm), x)                                              # :92-96
            x = _ff(c.fn, fn.layer_norm(x, c.norm.weight, c.norm.bias, c.norm.eps), x)      # :98-104
        if self.pooling:
            x = self.patch_merge[1](x)
        return x


class SparseMLP(nn.Module):
    def __init__(self, image_size=224, patch_size=4, in_channels=3, num_classes=1000, d_model=96, depth=[2, 10, 24, 2],
                 expansion_factor=2, patcher_norm=False):
        image_size = pair(image_size)
        patch_size = pair(patch_size)
        assert (image_size[0] % patch_size[0]) == 0, 'image must be divisible by patch size'
        assert (image_size[1] % patch_size[1]) == 0, 'image must be divisible by patch size'
        height = image_size[0] // patch_size[0]
        width = image_size[1] // patch_size[1]
        super().__init__()
        self.patcher = nn.Sequential(
            nn.Conv2d(in_channels, d_model, kernel_size=patch_size, stride=patch_size),
            nn.Identity() if (not patcher_norm) else nn.Sequential(nn.Identity(), nn.LayerNorm(d_model), nn.Identity()))
        self.layers = nn.ModuleList()
        for i_layer in range(len(depth)):
            self.layers.append(sMLPStage(height // (2 ** i_layer), width // (2 ** i_layer), d_model, depth[i_layer],
                                         expansion_factor=expansion_factor, pooling=((i_layer + 1) < len(depth))))
            if (i_layer + 1) < len(depth):
                d_model = d_model * 2
        self.mlp_head = nn.Sequential(nn.Identity(), nn.LayerNorm(d_model), nn.Identity(), nn.Linear(d_model, num_classes))

    def _bn_buffers_fp32(self):
        # running statistics are updated in fp32 by the kernels even when the module was cast with .bfloat16()
        for m in self.layers.modules():
            if isinstance(m, nn.BatchNorm2d) and m.running_mean is not None and m.running_mean.dtype != torch.float32:
                m.running_mean = m.running_mean.float()
                m.running_var = m.running_var.float()

    def forward(self, x):
        self._bn_buffers_fp32()
        x = self.patcher[0](x.contiguous(memory_format=torch.channels_last)).permute(0, 2, 3, 1).contiguous()
        if not isinstance(self.patcher[1], nn.Identity):
            ln = self.patcher[1][1]
            x = fn.layer_norm(x, ln.weight, ln.bias, ln.eps)
        for layer in self.layers:
            x = layer(x)
        ln = self.mlp_head[1]
        return fn.head(fn.layer_norm(x, ln.weight, ln.bias, ln.eps), self.mlp_head[3])
