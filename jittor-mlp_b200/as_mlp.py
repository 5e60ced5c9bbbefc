"""AS-MLP with the block bodies on the sm_100a path.

Drop-in for /root/reference/models_pytorch/as_mlp.py (same classes, constructor signatures, defaults, state_dict
keys).  Internally activations are channels-last [B, H, W, C] rows, so every 1x1 Conv2d is a K-major GEMM, the axial
shift (reference: the CuPy kernel of utils/shift_cuda.py:44-103, two extra read+write passes per call) is one
vectorised gather with per-channel-group offsets, and GroupNorm(1, C) keeps its whole-sample statistics
(as_mlp.py:343-344).  Only the logits leave the model (as_mlp.py:440-443), so the layout change is invisible.
"""
import torch
import torch.utils.checkpoint as checkpoint
from torch import nn

from . import fn


def to_2tuple(v):
    return v if isinstance(v, tuple) else (v, v)


class DropPath(nn.Module):
    """timm.models.layers.DropPath semantics (as_mlp.py:5,145): per-sample stochastic depth, identity in eval."""

    def __init__(self, drop_prob=0.):
        super().__init__()
        self.drop_prob = float(drop_prob or 0.)

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


def _drop_active(m):
    return isinstance(m, DropPath) and m.training and m.drop_prob > 0.


def MyNorm(dim):
    return nn.GroupNorm(1, dim)


def _gn(norm, x, gelu=False):
    if not (isinstance(norm, nn.GroupNorm) and norm.num_groups == 1):
        raise ValueError("only norm_layer = GroupNorm(1, C) (the reference's MyNorm) is implemented")
    return fn.group_norm1(x, norm.weight, norm.bias, norm.eps, gelu)


class Mlp(nn.Module):
    """Parameter container (as_mlp.py:8-24): Conv2d 1x1 -> GELU -> Conv2d 1x1."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        if drop != 0.:
            raise ValueError("the fused blocks implement drop = 0 only (the reference default)")
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Conv2d(in_features, hidden_features, 1, 1)
        self.act = act_layer()
        self.fc2 = nn.Conv2d(hidden_features, out_features, 1, 1)
        self.drop = nn.Dropout(drop)

    def run(self, x, res):
        return fn.mlp(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, res)


class AxialShift(nn.Module):
    def __init__(self, dim, shift_size, as_bias=True, proj_drop=0.):
        super().__init__()
        self.dim = dim
        self.shift_size = shift_size
        self.pad = shift_size // 2
        self.conv1 = nn.Conv2d(dim, dim, 1, 1, 0, groups=1, bias=as_bias)
        self.conv2_1 = nn.Conv2d(dim, dim, 1, 1, 0, groups=1, bias=as_bias)
        self.conv2_2 = nn.Conv2d(dim, dim, 1, 1, 0, groups=1, bias=as_bias)
        self.conv3 = nn.Conv2d(dim, dim, 1, 1, 0, groups=1, bias=as_bias)
        self.actn = nn.GELU()
        self.norm1 = MyNorm(dim)
        self.norm2 = MyNorm(dim)

    def run(self, x, res):
        """conv3(GN(gelu(conv2_1(shift_W(t))) + gelu(conv2_2(shift_H(t))))) (+ res), t = gelu(GN(conv1(x)))
        (as_mlp.py:55-95)."""
        t = _gn(self.norm1, fn.linear(x, self.conv1.weight, self.conv1.bias), gelu=True)
        x_lr = fn.linear_gelu(fn.axial_shift(t, self.shift_size, 3), self.conv2_1.weight, self.conv2_1.bias)
        x_td = fn.linear_gelu(fn.axial_shift(t, self.shift_size, 2), self.conv2_2.weight, self.conv2_2.bias)
        s = _gn(self.norm2, x_lr + x_td)
        return fn.linear(s, self.conv3.weight, self.conv3.bias, res)

    def extra_repr(self) -> str:
        return f'dim={self.dim}, shift_size={self.shift_size}'


class AxialShiftedBlock(nn.Module):
    def __init__(self, dim, input_resolution, shift_size=7, mlp_ratio=4., as_bias=True, drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.shift_size = shift_size
        self.mlp_ratio = mlp_ratio
        self.norm1 = norm_layer(dim)
        self.axial_shift = AxialShift(dim, shift_size=shift_size, as_bias=as_bias, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = Mlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop)

    def forward(self, x):                                   # x: [B, H, W, C] channels-last
        if _drop_active(self.drop_path):
            x = x + self.drop_path(self.axial_shift.run(_gn(self.norm1, x), None))
            return x + self.drop_path(self.mlp.run(_gn(self.norm2, x), None))
        x = self.axial_shift.run(_gn(self.norm1, x), x)     # residual fused into conv3's epilogue
        return self.mlp.run(_gn(self.norm2, x), x)


class PatchMerging(nn.Module):
    def __init__(self, input_resolution, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.input_resolution = input_resolution
        self.dim = dim
        self.reduction = nn.Conv2d(4 * dim, 2 * dim, 1, 1, bias=False)
        self.norm = norm_layer(4 * dim)

    def forward(self, x):                                   # [B, H, W, C] -> [B, H/2, W/2, 2C]
        B, H, W, C = x.shape
        assert H % 2 == 0 and W % 2 == 0, f"x size ({H}*{W}) are not even."
        # 2x2 space-to-depth in the reference's channel order x0, x1, x2, x3 (as_mlp.py:204-209)
        x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
        return fn.linear(_gn(self.norm, x), self.reduction.weight, None)


class BasicLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, shift_size, mlp_ratio=4., as_bias=True, drop=0., drop_path=0.,
                 norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.dim = dim
        self.input_resolution = input_resolution
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            AxialShiftedBlock(dim=dim, input_resolution=input_resolution, shift_size=shift_size, mlp_ratio=mlp_ratio,
                              as_bias=as_bias, drop=drop,
                              drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                              norm_layer=norm_layer)
            for i in range(depth)])
        self.downsample = downsample(input_resolution, dim=dim, norm_layer=norm_layer) if downsample is not None else None

    def forward(self, x):
        for blk in self.blocks:
            x = checkpoint.checkpoint(blk, x, use_reentrant=False) if self.use_checkpoint else blk(x)
        if self.downsample is not None:
            x = self.downsample(x)
        return x


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        img_size = to_2tuple(img_size)
        patch_size = to_2tuple(patch_size)
        patches_resolution = [img_size[0] // patch_size[0], img_size[1] // patch_size[1]]
        self.img_size = img_size
        self.patch_size = patch_size
        self.patches_resolution = patches_resolution
        self.num_patches = patches_resolution[0] * patches_resolution[1]
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        B, C, H, W = x.shape
        assert H == self.img_size[0] and W == self.img_size[1], \
            f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]})."
        # stem conv stays on cuDNN, fed channels-last so that it runs its NHWC kernel directly and the permute below is a
        # view (an NCHW input costs cuDNN's own nchwToNhwc pass plus a strided copy of the [B, C, H/4, W/4] output)
        x = self.proj(x.contiguous(memory_format=torch.channels_last)).permute(0, 2, 3, 1).contiguous()
        if self.norm is not None:
            x = _gn(self.norm, x)
        return x


class AS_MLP(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=[2, 2, 6, 2],
                 shift_size=5, mlp_ratio=4., as_bias=True, drop_rate=0., drop_path_rate=0.1, norm_layer=MyNorm,
                 patch_norm=True, use_checkpoint=False, **kwargs):
        super().__init__()
        if drop_rate != 0.:
            raise ValueError("the fused blocks implement drop_rate = 0 only (the reference default)")
        self.num_classes = num_classes
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.patch_norm = patch_norm
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.mlp_ratio = mlp_ratio
        self.patch_embed = PatchEmbed(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if self.patch_norm else None)
        patches_resolution = self.patch_embed.patches_resolution
        self.patches_resolution = patches_resolution
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        for i_layer in range(self.num_layers):
            self.layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i_layer),
                input_resolution=(patches_resolution[0] // (2 ** i_layer), patches_resolution[1] // (2 ** i_layer)),
                depth=depths[i_layer], shift_size=shift_size, mlp_ratio=self.mlp_ratio, as_bias=as_bias, drop=drop_rate,
                drop_path=dpr[sum(depths[:i_layer]):sum(depths[:i_layer + 1])], norm_layer=norm_layer,
                downsample=PatchMerging if (i_layer < self.num_layers - 1) else None, use_checkpoint=use_checkpoint))
        self.norm = norm_layer(self.num_features)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.head = nn.Linear(self.num_features, num_classes) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward_features(self, x):
        x = self.patch_embed(x)
        for layer in self.layers:
            x = layer(x)
        x = _gn(self.norm, x)                               # [B, H, W, C]
        return fn.TokenMeanFn.apply(x.contiguous())         # AdaptiveAvgPool2d(1) + flatten

    def forward(self, x):
        x = self.forward_features(x)
        if isinstance(self.head, nn.Linear) and self.head.in_features % 8 == 0 and self.head.out_features % 8 == 0:
            return fn.linear(x, self.head.weight, self.head.bias)
        return self.head(x)
