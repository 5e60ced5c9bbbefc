"""gMLP with the block body on the fused sm_100a path.

Drop-in for /root/reference/models_pytorch/g_mlp.py (same classes, constructor signatures, defaults incl.
``image_size=256``, state_dict keys).  ``sgu.spatial_proj.bias`` is initialised to 1.0 like g_mlp.py:15.
"""
from torch import nn

from . import fn, ops
from .utils import check_sizes


class SpatialGatingUnit(nn.Module):
    """Parameter container for g_mlp.py:10-22."""

    def __init__(self, d_ffn, seq_len):
        super().__init__()
        self.norm = nn.LayerNorm(d_ffn)
        self.spatial_proj = nn.Conv1d(seq_len, seq_len, kernel_size=1)
        nn.init.constant_(self.spatial_proj.bias, 1.0)


class gMLPBlock(nn.Module):
    def __init__(self, d_model, d_ffn, seq_len):
        super().__init__()
        self.norm = nn.LayerNorm(d_model)
        self.channel_proj1 = nn.Linear(d_model, d_ffn * 2)
        self.channel_proj2 = nn.Linear(d_ffn, d_model)
        self.sgu = SpatialGatingUnit(d_ffn, seq_len)

    def forward(self, x):
        return ops.GMLPBlockFn.apply(
            x.contiguous(), self.norm.eps, self.sgu.norm.eps, self.norm.weight, self.norm.bias,
            self.channel_proj1.weight, self.channel_proj1.bias, self.channel_proj2.weight, self.channel_proj2.bias,
            self.sgu.norm.weight, self.sgu.norm.bias, self.sgu.spatial_proj.weight, self.sgu.spatial_proj.bias)


class gMLP(nn.Module):
    def __init__(self, d_model=256, d_ffn=1536, seq_len=256, depth=30):
        super().__init__()
        self.model = nn.Sequential(*[gMLPBlock(d_model, d_ffn, seq_len) for _ in range(depth)])

    def forward(self, x):
        return self.model(x)


class gMLPForImageClassification(gMLP):
    def __init__(self, image_size=256, patch_size=16, in_channels=3, num_classes=1000, d_model=256, d_ffn=1536,
                 depth=30):
        num_patches = check_sizes(image_size, patch_size)
        super().__init__(d_model, d_ffn, num_patches, depth)
        self.patcher = nn.Sequential(nn.Conv2d(in_channels, d_model, kernel_size=patch_size, stride=patch_size))
        self.mlp_head = nn.Sequential(nn.Linear(d_model, num_classes))

    def forward(self, x):
        patches = fn.patch_embed(x, self.patcher[0])        # stem conv as gather + GEMM -> contiguous [B, N, C]
        embedding = self.model(patches)
        return fn.head(embedding, self.mlp_head[0])           # token mean + Linear
