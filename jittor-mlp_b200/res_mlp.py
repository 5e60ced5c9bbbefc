"""ResMLP with the block body on the fused sm_100a path.

Drop-in for /root/reference/models_pytorch/res_mlp.py (same classes, constructor signatures, defaults, state_dict
keys).  Quirks preserved: the token-mixing residual is taken after the pre-affine (res_mlp.py:53-54);
``ResMLPForImageClassification.affine`` exists but is never applied (res_mlp.py:86, forward :91-99).
"""
import torch
from torch import nn

from . import fn, ops
from .utils import check_sizes


class Aff(nn.Module):
    """Parameter container for res_mlp.py:11-19 (x * alpha + beta)."""

    def __init__(self, dim):
        super().__init__()
        self.alpha = nn.Parameter(torch.ones([1, 1, dim]))
        self.beta = nn.Parameter(torch.zeros([1, 1, dim]))


class FeedForward(nn.Module):
    """Parameter container for res_mlp.py:21-32."""

    def __init__(self, dim, hidden_dim, dropout=0.):
        super().__init__()
        if dropout != 0.:
            raise ValueError("the fused block implements dropout = 0 only (the reference default)")
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout), nn.Linear(hidden_dim, dim),
                                 nn.Dropout(dropout))


class MLPblock(nn.Module):
    def __init__(self, num_patch, dim, mlp_dim, dropout=0., depth=18):
        super().__init__()
        if depth <= 18:
            init_values = 0.1
        elif depth > 18 and depth <= 24:
            init_values = 1e-5
        else:
            init_values = 1e-6
        self.pre_affine = Aff(dim)
        self.token_mix = nn.Conv1d(num_patch, num_patch, kernel_size=1)
        self.ff = FeedForward(dim, mlp_dim, dropout)
        self.post_affine = Aff(dim)
        self.gamma_1 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)
        self.gamma_2 = nn.Parameter(init_values * torch.ones((dim)), requires_grad=True)

    def forward(self, x):
        return ops.ResMLPBlockFn.apply(
            x.contiguous(), self.pre_affine.alpha, self.pre_affine.beta, self.token_mix.weight, self.token_mix.bias,
            self.ff.net[0].weight, self.ff.net[0].bias, self.ff.net[3].weight, self.ff.net[3].bias,
            self.post_affine.alpha, self.post_affine.beta, self.gamma_1, self.gamma_2)


class ResMLP(nn.Module):
    def __init__(self, num_patch, d_model, depth, expansion_factor):
        super().__init__()
        self.model = nn.Sequential(
            *[MLPblock(num_patch, d_model, d_model * expansion_factor, depth=depth) for _ in range(depth)])

    def forward(self, x):
        return self.model(x)


class ResMLPForImageClassification(ResMLP):
    def __init__(self, in_channels=3, d_model=384, num_classes=1000, patch_size=16, image_size=224, depth=12,
                 expansion_factor=4):
        num_patches = check_sizes(image_size, patch_size)
        super().__init__(num_patches, d_model, depth, expansion_factor)
        self.patcher = nn.Sequential(nn.Conv2d(in_channels, d_model, kernel_size=patch_size, stride=patch_size))
        self.affine = Aff(d_model)      # created, never used -- exactly like the reference
        self.mlp_head = nn.Sequential(nn.Linear(d_model, num_classes))

    def forward(self, x):
        patches = fn.patch_embed(x, self.patcher[0])        # stem conv as gather + GEMM -> contiguous [B, N, C]
        embedding = self.model(patches)
        return fn.head(embedding, self.mlp_head[0])           # token mean + Linear
