"""MLP-Mixer with the block body replaced by the fused sm_100a path.

Drop-in for /root/reference/models_pytorch/mlp_mixer.py: same class names, constructor
signatures, defaults and ``state_dict`` keys (``model.{i}.0.fn.net.0.weight`` ...), so
``load_state_dict(reference.state_dict(), strict=True)`` is the integration test.
The module tree exists to own parameters under the reference's names; ``MixerBlock.forward``
hands them to one C-ABI call instead of running the ~12 ATen ops of mlp_mixer.py:12-25.
"""
from functools import partial

from torch import nn

from . import fn, ops
from .utils import check_sizes


class PreNormResidual(nn.Module):
    """Parameter container mirroring mlp_mixer.py:6-13 (fn(norm(x)) + x)."""

    def __init__(self, dim, fn):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)


class FeedForward(nn.Module):
    """Parameter container mirroring mlp_mixer.py:16-27 (dense, GELU, Dropout, dense, Dropout)."""

    def __init__(self, dim, hidden_dim, dropout=0., dense=nn.Linear):
        super().__init__()
        if dropout != 0.:
            raise ValueError("the fused block implements dropout = 0 only (the reference default)")
        self.net = nn.Sequential(dense(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout), dense(hidden_dim, dim),
                                 nn.Dropout(dropout))


class MixerBlock(nn.Sequential):
    """One element of MLPMixer.model (mlp_mixer.py:36-39): token-mixing then channel-mixing half."""

    def forward(self, x):
        tok, chn = self[0], self[1]
        if tok.norm.eps != chn.norm.eps:
            raise ValueError("both LayerNorms must share eps")
        return ops.mixer_block(
            x.contiguous(), tok.norm.eps,            # block 0 receives a permuted view (mlp_mixer.py:70-71)
            tok.norm.weight, tok.norm.bias,
            tok.fn.net[0].weight, tok.fn.net[0].bias, tok.fn.net[3].weight, tok.fn.net[3].bias,
            chn.norm.weight, chn.norm.bias,
            chn.fn.net[0].weight, chn.fn.net[0].bias, chn.fn.net[3].weight, chn.fn.net[3].bias)


class MLPMixer(nn.Module):
    def __init__(self, num_patches, d_model, depth, expansion_factor=4, dropout=0.):
        super().__init__()
        chan_first, chan_last = partial(nn.Conv1d, kernel_size=1), nn.Linear
        self.model = nn.Sequential(
            *[MixerBlock(
                PreNormResidual(d_model, FeedForward(num_patches, num_patches * expansion_factor, dropout, chan_first)),
                PreNormResidual(d_model, FeedForward(d_model, d_model * expansion_factor, dropout, chan_last)),
            ) for _ in range(depth)])

    def forward(self, x):
        return self.model(x)


class MLPMixerForImageClassification(MLPMixer):
    def __init__(self, in_channels=3, d_model=512, num_classes=1000, patch_size=16, image_size=224, depth=12,
                 expansion_factor=4):
        num_patches = check_sizes(image_size, patch_size)
        super().__init__(num_patches, d_model, depth, expansion_factor)
        self.patcher = nn.Sequential(nn.Conv2d(in_channels, d_model, kernel_size=patch_size, stride=patch_size))
        self.active = nn.LayerNorm(d_model)
        self.mlp_head = nn.Sequential(nn.Linear(d_model, num_classes))

    def forward(self, x):
        patches = fn.patch_embed(x, self.patcher[0])        # stem conv as gather + GEMM -> contiguous [B, N, C]
        embedding = self.model(patches)
        embedding = fn.layer_norm(embedding, self.active.weight, self.active.bias, self.active.eps)
        return fn.head(embedding, self.mlp_head[0])           # token mean + Linear (mlp_mixer.py:75-76)
