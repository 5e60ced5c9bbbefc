"""ConvMixer with the block bodies on the sm_100a path.

Drop-in for /root/reference/models_pytorch/conv_mixer.py (same classes, constructor signature, state_dict keys incl. the
BatchNorm buffers).  Blocks run channels-last: the depthwise k x k conv is a shared-memory stencil, the 1x1 conv a
K-major GEMM with bias+GELU in the epilogue, BatchNorm uses batch statistics in train() (updating running_mean /
running_var / num_batches_tracked like nn.BatchNorm2d) and running statistics in eval().
"""
import torch
import torch.nn as nn

from . import fn, fn_spatial


class Residual(nn.Module):
    def __init__(self, fn_):
        super().__init__()
        self.fn = fn_


def _bn(a, bn, res=None, link=None):
    if bn.training:
        if bn.track_running_stats:
            bn.num_batches_tracked += 1
        mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
        return fn_spatial.BatchNormFn.apply(a, bn.weight, bn.bias, bn.running_mean if bn.track_running_stats else None,
                                            bn.running_var if bn.track_running_stats else None, mom, bn.eps, res, link)
    return fn_spatial.batch_norm_eval(a, bn, res)


class ConvMixer(nn.Module):
    def __init__(self, dim, depth, kernel_size=9, patch_size=7, n_classes=1000):
        super().__init__()
        if kernel_size not in (3, 5, 7, 9):
            raise ValueError("depthwise kernel_size must be one of 3, 5, 7, 9")
        self.embedding = nn.Sequential(
            nn.Conv2d(3, dim, kernel_size=patch_size, stride=patch_size, padding=patch_size // 2), nn.GELU(),
            nn.BatchNorm2d(dim))
        self.blocks = nn.Sequential(
            *[nn.Sequential(
                Residual(nn.Sequential(nn.Conv2d(dim, dim, kernel_size, groups=dim, padding="same"), nn.GELU(),
                                       nn.BatchNorm2d(dim))),
                nn.Conv2d(dim, dim, kernel_size=1), nn.GELU(), nn.BatchNorm2d(dim)) for i in range(depth)])
        self.classifier = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Flatten(), nn.Linear(dim, n_classes))

    def _bn_buffers_fp32(self):
        # running statistics are updated in fp32 by the kernels even when the module was cast with .bfloat16()
        for m in self.blocks.modules():
            if isinstance(m, nn.BatchNorm2d) and m.running_mean is not None and m.running_mean.dtype != torch.float32:
                m.running_mean = m.running_mean.float()
                m.running_var = m.running_var.float()

    def forward(self, x):
        self._bn_buffers_fp32()
        x = self.embedding(x.contiguous(memory_format=torch.channels_last))   # stem: cuDNN conv + GELU + BatchNorm (torch), NHWC
        x = x.permute(0, 2, 3, 1).contiguous()              # channels-last rows from here on
        for blk in self.blocks:
            dw, bn1 = blk[0].fn[0], blk[0].fn[2]
            # each GELU output feeds exactly one BatchNorm: its backward also applies gelu' (fn.GeluLink)
            l1, l2 = fn.GeluLink(), fn.GeluLink()
            a = fn_spatial.DwConvGeluFn.apply(x, dw.weight, dw.bias, l1)
            x = _bn(a, bn1, x, l1)                          # BN(GELU(dwconv(x))) + x   (conv_mixer.py:23-26)
            a = fn.linear_gelu(x, blk[1].weight, blk[1].bias, l2)
            x = _bn(a, blk[3], None, l2)                    # BN(GELU(conv1x1(x)))      (conv_mixer.py:28-31)
        return fn.head(x, self.classifier[2])
