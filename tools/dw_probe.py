"""One ConvMixer-768 depthwise layer forward + backward (profiling target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jittor_mlp_b200 import fn_spatial  # noqa: E402

B = 256
x = torch.randn(B, 32, 32, 768, device="cuda").bfloat16().requires_grad_(True)
w = (torch.randn(768, 1, 7, 7, device="cuda") * 0.1).bfloat16().requires_grad_(True)
b = torch.randn(768, device="cuda").bfloat16().requires_grad_(True)
for _ in range(2):
    y = fn_spatial.DwConvGeluFn.apply(x, w, b)
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()
