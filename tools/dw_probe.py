import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jittor_mlp_b200 import fn_spatial
x = torch.randn(64, 32, 32, 768, device="cuda").bfloat16()
w = (torch.randn(768, 1, 7, 7, device="cuda") * 0.1).bfloat16(); b = torch.randn(768, device="cuda").bfloat16()
with torch.no_grad():
    for _ in range(3):
        y = fn_spatial.DwConvGeluFn.apply(x, w, b)
torch.cuda.synchronize()
