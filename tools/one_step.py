"""One fwd+bwd step of a bench preset between cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
model = bench.build_model(name, dev)
x = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(1)).bfloat16().to(dev)
for i in range(3):
    if i == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    model.zero_grad(set_to_none=True)
    bench.loss_fn(model(x)).backward()
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
