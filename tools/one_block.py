"""One MixerBlock forward+backward at the Mixer-B/16 batch-256 shapes (profiling target for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jittor_mlp_b200 as J  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
m = J.MLPMixer(196, 768, 1).cuda().bfloat16()
x = torch.randn(B, 196, 768, device="cuda").bfloat16().requires_grad_(True)
dy = torch.randn(B, 196, 768, device="cuda").bfloat16()
for _ in range(reps):
    y = m(x)
    y.backward(dy)
    m.zero_grad(set_to_none=True)
    x.grad = None
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.cudart().cudaProfilerStart()          # ncu --profile-from-start off: exactly `reps` blocks are captured
e0.record()
for _ in range(reps):
    y = m(x)
    y.backward(dy)
    m.zero_grad(set_to_none=True)
    x.grad = None
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"block fwd+bwd: {e0.elapsed_time(e1) / reps:.3f} ms  (tensor roofline 1.277 ms at the measured sustained peak)")
