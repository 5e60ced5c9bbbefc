"""In-situ kernel times of one training step (CUPTI through torch.profiler): unlike the serialised, cold-cache ncu launch
list this runs the real back-to-back step under the sustained power cap, so per-kernel SHARES and the total idle gap
between kernels can be read off.  python tools/step_timeline.py [preset] [batch] [graph:0|1]  -> one JSON line + table."""
import collections
import json
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import jittor_mlp_b200 as J  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mixer_b16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
use_graph = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
dev = torch.device("cuda", 0)
model = bench.build_model(name, dev)
x = torch.randn(B, 3, 224, 224, device=dev).bfloat16()
if use_graph:
    gs = J.GraphedStep(model, x, bench.loss_fn)
    step = gs.run
else:
    def step():
        model.zero_grad(set_to_none=True)
        bench.loss_fn(model(x)).backward()
for _ in range(5):
    step()
torch.cuda.synchronize()
STEPS = 4
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(STEPS):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0]
ev.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
busy = 0.0
for e in ev:
    k = re.sub(r"\(.*", "", re.sub(r"^void ", "", e.name))[:90]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e.device_time_total
    busy += e.device_time_total
span = ev[-1].time_range.end - ev[0].time_range.start
print(json.dumps({"preset": name, "batch": B, "graph": use_graph, "steps": STEPS, "span_ms_per_step": round(span / STEPS / 1e3, 3),
                  "kernel_busy_ms_per_step": round(busy / STEPS / 1e3, 3), "launches_per_step": len(ev) // STEPS}))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{t / STEPS / 1e3:9.3f} ms {100 * t / busy:5.1f}% {n // STEPS:5d}  {k}")
