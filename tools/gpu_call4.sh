#!/bin/bash
mkdir -p gpurun_out
python tools/bench_tokmix.py > gpurun_out/bench_tokmix.log 2>&1; tail -3 gpurun_out/bench_tokmix.log
# ncu: one launch of each fused kernel, full set with source
TOKMIX_ONLY=fused_fwd,fused_bwd timeout 600 ncu --set full --clock-control none --import-source on -k regex:tokmix_ --launch-skip 6 -c 2 -o gpurun_out/r02_tokmix_v1 -f python tools/bench_tokmix.py > gpurun_out/ncu_tokmix.log 2>&1
tail -3 gpurun_out/ncu_tokmix.log
python tools/step_timeline.py mixer_b16 256 1 > gpurun_out/timeline_fused.log 2>&1; head -24 gpurun_out/timeline_fused.log
VMLP_TOKMIX=0 python tools/step_timeline.py mixer_b16 256 1 > gpurun_out/timeline_unfused.log 2>&1; head -24 gpurun_out/timeline_unfused.log
