#!/bin/bash
mkdir -p gpurun_out
python tools/tokmix_check.py --all > gpurun_out/tokmix_bringup.log 2>&1; grep -E "^TOKMIX" gpurun_out/tokmix_bringup.log | tail -4
python tools/bench_tokmix.py > gpurun_out/bench_tokmix.log 2>&1; tail -2 gpurun_out/bench_tokmix.log
TOKMIX_ONLY=fused_fwd timeout 300 ncu --set full --clock-control none --import-source on -k regex:tokmix_fwd --launch-skip 3 -c 1 -o gpurun_out/r02_tokmix_fwd_v2 -f python tools/bench_tokmix.py > gpurun_out/ncu_tokmix_fwd.log 2>&1
TOKMIX_ONLY=fused_bwd timeout 300 ncu --set full --clock-control none --import-source on -k regex:tokmix_bwd --launch-skip 3 -c 1 -o gpurun_out/r02_tokmix_bwd_v2 -f python tools/bench_tokmix.py > gpurun_out/ncu_tokmix_bwd.log 2>&1
tail -2 gpurun_out/ncu_tokmix_bwd.log
( timeout 900 python bench.py ) > gpurun_out/bench_fused.log 2> gpurun_out/bench_fused.err; tail -c 1200 gpurun_out/bench_fused.log; tail -3 gpurun_out/bench_fused.err
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
