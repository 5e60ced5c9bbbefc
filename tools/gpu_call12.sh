#!/bin/bash
mkdir -p gpurun_out
python tools/tokmix_check.py --all > gpurun_out/tokmix_bringup.log 2>&1; grep -E "^TOKMIX" gpurun_out/tokmix_bringup.log | grep -v "fwd.*0.00" | tail -6 | cut -c1-300
python tools/bench_tokmix.py > gpurun_out/bench_tokmix.log 2>&1; tail -1 gpurun_out/bench_tokmix.log | cut -c1-900
if grep -q "all green" gpurun_out/tokmix_bringup.log; then
  ( timeout 900 python bench.py ) > gpurun_out/bench_fused.log 2> gpurun_out/bench_fused.err; tail -c 600 gpurun_out/bench_fused.log | head -c 400; tail -2 gpurun_out/bench_fused.err
  ( VMLP_TOKMIX=0 timeout 900 python bench.py --no-cpu-baseline ) > gpurun_out/bench_unfused.log 2> gpurun_out/bench_unfused.err
  ( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
fi
