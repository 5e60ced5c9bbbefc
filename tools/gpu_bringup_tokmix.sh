#!/bin/bash
# GPU call: bring-up of the fused token-mixing kernels, then the parity suite and the bench (fused and unfused token path)
mkdir -p gpurun_out
python tools/tokmix_check.py --all > gpurun_out/tokmix_bringup.log 2>&1
rc=$?
tail -40 gpurun_out/tokmix_bringup.log
if [ $rc -eq 0 ]; then
  ( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
  tail -30 gpurun_out/pytest_gpu.log
  ( timeout 600 python bench.py ) > gpurun_out/bench_fused.log 2> gpurun_out/bench_fused.err
  tail -c 1500 gpurun_out/bench_fused.log; tail -3 gpurun_out/bench_fused.err
  ( VMLP_TOKMIX=0 timeout 600 python bench.py ) > gpurun_out/bench_unfused.log 2> gpurun_out/bench_unfused.err
  tail -c 600 gpurun_out/bench_unfused.log
else
  ( time VMLP_TOKMIX=0 timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_tokmix_gpu.py ) > gpurun_out/pytest_gpu.log 2>&1
  tail -30 gpurun_out/pytest_gpu.log
  ( VMLP_TOKMIX=0 timeout 600 python bench.py ) > gpurun_out/bench_unfused.log 2> gpurun_out/bench_unfused.err
  tail -c 600 gpurun_out/bench_unfused.log
fi
