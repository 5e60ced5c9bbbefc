#!/bin/bash
# round 2, second half: new rows (ViP, fused optimizer), the re-scheduled forward token epilogue, regression of the split attention
mkdir -p gpurun_out
timeout 600 python tools/tokmix_check.py --all > gpurun_out/tokmix_check.log 2>&1; tail -3 gpurun_out/tokmix_check.log
timeout 300 python tools/bench_tokmix.py > gpurun_out/bench_tokmix.json 2> gpurun_out/bench_tokmix.err; cat gpurun_out/bench_tokmix.json
timeout 900 python -m pytest tests/test_vip_gpu.py tests/test_shift_family_gpu.py tests/test_tokmix_gpu.py tests/test_kernels2_gpu.py -q -x -m gpu 2>&1 | tail -15
