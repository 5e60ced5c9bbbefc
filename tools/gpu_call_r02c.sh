#!/bin/bash
# 2-GPU call: lock-step backward re-check, then the data-parallel step with the gradient arena
mkdir -p gpurun_out
timeout 300 python tools/bench_tokmix.py > gpurun_out/bench_tokmix.json 2> gpurun_out/bench_tokmix.err; cat gpurun_out/bench_tokmix.json | cut -c1-330
timeout 600 python -m pytest tests/test_tokmix_gpu.py tests/test_mixer_gpu.py tests/test_dp_gpu.py -q -x -m gpu > gpurun_out/pytest_b.log 2>&1; tail -3 gpurun_out/pytest_b.log
N=2 bash tools/gpu_dp_r02.sh
