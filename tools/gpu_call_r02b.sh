#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/tokmix_stress.py bwd 40 2>/dev/null | tail -2
timeout 600 python tools/tokmix_check.py --all > gpurun_out/tokmix_check.log 2>&1; tail -2 gpurun_out/tokmix_check.log
timeout 300 python tools/bench_tokmix.py > gpurun_out/bench_tokmix.json 2> gpurun_out/bench_tokmix.err; cat gpurun_out/bench_tokmix.json; tail -3 gpurun_out/bench_tokmix.err
timeout 300 python tools/tokmix_trace.py bwd > gpurun_out/tokmix_trace_bwd.log 2>&1; head -3 gpurun_out/tokmix_trace_bwd.log
timeout 600 python -m pytest tests/test_tokmix_gpu.py tests/test_mixer_gpu.py -q -x -m gpu > gpurun_out/pytest_b.log 2>&1; tail -3 gpurun_out/pytest_b.log
