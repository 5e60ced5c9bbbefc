#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/tokmix_stress.py bwd 24 2>/dev/null | tail -1
TOKMIX_ONLY=fused_bwd timeout 200 python tools/bench_tokmix.py 2>/dev/null | cut -c1-200
timeout 600 python -m pytest tests/test_tokmix_gpu.py tests/test_mixer_gpu.py -q -x -m gpu 2>&1 | tail -2
