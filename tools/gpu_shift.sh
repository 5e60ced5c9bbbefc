#!/bin/bash
# shift-family iteration loop: parity tests, bench lines, launch lists
mkdir -p gpurun_out
python -m pytest tests/test_shift_family_gpu.py tests/test_rowwise_gpu.py -x -q 2>&1 | tail -3
MODELS="${MODELS:-as_mlp_t s2mlpv2 hire_t}" PROFILE_MODELS="${PROFILE_MODELS:-as_mlp_t s2mlpv2 hire_t}" bash tools/gpu_models.sh
