#!/bin/bash
# compute-sanitizer over the kernels with hand-rolled synchronisation: memcheck on the GEMM / fused
# token kernels / LayerNorm / depthwise stencil tests, racecheck (shared-memory hazards) on the row-wise and stencil tests.
# tcgen05 / TMA traffic is not instrumented by the tool, generic loads/stores, shared-memory accesses and barriers are.
mkdir -p gpurun_out
L=gpurun_out/sanitize.log
: > $L
run() { echo "== $*" >> $L; timeout 900 "$@" >> $L 2>&1; echo "exit $?" >> $L; }
if [ -z "$SANITIZE_NEW" ]; then
run compute-sanitizer --tool memcheck --error-exitcode 3 python tools/tokmix_check.py 3 64 128 256 both
run compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_rowwise_gpu.py tests/test_gemm_gpu.py -q -x -k "plain_store or bias_modes or layernorm or token_weight_gradient"
run compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_rowwise_gpu.py tests/test_kernels2_gpu.py -q -x -k "layernorm or depthwise or colsum"
fi
# round-2 additions: strided copies, optimizer step, head mean, token-axis Linear, plain depthwise conv (SANITIZE_NEW=1: only these)
run compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_vip_gpu.py -q -x -k "permute5 or optimizer_matches or token_mean or token_linear or concat_channels"
grep -E "^==|ERROR SUMMARY|exit |passed|failed" $L
