#!/bin/bash
# same-box A/B of the whole Mixer-B/16 step: fused token-mixing kernels (default) vs the unfused GEMM sequence (VMLP_TOKMIX=0),
# interleaved A B A B so that clock drift under the power cap hits both arms alike
mkdir -p gpurun_out
: > gpurun_out/r02_ab_tokmix.jsonl
for rep in 1 2; do
  for arm in 1 0; do
    VMLP_TOKMIX=$arm python bench.py --no-cpu-baseline --no-kernels --steps 20 2>/dev/null | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'fused_token_kernels': bool($arm), 'rep': $rep, 'images_per_s': d['value'], 'ms_per_step': d['ms_per_step'], 'e2e': d['e2e']['value'], 'sm_mhz': d['clocks']['sm_mhz'], 'reasons': d['clocks']['reasons']}))" | tee -a gpurun_out/r02_ab_tokmix.jsonl
  done
done
