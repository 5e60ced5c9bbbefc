#!/bin/bash
# quick iteration loop: GEMM + Mixer parity, A/B of one block against a previous build, per-kernel bench table
mkdir -p gpurun_out
python -m pytest tests/test_gemm_gpu.py tests/test_mixer_gpu.py tests/test_resmlp_gmlp_gpu.py -x -q 2>&1 | tail -3
bash tools/ab_block.sh
python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_quick.log 2>/dev/null
python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_quick.log') if x.startswith('{')]
d=json.loads(l[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
for k,v in d.get('kernels',{}).items(): print(f"{k:34s} {v['ms']:.4f} {v['tflops']:7.1f} {v['frac_of_sustained_peak']}")
print(d.get('block_gemm_ms'))
PY
