#!/bin/bash
# bench line of every preset + per-kernel launch lists of the shift family (one step each)
mkdir -p gpurun_out
: > gpurun_out/models_bench.jsonl
for m in ${MODELS:-resmlp_24 gmlp_s as_mlp_t s2mlpv2 hire_t s2mlpv1_deep convmixer_768_32}; do
  python bench.py --model $m --no-kernels --no-cpu-baseline --steps 10 2>/dev/null | grep '^{' >> gpurun_out/models_bench.jsonl
  tail -1 gpurun_out/models_bench.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m', d['value'], d['ms_per_step'], d['e2e']['value'], d['model_frac_of_sustained_peak'], d['gpu_launches_per_step'], d['clocks']['sm_mhz'])"
done
for m in ${PROFILE_MODELS:-as_mlp_t s2mlpv2 hire_t convmixer_768_32}; do
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$m.csv python tools/one_step.py $m > /dev/null 2>&1
  echo "== $m"; python tools/summarize_launches.py gpurun_out/launches_$m.csv 22
done
