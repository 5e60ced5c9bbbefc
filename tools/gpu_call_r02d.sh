#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02_bench_models_b.jsonl
for m in sparse_mlp_t vip_s; do
  python bench.py --model $m --no-cpu-baseline --steps 10 2>gpurun_out/bench_$m.err | grep '^{' >> gpurun_out/r02_bench_models_b.jsonl
  tail -1 gpurun_out/r02_bench_models_b.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'] if d.get('roofline') else None, d['clocks']['sm_mhz'], d['config'].get('cuda_graph'))"
  tail -2 gpurun_out/bench_$m.err | cut -c1-300
done
