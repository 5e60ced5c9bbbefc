#!/bin/bash
# round-2 evidence: ncu full set of one MixerBlock fwd+bwd (summary CSV), step launch list, bench lines of every preset
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/r02_block -f python tools/one_block.py 1 > gpurun_out/ncu_block.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_block.ncu-rep > gpurun_out/r02_ncu_full_mixer_block_v1_summary.csv 2>> gpurun_out/ncu_block.log
cat gpurun_out/r02_ncu_full_mixer_block_v1_summary.csv | cut -c1-160
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_mixer_b16.csv python tools/one_step.py mixer_b16 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_mixer_b16.csv 16
python tools/step_timeline.py mixer_b16 256 1 > gpurun_out/r02_timeline_mixer_b16.log 2>&1; head -16 gpurun_out/r02_timeline_mixer_b16.log
: > gpurun_out/r02_bench_models.jsonl
for m in ${MODELS:-mixer_l16 resmlp_24 gmlp_s as_mlp_t s2mlpv2 hire_t s2mlpv1_deep convmixer_768_32 vip_s}; do
  python bench.py --model $m --no-cpu-baseline --steps 10 2>/dev/null | grep '^{' >> gpurun_out/r02_bench_models.jsonl
  tail -1 gpurun_out/r02_bench_models.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$m', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'] if d.get('roofline') else None, d['clocks']['sm_mhz'])"
done
