#!/bin/bash
# end-of-round check on one GPU: smoke(), the whole GPU suite, the default bench line, a bench line with the fused optimizer
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
bash tools/gpu_validate.sh
python bench.py --optimizer adamw --no-kernels --no-cpu-baseline --steps 10 2>/dev/null | grep '^{' > gpurun_out/r02_bench_mixer_b16_adamw.json; python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_mixer_b16_adamw.json').read()); print('adamw', d['value'], d['ms_per_step'], d['config']['optimizer'])"
