#!/bin/bash
# 2-GPU data-parallel bench: eager step vs the CUDA-graph-captured step (NCCL inside the graph); each under a hard timeout
mkdir -p gpurun_out
for g in 0 1; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29511 + g)) \
    bench.py --gpus 2 --steps 20 --warmup 3 --graph $g --no-kernels > gpurun_out/bench_dp2_graph$g.log 2> gpurun_out/bench_dp2_graph$g.err
  echo "graph=$g rc=$?"; grep '^{' gpurun_out/bench_dp2_graph$g.log | python -c "import json,sys; [print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'], d['clocks']['sm_mhz']) for d in map(json.loads, sys.stdin)]"
  grep -i "capture failed\|error" gpurun_out/bench_dp2_graph$g.err | head -3
done
timeout 120 python bench.py --steps 20 --no-kernels --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "import json,sys; [print('1gpu', d['value'], d['ms_per_step']) for d in map(json.loads, sys.stdin)]"
