#!/bin/bash
# N-GPU evidence (run with gpurun --gpus N): gradient equality of the DP step (eager and graph mode), then the bench line
N=${N:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tools/dp_check.py > gpurun_out/r02_dp_check_${N}gpu.log 2>&1
grep dp_check gpurun_out/r02_dp_check_${N}gpu.log; tail -3 gpurun_out/r02_dp_check_${N}gpu.log | cut -c1-300
for m in ${MODELS:-mixer_b16}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --model $m --no-kernels > gpurun_out/r02_bench_${m}_${N}gpu.json 2> gpurun_out/r02_bench_${m}_${N}gpu.err
  tail -c 700 gpurun_out/r02_bench_${m}_${N}gpu.json; tail -3 gpurun_out/r02_bench_${m}_${N}gpu.err | cut -c1-300
done
for m in ${MODELS:-mixer_b16}; do
  python bench.py --gpus 1 --model $m --no-kernels --no-cpu-baseline > gpurun_out/r02_bench_${m}_1gpu_samebox_n${N}.json 2>/dev/null; tail -c 400 gpurun_out/r02_bench_${m}_1gpu_samebox_n${N}.json
done
