#!/bin/bash
mkdir -p gpurun_out
python tools/tokmix_check.py --all > gpurun_out/tokmix_bringup.log 2>&1; grep -E "^TOKMIX" gpurun_out/tokmix_bringup.log | tail -3 | cut -c1-300
python tools/bench_tokmix.py > gpurun_out/bench_tokmix.log 2>&1; tail -1 gpurun_out/bench_tokmix.log
: > gpurun_out/tokmix_flags.log
for f in 0 1 2 4 8 16 9 3 31; do
  echo -n "flags $f " >> gpurun_out/tokmix_flags.log
  VMLP_TM_FLAGS=$f TOKMIX_ONLY=fused_fwd,fused_bwd timeout 120 python tools/bench_tokmix.py 2>&1 | tail -1 >> gpurun_out/tokmix_flags.log
done
cat gpurun_out/tokmix_flags.log
TOKMIX_ONLY=fused_fwd timeout 300 ncu --set full --clock-control none --import-source on -k regex:tokmix_fwd --launch-skip 3 -c 1 -o gpurun_out/r02_tokmix_fwd_v3 -f python tools/bench_tokmix.py > gpurun_out/ncu_tokmix_fwd.log 2>&1
TOKMIX_ONLY=fused_bwd timeout 300 ncu --set full --clock-control none --import-source on -k regex:tokmix_bwd --launch-skip 3 -c 1 -o gpurun_out/r02_tokmix_bwd_v3 -f python tools/bench_tokmix.py > gpurun_out/ncu_tokmix_bwd.log 2>&1
( timeout 900 python bench.py ) > gpurun_out/bench_fused.log 2> gpurun_out/bench_fused.err; tail -c 1500 gpurun_out/bench_fused.log; tail -3 gpurun_out/bench_fused.err
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
