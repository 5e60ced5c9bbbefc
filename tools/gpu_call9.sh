#!/bin/bash
mkdir -p gpurun_out
python tools/tokmix_check.py --all > gpurun_out/tokmix_bringup.log 2>&1; grep -E "^TOKMIX" gpurun_out/tokmix_bringup.log | tail -3 | cut -c1-400
: > gpurun_out/tokmix_flags.log
run() { echo -n "$1 | " >> gpurun_out/tokmix_flags.log; env $1 TOKMIX_ONLY=fused_fwd,fused_bwd timeout 120 python tools/bench_tokmix.py 2>&1 | tail -1 | cut -c50-200 >> gpurun_out/tokmix_flags.log; }
run "VMLP_TM_FLAGS=0"
run "VMLP_TM_FLAGS=1"
run "VMLP_TM_FLAGS=4"
run "VMLP_TM_FLAGS=31"
run "VMLP_TM_FLAGS=8"
cat gpurun_out/tokmix_flags.log
python tools/tokmix_trace.py > gpurun_out/trace_full.log 2>&1; sed -n 1,16p gpurun_out/trace_full.log; grep -n "^EPI" -A 14 gpurun_out/trace_full.log
python tools/bench_tokmix.py > gpurun_out/bench_tokmix.log 2>&1; tail -1 gpurun_out/bench_tokmix.log | cut -c1-900
