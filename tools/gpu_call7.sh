#!/bin/bash
mkdir -p gpurun_out
python tools/tokmix_check.py --all > gpurun_out/tokmix_bringup.log 2>&1; grep -E "^TOKMIX" gpurun_out/tokmix_bringup.log | tail -2 | cut -c1-300
: > gpurun_out/tokmix_flags.log
run() { echo -n "$1 | " >> gpurun_out/tokmix_flags.log; env $1 TOKMIX_ONLY=fused_fwd,fused_bwd timeout 120 python tools/bench_tokmix.py 2>&1 | tail -1 | cut -c50-200 >> gpurun_out/tokmix_flags.log; }
run "VMLP_TM_FLAGS=0"
run "VMLP_TM_FLAGS=64"
run "VMLP_TM_FLAGS=0 VMLP_TM_DEPTH=2 VMLP_TM_NHB=2"
run "VMLP_TM_FLAGS=0 VMLP_TM_DEPTH=2 VMLP_TM_NHB=1"
run "VMLP_TM_FLAGS=31"
run "VMLP_TM_FLAGS=63"
run "VMLP_TM_FLAGS=95"
run "VMLP_TM_FLAGS=127"
run "VMLP_TM_FLAGS=4"
run "VMLP_TM_FLAGS=5"
cat gpurun_out/tokmix_flags.log
python tools/bench_tokmix.py > gpurun_out/bench_tokmix.log 2>&1; tail -1 gpurun_out/bench_tokmix.log | cut -c1-900
