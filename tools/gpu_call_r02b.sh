#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tokmix_check.py --all > gpurun_out/tokmix_check.log 2>&1; tail -4 gpurun_out/tokmix_check.log
timeout 300 python tools/bench_tokmix.py > gpurun_out/bench_tokmix.json 2> gpurun_out/bench_tokmix.err; cat gpurun_out/bench_tokmix.json
timeout 300 python tools/tokmix_trace.py > gpurun_out/tokmix_trace.log 2>&1; head -3 gpurun_out/tokmix_trace.log
