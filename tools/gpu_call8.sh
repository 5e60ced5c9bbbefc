#!/bin/bash
mkdir -p gpurun_out
python tools/tokmix_trace.py > gpurun_out/trace_full.log 2>&1
VMLP_TM_FLAGS=31 python tools/tokmix_trace.py > gpurun_out/trace_skel.log 2>&1
head -50 gpurun_out/trace_skel.log
