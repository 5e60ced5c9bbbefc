#!/bin/bash
# one GPU call: parity tests + default bench line (run from the repo root on the GPU box)
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time python bench.py ) > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err
tail -c 3000 gpurun_out/bench_default.log
tail -5 gpurun_out/bench_default.err
