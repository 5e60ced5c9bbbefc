#!/bin/bash
# bring-up of the fused token-mixing kernels: every shape in its own process, bounded by timeout
mkdir -p gpurun_out
L=gpurun_out/tokmix_bringup.log
: > $L
for s in "2 16 64 64" "2 64 128 256" "3 64 128 256" "2 49 200 200" "5 80 256 136" "2 100 128 320" "2 20 128 128" "4 196 768 784" "2 196 1024 784" "1 256 384 1024" "256 196 768 784"; do
  for w in fwd bwd; do
    echo "== $s $w" >> $L
    timeout 180 python tools/tokmix_check.py $s $w >> $L 2>&1
    echo "exit $?" >> $L
  done
done
grep -E "TOKMIX|exit|timeout|error|Error" $L | tail -60
