#!/bin/bash
# A/B: the same one-block loop with two builds of the library, interleaved so both see the same thermal / power state
for i in 1 2 3; do
  for so in ${AB_LIBS:-libvmlp_base.so libvmlp_b200.so}; do
    echo -n "$so: "; VMLP_LIB_PATH=$PWD/jittor-mlp_b200/$so python tools/one_block.py 30 2>&1 | tail -1
  done
done
