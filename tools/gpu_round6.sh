#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 2 4; do
  echo "=== dbg_flags=$dbg"
  python tools/gpu_check.py rfdn f16 --size 256 256 --profile 20 --timeline 1 --dbg $dbg 2>&1 | grep -E "CHECK|PROF conv_tc:B1|PROF conv_tc:c |PROF conv_tc:LR|PROF conv_tc:up|TL|   |rror" | head -24
done > gpurun_out/r6_prof.txt 2>&1
cat gpurun_out/r6_prof.txt
