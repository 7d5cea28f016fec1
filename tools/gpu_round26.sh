#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|PROF|TL|^   |rror|esr:|Trace" | head -${LINES_MAX:-20}; }
{
for arch in rfdn rlfn; do
  run $arch f16 --size 33 47 --batch 2 --graph 1
  run $arch f16 --size 15 15 --graph 0
  run $arch f16 --size 70 201 --graph 0
done
run rfdn f16 --graph 1 --size 256 256 --time 2000 --nocheck 1
run rfdn f16 --graph 1 --size 256 256 --batch 16 --time 50 --nocheck 1
LINES_MAX=9 run rfdn f16 --size 256 256 --profile 20 --nocheck 1
LINES_MAX=9 run rfdn f16 --size 256 256 --batch 16 --profile 10 --nocheck 1
} > gpurun_out/r26.txt 2>&1
cat gpurun_out/r26.txt
