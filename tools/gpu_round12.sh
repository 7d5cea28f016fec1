#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|PROF|TL|^   |rror|esr:|Trace" | head -${LINES_MAX:-20}; }
{
for arch in rfdn imdn rlfn bsrn; do
  run $arch f16 --size 33 47 --batch 2
done
LINES_MAX=70 run rfdn f16 --size 256 256 --profile 20 --timeline 2
run rfdn f16 --graph 1 --size 256 256 --time 2000
LINES_MAX=70 run rfdn f16 --size 256 256 --batch 16 --profile 10 --timeline 1
run rfdn f16 --graph 1 --size 256 256 --batch 16 --time 50
} > gpurun_out/r12.txt 2>&1
cat gpurun_out/r12.txt
