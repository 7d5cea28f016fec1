#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|PROF|TL|^   |rror|esr:|Trace" | head -${LINES_MAX:-20}; }
{
for arch in rfdn imdn rlfn bsrn; do
  run $arch f32 --tc 0 --size 33 47 --batch 2 --graph 1
  run $arch f16 --size 33 47 --batch 2 --graph 1
  run $arch f16 --size 70 200 --graph 0
done
for pdl in 0 1; do
echo "== pdl $pdl"
run rfdn f16 --graph 1 --size 256 256 --time 2000 --nocheck 1 --pdl $pdl
run rfdn f16 --graph 0 --size 256 256 --time 2000 --nocheck 1 --pdl $pdl
run rfdn f16 --graph 1 --size 256 256 --batch 16 --time 50 --nocheck 1 --pdl $pdl
done
LINES_MAX=30 run rfdn f16 --size 256 256 --batch 16 --profile 10 --timeline 1 --nocheck 1
} > gpurun_out/r16.txt 2>&1
cat gpurun_out/r16.txt
