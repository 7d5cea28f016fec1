#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|PROF|TL|^   |rror|esr:|Trace" | head -${LINES_MAX:-20}; }
{
run rfdn f16 --size 70 200 --graph 1
run rlfn f16 --size 70 200 --graph 1
for pdl in 0 1; do
echo "== pdl $pdl"
run rfdn f16 --graph 1 --size 256 256 --time 2000 --nocheck 1 --pdl $pdl
run rfdn f16 --graph 1 --size 256 256 --batch 16 --time 50 --nocheck 1 --pdl $pdl
done
LINES_MAX=70 run rfdn f16 --size 256 256 --batch 16 --profile 10 --timeline 1 --nocheck 1
LINES_MAX=70 run rfdn f16 --size 256 256 --profile 20 --timeline 1 --nocheck 1
} > gpurun_out/r17.txt 2>&1
cat gpurun_out/r17.txt
