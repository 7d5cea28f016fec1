#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|PROF|TL|^   |rror|esr:|Trace" | head -${LINES_MAX:-20}; }
{
run rfdn f16 --size 33 47 --batch 2
run imdn f16 --size 64 64
run rlfn f16 --size 33 47 --batch 2
run bsrn f16 --size 64 64
LINES_MAX=80 run rfdn f16 --size 256 256 --profile 20 --timeline 3
run rfdn f16 --graph 1 --size 256 256 --time 2000
run rfdn f16 --graph 1 --size 256 256 --batch 16 --time 50
} > gpurun_out/r7.txt 2>&1
cat gpurun_out/r7.txt
