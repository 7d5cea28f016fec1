#!/bin/bash
mkdir -p gpurun_out
run() { timeout 180 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|rror|esr:|Trace" | head -20; }
{
run rfdn f16 --tc 1 --shift 1 --size 64 64
run rfdn f16 --tc 1 --shift 0 --size 64 64
run rfdn f16 --tc 1 --shift 1 --size 33 47 --batch 2
run rfdn f16 --tc 1 --shift 1 --size 200 300
run imdn f16 --tc 1 --shift 1 --size 64 64
run rlfn f16 --tc 1 --shift 1 --size 64 64
run bsrn f16 --tc 1 --shift 1 --size 64 64
run rfdn f16 --tc 1 --shift 1 --size 256 256 --time 50
run rfdn f16 --tc 1 --shift 1 --graph 1 --size 256 256 --time 50
run rfdn f16 --tc 1 --shift 1 --host 1 --size 64 64
} > gpurun_out/r2_check.txt 2>&1
cat gpurun_out/r2_check.txt
