#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|PROF|TL|^   |rror|esr:|Trace" | head -${LINES_MAX:-20}; }
{
LINES_MAX=60 run rfdn f16 --size 256 256 --batch 16 --profile 10 --timeline 2
} > gpurun_out/r9.txt 2>&1
cat gpurun_out/r9.txt
# one full forward (48 launches) at B=1, skipping the 48 launches of the first (warm-up) forward is not possible with -s on
# kernels from torch; capture our kernels by name instead
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_esa_apply|k_head_conv|k_conv16|k_maxpool" -c 6 -o gpurun_out/r9_small python tools/gpu_check.py rfdn f16 --size 256 256 > gpurun_out/r9_ncu1.log 2>&1; tail -1 gpurun_out/r9_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" -s 23 -c 3 -o gpurun_out/r9_tc16 python tools/gpu_check.py rfdn f16 --size 256 256 --batch 16 > gpurun_out/r9_ncu2.log 2>&1; tail -1 gpurun_out/r9_ncu2.log
