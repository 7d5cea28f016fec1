#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|PROF conv_tc:B1|TL|^   |rror|esr:|Trace" | head -${LINES_MAX:-20}; }
{
for slots in 2 4; do
echo "== slots $slots"
LINES_MAX=14 run rfdn f16 --size 256 256 --batch 16 --profile 10 --timeline 1 --slots $slots
echo "== slots $slots no-MMA"
LINES_MAX=14 run rfdn f16 --size 256 256 --batch 16 --profile 10 --timeline 1 --slots $slots --dbg 1
done
} > gpurun_out/r13.txt 2>&1
cat gpurun_out/r13.txt
