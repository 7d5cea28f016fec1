#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|PROF|TL|^   |rror|esr:|Trace" | head -${LINES_MAX:-20}; }
{
LINES_MAX=40 run rfdn f16 --size 256 256 --profile 20 --nocheck 1
run rfdn f16 --graph 1 --size 256 256 --time 2000 --nocheck 1 --pdl 0
run rfdn f16 --graph 1 --size 256 256 --time 2000 --nocheck 1 --pdl 1
for arch in imdn rlfn bsrn; do
run $arch f16 --graph 1 --size 256 256 --time 500 --nocheck 1
done
run rfdn f16 --graph 1 --size 339 510 --time 200 --nocheck 1
run rfdn f16 --graph 1 --size 256 256 --batch 4 --time 200 --nocheck 1
run rfdn f16 --graph 1 --size 256 256 --batch 32 --time 20 --nocheck 1
} > gpurun_out/r20.txt 2>&1
cat gpurun_out/r20.txt
