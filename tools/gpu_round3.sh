#!/bin/bash
mkdir -p gpurun_out
run() { timeout 180 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|rror|esr:|Trace" | head -20; }
{
run rfdn f16 --size 33 47 --batch 2
run rfdn f16 --size 200 300
run imdn f16 --size 64 64
run imdn f16 --size 33 47 --batch 2
run rlfn f16 --size 64 64
run rlfn f16 --size 33 47 --batch 2
run bsrn f16 --size 64 64
run bsrn f16 --size 33 47 --batch 2
run rfdn f16 --host 1 --size 64 64
run rfdn f16 --graph 1 --size 256 256 --time 50
run imdn f16 --graph 1 --size 256 256 --time 50
run rlfn f16 --graph 1 --size 256 256 --time 50
run bsrn f16 --graph 1 --size 256 256 --time 50
run rfdn f16 --graph 1 --size 256 256 --batch 16 --time 10
} > gpurun_out/r3_check.txt 2>&1
cat gpurun_out/r3_check.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launches.csv python tools/gpu_check.py rfdn f16 --size 256 256 > gpurun_out/r3_ncu.log 2>&1
tail -3 gpurun_out/r3_ncu.log
