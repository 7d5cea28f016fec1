#!/bin/bash
# session 4, run 1: split epilogue + chained conv launches
mkdir -p gpurun_out
(timeout 120 python tools/gpu_check.py rfdn f16 --size 64 64 --chain 0 --graph 0
 timeout 120 python tools/gpu_check.py rfdn f16 --size 64 64 --chain 1 --graph 0
 timeout 120 python tools/gpu_check.py rfdn f16 --size 256 256 --chain 0 --graph 1 --time 300
 timeout 120 python tools/gpu_check.py rfdn f16 --size 256 256 --chain 1 --graph 1 --time 300 --profile 20 --timeline 6
 timeout 120 python tools/gpu_check.py rfdn f16 --size 256 256 --batch 16 --chain 0 --graph 1 --time 50 --nocheck 1 --profile 10
 timeout 120 python tools/gpu_check.py rfdn f16 --size 256 256 --batch 16 --chain 1 --graph 1 --time 50 --nocheck 1 --profile 10 --timeline 1
) 2>&1 | grep -E "CHECK|TIME|PROF|TL|^   |rror|esr:|Trace|NOCHECK" | tee gpurun_out/s4_1_check.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/s4_1_pytest.txt
