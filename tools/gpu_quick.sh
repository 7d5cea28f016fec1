#!/bin/bash
# developer loop on the GPU box: parity suite, then timing / per-launch profile / timeline of RFDN at batch 1 and 16
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/quick_pytest.txt
(timeout 120 python tools/gpu_check.py rfdn f16 --size 256 256 --graph 1 --time 300 --profile 20 --timeline 2
 timeout 120 python tools/gpu_check.py rfdn f16 --size 256 256 --batch 16 --graph 1 --time 50 --nocheck 1 --profile 10 --timeline 1
) 2>&1 | grep -E "CHECK|TIME|PROF|TL|^   |rror|esr:|NOCHECK" | tee gpurun_out/quick_check.txt
