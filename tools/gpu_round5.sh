#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_check.py rfdn f16 --size 256 256 --profile 20 --timeline 6 2>&1 | grep -E "CHECK|PROF|TL|   |rror" > gpurun_out/r5_prof.txt
cat gpurun_out/r5_prof.txt | head -90
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 2 -o gpurun_out/r5_conv_tc python tools/gpu_check.py rfdn f16 --size 256 256 > gpurun_out/r5_ncu.log 2>&1; tail -2 gpurun_out/r5_ncu.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r5_pytest.txt
