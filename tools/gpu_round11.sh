#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" -s 1 -c 2 -o gpurun_out/r11_tc16 python tools/gpu_check.py rfdn f16 --size 256 256 --batch 16 > gpurun_out/r11_ncu.log 2>&1; tail -2 gpurun_out/r11_ncu.log
ls -la gpurun_out/*.ncu-rep
