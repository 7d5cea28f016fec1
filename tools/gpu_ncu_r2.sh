#!/bin/bash
# round-2 ncu evidence (one GPU): launch list of the bench command, full captures of the fused chain kernel (RFDB chain
# and tail chain at batch 1) and of the remaining per-layer kernel (c5) next to it.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv -k regex:"conv_chain|conv_tc|k_" -c 400 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_ncu_launch.log 2>&1; tail -1 gpurun_out/r2_ncu_launch.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_chain" -s 8 -c 2 -f -o gpurun_out/r2_conv_chain_b1 python tools/gpu_check.py rfdn f16 --size 256 256 --nocheck 1 --time 3 > gpurun_out/r2_ncu1.log 2>&1; tail -1 gpurun_out/r2_ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" -s 5 -c 1 -f -o gpurun_out/r2_conv_tc_c5_b1 python tools/gpu_check.py rfdn f16 --size 256 256 --nocheck 1 --time 3 > gpurun_out/r2_ncu2.log 2>&1; tail -1 gpurun_out/r2_ncu2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"esa" -s 3 -c 3 -f -o gpurun_out/r2_esa_b1 python tools/gpu_check.py rfdn f16 --size 256 256 --nocheck 1 --time 3 > gpurun_out/r2_ncu3.log 2>&1; tail -1 gpurun_out/r2_ncu3.log
python tools/gpu_check.py rfdn f16 --size 256 256 --graph 1 --profile 20 --nocheck 1 2>&1 | grep -E "PROF|^   " > gpurun_out/r2_profile_b1.txt
