#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r21_pytest.txt
timeout 600 python bench.py > gpurun_out/r21_bench.json 2> gpurun_out/r21_bench.err; tail -3 gpurun_out/r21_bench.err; cat gpurun_out/r21_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r21_bench_ref.json 2>> gpurun_out/r21_bench.err; cat gpurun_out/r21_bench_ref.json
# launch list of the bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r21_launches.csv -k regex:"conv_tc|k_" -c 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r21_ncu_launch.log 2>&1; tail -2 gpurun_out/r21_ncu_launch.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" -s 1 -c 1 -o gpurun_out/r21_conv_tc_b1 python tools/gpu_check.py rfdn f16 --size 256 256 --nocheck 1 > gpurun_out/r21_ncu1.log 2>&1; tail -1 gpurun_out/r21_ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" -s 1 -c 1 -o gpurun_out/r21_conv_tc_b16 python tools/gpu_check.py rfdn f16 --size 256 256 --batch 16 --nocheck 1 > gpurun_out/r21_ncu2.log 2>&1; tail -1 gpurun_out/r21_ncu2.log
