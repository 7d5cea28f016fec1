#!/bin/bash
# round-end evidence run (one GPU): tests, smoke, both bench arms, ncu launch list of the bench command, ncu full
# captures of the dominant kernel, per-launch CUDA-event profiles and in-kernel timelines, multi-stream probe.
# Everything lands in gpurun_out/final_*; the summaries worth keeping are copied to profiles/ afterwards.
mkdir -p gpurun_out
python ntire2022_esr_b200/build.py > /dev/null
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/final_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.txt
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -3 gpurun_out/final_bench.err; cat gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_bench_ref.json 2>> gpurun_out/final_bench.err; cat gpurun_out/final_bench_ref.json
timeout 600 python bench.py --batch 16 --steps 60 --warmup 5 > gpurun_out/final_bench_b16.json 2>> gpurun_out/final_bench.err; cut -c1-400 gpurun_out/final_bench_b16.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv -k regex:"conv_tc|k_" -c 400 python bench.py --steps 2 --warmup 3 > gpurun_out/final_ncu_launch.log 2>&1; tail -1 gpurun_out/final_ncu_launch.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" -s 1 -c 1 -f -o gpurun_out/final_conv_tc_b1 python tools/gpu_check.py rfdn f16 --size 256 256 --nocheck 1 > gpurun_out/final_ncu1.log 2>&1; tail -1 gpurun_out/final_ncu1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc" -s 1 -c 1 -f -o gpurun_out/final_conv_tc_b16 python tools/gpu_check.py rfdn f16 --size 256 256 --batch 16 --nocheck 1 > gpurun_out/final_ncu2.log 2>&1; tail -1 gpurun_out/final_ncu2.log
python tools/gpu_check.py rfdn f16 --size 256 256 --graph 1 --profile 20 --timeline 2 --nocheck 1 2>&1 | grep -E "PROF|TL|^   " > gpurun_out/final_profile_b1.txt
python tools/gpu_check.py rfdn f16 --size 256 256 --batch 16 --graph 1 --profile 10 --timeline 1 --nocheck 1 2>&1 | grep -E "PROF|TL|^   " > gpurun_out/final_profile_b16.txt
python tools/gpu_two_streams.py 1 2>&1 | tail -4 | tee gpurun_out/final_streams.txt
timeout 600 python tools/gpu_configs.py 2>/dev/null | tee gpurun_out/final_configs.jsonl | cut -c1-200
