#!/bin/bash
# round-2 evidence run (one GPU): tests, smoke, both bench arms (default config), the other BASELINE configs, FMEN, the ncu
# launch list of the bench command and full captures of the fused chain / c5 / ESA kernels.  Everything lands in
# gpurun_out/r2f_*; the summaries worth keeping are copied to profiles/ afterwards.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2f_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2f_smoke.txt
timeout 600 python bench.py > gpurun_out/r2f_bench_config1.json 2> gpurun_out/r2f_bench.err; tail -2 gpurun_out/r2f_bench.err; cut -c1-260 gpurun_out/r2f_bench_config1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2f_bench_reference_arm.json 2>> gpurun_out/r2f_bench.err; cut -c1-200 gpurun_out/r2f_bench_reference_arm.json
for c in 0 2 3 4; do timeout 400 python bench.py --config $c --steps 30 > gpurun_out/r2f_bench_config$c.json 2>> gpurun_out/r2f_bench.err; cut -c1-160 gpurun_out/r2f_bench_config$c.json; done
timeout 300 python bench.py --model fmen --steps 200 > gpurun_out/r2f_bench_fmen_b1.json 2>> gpurun_out/r2f_bench.err
timeout 300 python bench.py --model imdn --steps 200 > gpurun_out/r2f_bench_imdn_b1.json 2>> gpurun_out/r2f_bench.err
timeout 300 python bench.py --model rlfn --steps 200 > gpurun_out/r2f_bench_rlfn_b1.json 2>> gpurun_out/r2f_bench.err
bash tools/gpu_ncu_r2.sh
