#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|rror|esr:|Trace" | head -20; }
{
run bsrn f16 --size 64 64
run rfdn f16 --graph 1 --size 256 256 --time 3000
run rfdn f16 --graph 0 --size 256 256 --time 3000
} > gpurun_out/r4_check.txt 2>&1
cat gpurun_out/r4_check.txt
timeout 600 python bench.py > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err; tail -3 gpurun_out/r4_bench.err; cat gpurun_out/r4_bench.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r4_pytest.txt
