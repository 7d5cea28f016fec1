#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/r18_pytest.txt
timeout 600 python bench.py > gpurun_out/r18_bench.json 2> gpurun_out/r18_bench.err; tail -3 gpurun_out/r18_bench.err; cat gpurun_out/r18_bench.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
