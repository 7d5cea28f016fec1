#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|PROF|TL|^   |rror|esr:|Trace" | head -${LINES_MAX:-20}; }
{
for arch in rfdn imdn rlfn bsrn; do
  run $arch f32 --tc 0 --size 33 47 --batch 2
  run $arch f16 --size 33 47 --batch 2
done
LINES_MAX=30 run rfdn f16 --size 256 256 --profile 20
run rfdn f16 --graph 1 --size 256 256 --time 2000
run rfdn f16 --graph 1 --size 256 256 --batch 16 --time 50
run rfdn f32 --size 256 256 --time 20
} > gpurun_out/r8.txt 2>&1
cat gpurun_out/r8.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r8_pytest.txt
