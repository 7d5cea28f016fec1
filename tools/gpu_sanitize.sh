#!/bin/bash
# compute-sanitizer passes over a reduced set of forwards: every network, fp16 (tcgen05 path) and
# fp32 (CUDA-core path), small odd shapes so that border tiles, partial strips and multi-item CTAs are exercised.
# Summaries land in gpurun_out/sanitize_<tool>.txt; copy them to profiles/ to keep.
mkdir -p gpurun_out
TAG=${1:-r2}
for tool in memcheck racecheck synccheck initcheck; do
  out=gpurun_out/sanitize_${TAG}_${tool}.txt
  : > $out
  for cfg in "rfdn f16 --size 40 150" "rfdn f16 --size 33 47 --batch 3 --graph 0" "rlfn f16 --size 36 131" "bsrn f16 --size 31 140" "imdn f16 --size 24 129" "fmen f16 --size 21 135" "rfdn f32 --size 20 33"; do
    echo "=== $tool: gpu_check.py $cfg" >> $out
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python tools/gpu_check.py $cfg --nocheck 1 2>&1 \
      | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|NOCHECK|Invalid|Uninitialized|Barrier" | head -30 >> $out
    echo "exit=$?" >> $out
  done
  tail -n 3 $out
done
