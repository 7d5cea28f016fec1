#!/bin/bash
# first GPU contact: generic fp32/fp16 parity for all archs, then the tcgen05 path in both shift modes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r1_gpu.txt 2>&1
ls /root/reference > gpurun_out/r1_refdir.txt 2>&1
nproc >> gpurun_out/r1_gpu.txt; lscpu | grep "Model name" >> gpurun_out/r1_gpu.txt
run() { timeout 180 python tools/gpu_check.py "$@" 2>&1 | grep -E "CHECK|TIME|rror|esr:|Trace" | head -20; }
{
for arch in rfdn imdn rlfn bsrn; do
  run $arch f32 --tc 0 --size 33 47 --batch 2
  run $arch f16 --tc 0 --size 33 47 --batch 2
done
run rfdn f16 --tc 1 --shift 1 --size 64 64
run rfdn f16 --tc 1 --shift 0 --size 64 64
run rfdn f16 --tc 1 --shift 1 --size 33 47 --batch 2
run rfdn f16 --tc 1 --shift 1 --size 200 300
run imdn f16 --tc 1 --shift 1 --size 64 64
run rlfn f16 --tc 1 --shift 1 --size 64 64
run bsrn f16 --tc 1 --shift 1 --size 64 64
run rfdn f16 --tc 0 --size 256 256 --time 20
run rfdn f32 --tc 0 --size 256 256 --time 5
run rfdn f16 --tc 1 --shift 1 --size 256 256 --time 50
run rfdn f16 --tc 1 --shift 1 --graph 1 --size 256 256 --time 50
run rfdn f16 --tc 1 --shift 1 --host 1 --size 64 64
} > gpurun_out/r1_check.txt 2>&1
cat gpurun_out/r1_check.txt
