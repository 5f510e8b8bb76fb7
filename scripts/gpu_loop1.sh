#!/bin/bash
# round-2 experiment batch 1: partner-loop variants in isolation + conversion micro-benchmarks
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
./build/lb/ubench
for v in 0 1 5 9 13; do
  ./build/lb/loopbench_v$v 256 65536 1500 16
  ./build/lb/loopbench_v$v 256 8192 1500 16
  ./build/lb/loopbench_v$v 64 65536 4000 16
done
} > gpurun_out/r2_loop1.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests1.log 2>&1
tail -3 gpurun_out/r2_gputests1.log
