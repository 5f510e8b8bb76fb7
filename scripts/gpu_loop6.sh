#!/bin/bash
# round-2 experiment: minimum image with FRND.F64 (conversion pipe) instead of the 2^52 trick
mkdir -p gpurun_out
{
for v in base frnd; do
  ./build/lb/loopbench_$v 256 65536 1500 16
  ./build/lb/loopbench_$v 256 8192 1500 16
  ./build/lb/loopbench_$v 64 65536 4000 16
done
} > gpurun_out/r2_loop6.log 2>&1
grep mixed gpurun_out/r2_loop6.log
