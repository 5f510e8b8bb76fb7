#!/bin/bash
mkdir -p gpurun_out
{
for v in 33 97 161; do
  ./build/lb/loopbench_v$v 256 65536 1500 16
  ./build/lb/loopbench_v$v 256 8192 1500 16
  ./build/lb/loopbench_v$v 64 65536 4000 16
done
} > gpurun_out/r2_loop3.log 2>&1
grep mixed gpurun_out/r2_loop3.log
