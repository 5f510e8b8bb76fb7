#!/bin/bash
mkdir -p gpurun_out
{
for v in 0 1 17 33 49 53; do
  ./build/lb/loopbench_v$v 256 65536 1500 16
  ./build/lb/loopbench_v$v 256 8192 1500 16
  ./build/lb/loopbench_v$v 64 65536 4000 16
done
} > gpurun_out/r2_loop2.log 2>&1
tail -4 gpurun_out/r2_loop2.log
