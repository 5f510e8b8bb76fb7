#!/bin/bash
# round-2 experiment: in-range compaction of the partner loop (PIGS_LOOPV bit 512)
mkdir -p gpurun_out
{
for v in 33 545; do
  ./build/lb/loopbench_v${v}_t512 256 65536 1500 16
  ./build/lb/loopbench_v${v}_t512 256 8192 1500 16
  ./build/lb/loopbench_v${v}_t512 64 65536 4000 16
done
} > gpurun_out/r2_loop5.log 2>&1
cat gpurun_out/r2_loop5.log
