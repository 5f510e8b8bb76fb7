#!/bin/bash
# round-2 experiment: two partner blocks per loop iteration (PIGS_LOOPV bit 256), partner registers carried across beads
# build first: nvcc ... -DPIGS_LOOPV={33,289} -DLB_THREADS=512 -o build/lb/loopbench_v{33,289}_t512 scripts/loopbench.cu
mkdir -p gpurun_out
{
for c in "33 512 16" "289 512 16"; do
  set -- $c
  ./build/lb/loopbench_v$1_t$2 256 65536 1500 $3
  ./build/lb/loopbench_v$1_t$2 256 8192 1500 $3
  ./build/lb/loopbench_v$1_t$2 64 65536 4000 $3
done
} > gpurun_out/r2_loop4.log 2>&1
grep mixed gpurun_out/r2_loop4.log
