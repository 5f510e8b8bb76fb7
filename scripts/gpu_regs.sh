#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline "$@" 2>>gpurun_out/r2_regs.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$(basename ${PIGS_CUDA_LIB:-default}) $*', '->', round(d['value']/1e6,1), 'M/s  e2e', round(d['e2e']['value']/1e6,1), ' frac', round(d['roofline']['frac'],4), d['config'].get('chains_per_gpu'))"; }
{
run --workload C3
PIGS_CUDA_LIB=$PWD/pathintegralgroundstate_b200/libpigs_cuda_544.so run --workload C3 --chains 2516
PIGS_CUDA_LIB=$PWD/pathintegralgroundstate_b200/libpigs_cuda_576.so run --workload C3 --chains 2664
PIGS_CUDA_LIB=$PWD/pathintegralgroundstate_b200/libpigs_cuda_608.so run --workload C3 --chains 2812
PIGS_CUDA_LIB=$PWD/pathintegralgroundstate_b200/libpigs_cuda_544.so run --workload C2 --chains 2516 --mc-steps 8
PIGS_CUDA_LIB=$PWD/pathintegralgroundstate_b200/libpigs_cuda_576.so run --workload C2 --chains 2664 --mc-steps 8
} > gpurun_out/r2_regs.log 2>&1
cat gpurun_out/r2_regs.log
