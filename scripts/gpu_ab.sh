#!/bin/bash
# A/B of library builds: usage gpu_ab.sh <tag> [<tag> ...]; tag "" = the shipped libpigs_cuda.so
mkdir -p gpurun_out
run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline "$@" 2>>gpurun_out/r2_ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('${TAG:-base} $*', '->', round(d['value']/1e6,1), 'M/s  e2e', round(d['e2e']['value']/1e6,1), ' frac', round(d['roofline']['frac'],4), d['config'].get('schedule'), d['config'].get('chains_per_gpu'))"; }
{
for TAG in "" "$@"; do
  if [ -n "$TAG" ]; then export PIGS_CUDA_LIB=$PWD/pathintegralgroundstate_b200/libpigs_cuda_$TAG.so; else unset PIGS_CUDA_LIB; fi
  run --workload C3
  run --workload C2 --mc-steps 8
  run --workload C5 --chains 4096 --mc-steps 8
  run --workload C5 --chains 512 --mc-steps 20
  run --workload C3 --chains 512 --mc-steps 4
done
} > gpurun_out/r2_ab.log 2>&1
cat gpurun_out/r2_ab.log; tail -3 gpurun_out/r2_ab.err
