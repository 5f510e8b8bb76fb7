#!/bin/bash
# prefetch distance sweep (PIGS_PFDIST: how many evaluated slices ahead the rolling L2 bulk prefetch runs)
mkdir -p gpurun_out
run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline "$@" 2>>gpurun_out/r2_pfd.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pfdist=$PIGS_PFDIST $*', '->', round(d['value']/1e6,1), 'M/s  e2e', round(d['e2e']['value']/1e6,1), ' frac', round(d['roofline']['frac'],4))"; }
{
for d in 1 2 3 4 6; do
  export PIGS_PFDIST=$d
  run --workload C3
  run --workload C2 --mc-steps 8
done
} > gpurun_out/r2_pfdist.log 2>&1
cat gpurun_out/r2_pfdist.log
