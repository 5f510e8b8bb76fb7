#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline "$@" 2>>gpurun_out/r2_quick.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$PIGS_CUDA_LIB $PIGS_PREFETCH $*', '->', round(d['value']/1e6,1), 'M/s  e2e', round(d['e2e']['value']/1e6,1), ' frac', round(d['roofline']['frac'],4), d['config'].get('schedule'), d['config'].get('chains_per_gpu'))"; }
{
run --workload C3
run --workload C2 --mc-steps 8
run --workload C5 --chains 4096 --mc-steps 8
run --workload C5 --chains 512 --schedule 1 --mc-steps 20
PIGS_PREFETCH=0 run --workload C3
} > gpurun_out/r2_quick.log 2>&1
cat gpurun_out/r2_quick.log; tail -3 gpurun_out/r2_quick.err
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
