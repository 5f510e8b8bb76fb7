#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline "$@" 2>>gpurun_out/r2_pf.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('PF=$PIGS_PREFETCH $*', '->', round(d['value']/1e6,1), 'M/s  e2e', round(d['e2e']['value']/1e6,1), d['config'].get('schedule'), d['config'].get('chains_per_gpu'))"; }
{
for pf in 3 0; do
export PIGS_PREFETCH=$pf
run --workload C2 --mc-steps 8
run --workload C5 --chains 4096 --mc-steps 8
run --workload C5 --chains 512 --mc-steps 20
run --workload C5 --chains 1024 --mc-steps 10
run --workload C3 --chains 512 --mc-steps 4
done
} > gpurun_out/r2_pf.log 2>&1
cat gpurun_out/r2_pf.log
