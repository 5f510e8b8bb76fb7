#!/bin/bash
# few-chains regime: schedule 0 (one warp per chain) against schedule 1 (team), and the full-size lines
mkdir -p gpurun_out
run() { python bench.py --steps 4 --warmup 3 --no-cpu-baseline "$@" 2>>gpurun_out/r2_team.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', '->', round(d['value']/1e6,1), 'M/s  e2e', round(d['e2e']['value']/1e6,1), ' frac', round(d['roofline']['frac'],4), d['config'].get('schedule'), d['config'].get('chains_per_gpu'))"; }
{
run --workload C5 --chains 512 --schedule 0 --mc-steps 20
run --workload C5 --chains 512 --schedule 1 --mc-steps 20
run --workload C5 --chains 592 --schedule 1 --mc-steps 20
run --workload C5 --chains 4096 --mc-steps 8
run --workload C5 --chains 2048 --mc-steps 8
run --workload C5 --chains 1024 --mc-steps 10 --schedule 0
run --workload C5 --chains 1024 --mc-steps 10 --schedule 1
run --workload C3 --chains 512 --schedule 0 --mc-steps 4
run --workload C3 --chains 512 --schedule 1 --mc-steps 4
run --workload C3 --chains 1 --schedule 0 --mc-steps 8
run --workload C3 --chains 1 --schedule 1 --mc-steps 8
run --workload C2 --mc-steps 8
} > gpurun_out/r2_team.log 2>&1
cat gpurun_out/r2_team.log; tail -5 gpurun_out/r2_team.err
