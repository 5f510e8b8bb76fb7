#!/bin/bash
# round-2 profile of the benchmarked command (C3, default bench shape, 2 MC steps per launch to keep ncu short)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --mc-steps 2 --no-cpu-baseline"
$CMD > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/ncu_r02_launches_bench_c3.csv $CMD > gpurun_out/r02_ncu1.log 2>&1
$CMD > gpurun_out/r02_plain2.json 2>> gpurun_out/r02_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:k_sweep_0_2 -s 4 -c 1 -o gpurun_out/prof_r02_c3 -f $CMD > gpurun_out/r02_ncu2.log 2>&1
ls -la gpurun_out/prof_r02_c3.ncu-rep; tail -2 gpurun_out/r02_ncu2.log; cat gpurun_out/r02_plain.json | cut -c1-300
