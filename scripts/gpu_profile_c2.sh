#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/prof_case.py c2 2368 32 2 2 1"
$CMD > gpurun_out/r02_c2_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_sweep_0_2 -s 2 -c 1 -o gpurun_out/prof_r02_c2 -f $CMD > gpurun_out/r02_c2_ncu.log 2>&1
cat gpurun_out/r02_c2_plain.log; tail -2 gpurun_out/r02_c2_ncu.log
