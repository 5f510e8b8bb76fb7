#!/bin/bash
mkdir -p gpurun_out
timeout 60 ./build/lb/tma_test > gpurun_out/r2_tma_test.log 2>&1; cat gpurun_out/r2_tma_test.log
# shared-memory race check of the team schedule (4 window workers per chain): 6 chains of the small worm configuration
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/prof_case.py cw 6 0 -1 3 1 > gpurun_out/r2_racecheck_team.log 2>&1
tail -6 gpurun_out/r2_racecheck_team.log
