#!/bin/bash
# GPU regression batch: parity suite, then the default bench (C3) without the CPU leg
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1
tail -3 gpurun_out/r2_gputests.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
cat gpurun_out/r2_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['clocks'])"
