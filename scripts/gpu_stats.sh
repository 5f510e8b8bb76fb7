#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_driver.py tests/test_host.py -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1
tail -5 gpurun_out/r2_gputests.log
python -m pytest tests/test_gpu_stats.py -m gpu -q -s > gpurun_out/r2_stats.log 2>&1
tail -60 gpurun_out/r2_stats.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
cat gpurun_out/r2_bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C3', d['value']/1e6, d['e2e']['value']/1e6, d['roofline']['frac'], d['clocks'])"
tail -3 gpurun_out/r2_bench.err
