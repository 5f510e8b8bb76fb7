#!/bin/bash
# what the driver runs at round end, in the same order: GPU tests, smoke(), reference arm, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r02_c3_reference_arm.json ) 2>&1 | grep real
( time python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r02_c3_1gpu.json ) 2>&1 | grep real
python - <<'PY'
import json
r=json.load(open('gpurun_out/bench_r02_c3_reference_arm.json')); d=json.load(open('gpurun_out/bench_r02_c3_1gpu.json'))
print('reference', round(r['value']/1e6,3), 'M/s', r['cpu_baseline']['cores'], 'cores', r['cpu_baseline']['spread'], 'port', round(r['cpu_baseline']['value_port_O2']/1e6,3), round(r['cpu_baseline']['value_port_native']/1e6,3))
print('ours', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'frac', round(d['roofline']['frac'],4), 'peak', d['roofline']['peak'], d['clocks'], 'launches', d['gpu_launches'], 'cpu_baseline', round(d['cpu_baseline']['value']/1e6,3))
print('e2e ratio', d['e2e']['value']/r['value'])
PY
