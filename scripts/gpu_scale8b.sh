#!/bin/bash
# one 8-GPU box, final build: C3 weak scaling and C5 strong scaling at 1 / 2 / 4 / 8 GPUs (torchrun), same box for every N
mkdir -p gpurun_out
{
nvidia-smi -L | wc -l; nproc
for N in 1 2 4 8; do for wl in C3 C5; do
if [ $N = 1 ]; then
python bench.py --gpus 1 --steps 5 --warmup 3 --workload $wl --no-cpu-baseline 2>gpurun_out/r2_scale_${wl}_$N.err | tail -1 > gpurun_out/bench_r02_${wl}_${N}gpu.json
else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --no-cpu-baseline 2>gpurun_out/r2_scale_${wl}_$N.err | tail -1 > gpurun_out/bench_r02_${wl}_${N}gpu.json
fi
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_${wl}_${N}gpu.json')); print('$wl', d['n_gpus'], 'GPUs', round(d['value']/1e6,1), 'M/s e2e', round(d['e2e']['value']/1e6,1), d['scaling'], d['config'].get('schedule'), d['config'].get('chains_per_gpu'), d['clocks'])"
done; done
} > gpurun_out/r2_scale8b.log 2>&1
cat gpurun_out/r2_scale8b.log
