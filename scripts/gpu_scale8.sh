#!/bin/bash
# one 8-GPU box: C3 weak scaling and C5 strong scaling at 4 and 8 GPUs (torchrun), the reference arm on this box's
# host cores, and vpi_cuda --gpus 8 (multi-GPU inside the C ABI)
mkdir -p gpurun_out
{
nvidia-smi -L | wc -l; nproc
for N in 4 8; do for wl in C3 C5; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --no-cpu-baseline 2>gpurun_out/r2_scale_${wl}_$N.err | tail -1 > gpurun_out/bench_r02_${wl}_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_${wl}_${N}gpu.json')); print('$wl', d['n_gpus'], 'GPUs', round(d['value']/1e6,1), 'M/s e2e', round(d['e2e']['value']/1e6,1), d['scaling'], d['config'].get('schedule'), d['config'].get('chains_per_gpu'), d['clocks'])"
done; done
mkdir -p /tmp/vpirun && sed -e 's/Nblock *= *[0-9]*/Nblock = 3/' -e 's/Nstep *= *[0-9]*/Nstep = 10/' examples/vpi.in > /tmp/vpirun/vpi_small.in
./pathintegralgroundstate_b200/vpi_cuda --workdir /tmp/vpirun/g8 --chains 4096 --gpus 8 --rng philox < /tmp/vpirun/vpi_small.in | grep -E "Markov|GPU throughput" | tail -3
python bench.py --impl reference --steps 10 --warmup 2 --workload C3 > gpurun_out/bench_r02_C3_reference_arm_8gpu_box.json 2>gpurun_out/r2_ref8.err
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_C3_reference_arm_8gpu_box.json')); c=d['cpu_baseline']; print('reference arm', round(d['value']/1e6,3), 'M/s', c['cores'], 'cores', c['cpu'], 'port O2', round(c['value_port_O2']/1e6,3), 'native', round(c['value_port_native']/1e6,3), c['spread'])"
} > gpurun_out/r2_scale8.log 2>&1
cat gpurun_out/r2_scale8.log
