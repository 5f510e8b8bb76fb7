#!/bin/bash
# N-GPU batch (gpurun --gpus N): the multi-GPU handle of the C ABI on real devices, vpi_cuda --gpus N, and the
# torchrun bench lines (C3 weak scaling, C5 strong scaling)
N=${1:-2}
mkdir -p gpurun_out
{
nvidia-smi -L
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_gpu_handle" 2>&1 | tail -3
mkdir -p /tmp/vpirun && cd /tmp/vpirun && sed -e 's/Nblock *= *[0-9]*/Nblock = 3/' -e 's/Nstep *= *[0-9]*/Nstep = 10/' $GRAFT_REPO_ROOT/examples/vpi.in > vpi_small.in
$GRAFT_REPO_ROOT/pathintegralgroundstate_b200/vpi_cuda --workdir /tmp/vpirun/g$N --chains 1024 --gpus $N --rng philox < vpi_small.in | grep -E "Markov|<E>  =|GPU throughput" | tail -6
$GRAFT_REPO_ROOT/pathintegralgroundstate_b200/vpi_cuda --workdir /tmp/vpirun/g1 --chains 1024 --gpus 1 --rng philox < vpi_small.in | grep -E "Markov|<E>  =|GPU throughput" | tail -6
cmp /tmp/vpirun/g$N/e_vpi.out /tmp/vpirun/g1/e_vpi.out && echo "e_vpi.out identical on 1 and $N GPUs"
cd $GRAFT_REPO_ROOT
for wl in C3 C5; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --no-cpu-baseline 2>gpurun_out/r2_multi_$wl.err | tail -1 > gpurun_out/bench_r02_${wl}_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r02_${wl}_${N}gpu.json')); print('$wl', d['n_gpus'], 'GPUs', round(d['value']/1e6,1), 'M/s e2e', round(d['e2e']['value']/1e6,1), d['scaling'], d['config'].get('schedule'), d['config'].get('chains_per_gpu'))"
done
} > gpurun_out/r2_multi_$N.log 2>&1
cat gpurun_out/r2_multi_$N.log
