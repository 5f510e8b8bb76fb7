#!/bin/bash
# production-shaped soak: the reference's own driver program over the C ABI at C3 scale (2 368 chains, blocks of 100
# steps; every chain starts from its own uniform random gas, vpi_mod.f90:232-236, so block 1 is a violent transient),
# a checkpointed stop and a resume, and the single-chain drop-in case in team mode; everything under `timeout`
mkdir -p gpurun_out /tmp/soak
sed -e 's/Nblock *= *[0-9]*/Nblock = 4/' examples/vpi.in > /tmp/soak/c3.in
{
echo "== C3, 2368 Philox chains, 4 blocks of 100 steps"
( cd /tmp/soak && timeout 600 /root/repo/pathintegralgroundstate_b200/vpi_cuda --workdir /tmp/soak/a --chains 2368 --rng philox < /tmp/soak/c3.in | grep -E "<E>|<Et>|GPU throughput|NaN|nan" | tail -16 )
echo "== resume: Nblock is the total, the run continues from block 5 to block 6"
sed -e 's/Nblock *= *[0-9]*/Nblock = 6/' -e 's/resume *= *F/resume = T/' examples/vpi.in > /tmp/soak/c3r.in
( cd /tmp/soak && timeout 600 /root/repo/pathintegralgroundstate_b200/vpi_cuda --workdir /tmp/soak/a --chains 2368 --rng philox < /tmp/soak/c3r.in | grep -E "<E>|<Et>|GPU throughput|NaN|nan|resum" | tail -8 )
echo "== one chain (drop-in), team schedule, 2 blocks of 100 steps"
sed -e 's/Nblock *= *[0-9]*/Nblock = 2/' examples/vpi.in > /tmp/soak/c1.in
( cd /tmp/soak && timeout 600 /root/repo/pathintegralgroundstate_b200/vpi_cuda --workdir /tmp/soak/b --chains 1 --rng philox < /tmp/soak/c1.in | grep -E "<E>|<Et>|GPU throughput|NaN|nan" | tail -6 )
echo "== one chain, MT19937 replay (the reference's own stream), 1 block of 100 steps"
sed -e 's/Nblock *= *[0-9]*/Nblock = 1/' examples/vpi.in > /tmp/soak/c1m.in
( cd /tmp/soak && timeout 900 /root/repo/pathintegralgroundstate_b200/vpi_cuda --workdir /tmp/soak/c --chains 1 --rng mt < /tmp/soak/c1m.in | grep -E "<E>|<Et>|GPU throughput|NaN|nan" | tail -4 )
grep -c . /tmp/soak/a/e_vpi.out; tail -3 /tmp/soak/a/e_vpi.out
} > gpurun_out/r2_soak.log 2>&1
cat gpurun_out/r2_soak.log
