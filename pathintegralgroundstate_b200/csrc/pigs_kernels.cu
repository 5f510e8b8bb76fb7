// pigs_kernels.cu -- instantiation of the persistent sweep kernel.
// Compiled once per (PIGS_INST_MT, PIGS_INST_VAR) pair; see Makefile.
#include "pigs_launch.h"

#ifndef PIGS_INST_MT
#error "define PIGS_INST_MT (0|1) and PIGS_INST_VAR (0..3)"
#endif

namespace pigs {

// One CTA per SM (tables in shared memory) of up to 1024 threads = G chain groups.
__global__ void __launch_bounds__(1024, 1)
PIGS_KNAME(const __grid_constant__ DevParams P, const __grid_constant__ SweepArgs A) {
    sweep_body<(PIGS_INST_MT != 0), PIGS_INST_VAR>(P, A);
}

cudaError_t PIGS_LNAME(int what, const DevParams* P, const SweepArgs* A, int grid, int block, size_t smem,
                       cudaStream_t st, int* out) {
    if (what == 0) {
        PIGS_KNAME<<<grid, block, smem, st>>>(*P, *A);
        return cudaGetLastError();
    }
    if (what == 1) return cudaFuncSetAttribute(PIGS_KNAME, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, PIGS_KNAME, block, smem);
}

}  // namespace pigs
