// pigs_kernels.cu -- instantiation of the persistent sweep kernel.
// Compiled once per (PIGS_INST_MT, PIGS_INST_VAR) pair; see Makefile.
#include "pigs_sweep.cuh"

#ifndef PIGS_INST_MT
#error "define PIGS_INST_MT (0|1) and PIGS_INST_VAR (0..3)"
#endif
#ifndef PIGS_MAXT
#define PIGS_MAXT 512
#endif

namespace pigs {

// One CTA per SM (tables in shared memory) of up to PIGS_MAXT threads = G chain groups.
// PIGS_MAXT = 512 gives the move engine 128 registers per thread: measured on
// B200, 16 spill-free warps per SM beat 24 (80 regs) and 32 (64 regs, spilling
// in the partner loop) for both the N=64 and the N=256 workloads.
#ifdef PIGS_MAXNREG
// tuning builds: an explicit register budget (the launcher still starts at most PIGS_MAXT threads per CTA)
__global__ void __maxnreg__(PIGS_MAXNREG) PIGS_KNAME() { sweep_body<(PIGS_INST_MT != 0), PIGS_INST_VAR>(); }
#else
__global__ void __launch_bounds__(PIGS_MAXT, 1) PIGS_KNAME() { sweep_body<(PIGS_INST_MT != 0), PIGS_INST_VAR>(); }
#endif

cudaError_t PIGS_LNAME(int what, const DevParams* P, const SweepArgs* A, int grid, int block, size_t smem,
                       cudaStream_t st, int* out) {
    if (what == 0) {
        // the attribute belongs to the function (per device), not to a handle: two handles with different
        // shared-memory footprints share this kernel, so it is set for every launch
        cudaError_t e = cudaFuncSetAttribute(PIGS_KNAME, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyToSymbolAsync(cP, P, sizeof(DevParams), 0, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyToSymbolAsync(cA, A, sizeof(SweepArgs), 0, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        PIGS_KNAME<<<grid, block, smem, st>>>();
        return cudaGetLastError();
    }
    if (what == 1) return cudaFuncSetAttribute(PIGS_KNAME, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (what == 3) { *out = PIGS_MAXT; return cudaSuccess; }
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(out, PIGS_KNAME, block, smem);
}

}  // namespace pigs
