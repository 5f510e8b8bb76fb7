// pigs_launch.h -- host-visible launchers of the kernels (one translation unit
// per sweep-kernel instantiation so they compile in parallel).  Every launcher
// uploads its translation unit's __constant__ parameter block on the same
// stream right before the launch; the C API serialises launches with a mutex.
#pragma once
#include "pigs_device.cuh"

namespace pigs {

// persistent sweep kernel; mt: 0 Philox / 1 MT19937 replay; var: see VarTraits
cudaError_t launch_sweep(int mt, int var, const DevParams& P, const SweepArgs& A, int grid, int block, size_t smem,
                         cudaStream_t st);
cudaError_t sweep_set_smem(int mt, int var, size_t smem);
cudaError_t sweep_occupancy(int mt, int var, int block, size_t smem, int* ctas_per_sm);
cudaError_t sweep_max_threads(int mt, int var, int* n);      // the instantiation's __launch_bounds__

// unit kernels (pigs_unit.cu)
enum UnitOp { U_LOCAL_ENERGY = 0, U_THERM_ENERGY = 1, U_PAIR_CORR = 2, U_SOFK = 3, U_OBDM = 4 };
struct UnitArgs {
    int op, n;
    const double* in;    // blocked SoA slices [n][NpS/32][3][32] / paths [n][S][NpS/32][3][32] / xend [n][2][3]
    double* out;         // [n][3] energies / histograms [n][...]
};
cudaError_t launch_unit(bool trap, const DevParams& P, const UnitArgs& A, cudaStream_t st);
cudaError_t launch_update_action(bool trap, bool smem_tables, const DevParams& P, int n, const double* Rsoa, const int* ip, const int* ib,
                                 const double* xnew, const double* xold, double* dS, cudaStream_t st);
// layout transposes between the ABI's Path(dim,Np,0:2Nb) and the internal SoA
cudaError_t launch_aos_to_soa(const DevParams& P, const double* aos, double* soa, int nchain, cudaStream_t st, int* flag);
cudaError_t launch_soa_to_aos(const DevParams& P, const double* soa, double* aos, int nchain, cudaStream_t st);
// chain-summed block vector [NE | NCNT | gr | Sk | nrho]
cudaError_t launch_reduce_block(const DevParams& P, double* vec, cudaStream_t st);
cudaError_t launch_zero_block(const DevParams& P, cudaStream_t st);
cudaError_t launch_dfma_peak(int blocks, int threads, int iters, double* sink, cudaStream_t st);

}  // namespace pigs
