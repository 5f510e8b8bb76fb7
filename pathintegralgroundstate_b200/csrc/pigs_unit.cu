// pigs_unit.cu -- unit-API kernels (one reference procedure per launch, on
// caller data), layout transposes, the chain reduction of block accumulators
// and the FP64 peak micro-benchmark.  They reuse the device functions of the
// persistent sweep kernel, so a unit-parity test exercises the production code.
#include "pigs_launch.h"
#include "pigs_sweep.cuh"
#include <cstring>

namespace pigs {

// dispatcher over the 8 separately compiled sweep instantiations
#define DECL(mt, var) cudaError_t sweep_l_##mt##_##var(int, const DevParams*, const SweepArgs*, int, int, size_t, cudaStream_t, int*);
DECL(0, 0) DECL(0, 1) DECL(0, 2) DECL(0, 3) DECL(1, 0) DECL(1, 1) DECL(1, 2) DECL(1, 3)
#undef DECL
typedef cudaError_t (*sweep_fn)(int, const DevParams*, const SweepArgs*, int, int, size_t, cudaStream_t, int*);
static sweep_fn pick(int mt, int var) {
    static const sweep_fn tab[2][4] = {{sweep_l_0_0, sweep_l_0_1, sweep_l_0_2, sweep_l_0_3},
                                       {sweep_l_1_0, sweep_l_1_1, sweep_l_1_2, sweep_l_1_3}};
    return tab[mt ? 1 : 0][var & 3];
}
cudaError_t launch_sweep(int mt, int var, const DevParams& P, const SweepArgs& A, int grid, int block, size_t smem, cudaStream_t st) {
    return pick(mt, var)(0, &P, &A, grid, block, smem, st, nullptr);
}
cudaError_t sweep_set_smem(int mt, int var, size_t smem) { return pick(mt, var)(1, nullptr, nullptr, 0, 0, smem, 0, nullptr); }
cudaError_t sweep_occupancy(int mt, int var, int block, size_t smem, int* n) { return pick(mt, var)(2, nullptr, nullptr, 0, block, smem, 0, n); }
cudaError_t sweep_max_threads(int mt, int var, int* n) { return pick(mt, var)(3, nullptr, nullptr, 0, 0, 0, 0, n); }

// ---------------------------------------------------------------- estimators on caller data
// one CTA (= one chain group of 128 threads) per configuration
template <int VAR>
__global__ void __launch_bounds__(128) k_unit(const __grid_constant__ UnitArgs A) {
    extern __shared__ __align__(16) double smem[];
    GS* gs = reinterpret_cast<GS*>(smem);
    if (threadIdx.x == 0) { gs->tabV = cP.vtab; gs->tabW = cP.logwf; gs->chain = 0; gs->gsize = 128; gs->gshift = 7; gs->gbar = 0; }
    __syncthreads();
    const size_t ss = (size_t)3 * cP.NpS;
    for (int n = blockIdx.x; n < A.n; n += gridDim.x) {
        if (A.op == U_LOCAL_ENERGY) {
            double e[3];
            LocalEnergy<VAR>(gs, A.in + n * ss, e);
            if (threadIdx.x == 0) { A.out[3 * n] = e[0]; A.out[3 * n + 1] = e[1]; A.out[3 * n + 2] = e[2]; }
        } else if (A.op == U_THERM_ENERGY) {
            double e[3];
            if (threadIdx.x == 0) gs->path = const_cast<double*>(A.in) + n * ss * cP.S;
            __syncthreads();
            ThermEnergy<VAR>(gs, e);
            if (threadIdx.x == 0) { A.out[3 * n] = e[0]; A.out[3 * n + 1] = e[1]; A.out[3 * n + 2] = e[2]; }
        } else if (A.op == U_PAIR_CORR) {
            PairCorrelation(gs, A.in + n * ss, A.out + (size_t)n * cP.Nbin);
        } else if (A.op == U_SOFK) {
            StructureFactor(gs, A.in + n * ss, A.out + (size_t)n * cP.Nk * cP.dim);
        } else if (A.op == U_OBDM) {
            OBDM(gs, A.in + (size_t)n * 6, A.out + (size_t)n * cP.Nbin * (cP.Npw + 1));
        }
        __syncthreads();
    }
}
static cudaError_t upload(const DevParams& P, int T, cudaStream_t st) {
    SweepArgs A;
    memset(&A, 0, sizeof A);
    A.chain_only = -1; A.groups_per_cta = 1; A.threads_per_chain = T;
    A.tshift = 0; while ((1 << A.tshift) < T) ++A.tshift;
    cudaError_t e = cudaMemcpyToSymbolAsync(cP, &P, sizeof(DevParams), 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbolAsync(cA, &A, sizeof(SweepArgs), 0, cudaMemcpyHostToDevice, st);
}
cudaError_t launch_unit(bool trap, const DevParams& P, const UnitArgs& A, cudaStream_t st) {
    cudaError_t e = upload(P, 128, st);
    if (e != cudaSuccess) return e;
    int grid = A.n < 1184 ? A.n : 1184;
    size_t sm = grp_smem_bytes(P.S, P.Np, 4);
    if (trap) {
        e = cudaFuncSetAttribute(k_unit<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        k_unit<3><<<grid, 128, sm, st>>>(A);
    } else {
        e = cudaFuncSetAttribute(k_unit<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        k_unit<0><<<grid, 128, sm, st>>>(A);
    }
    return cudaGetLastError();
}

// UpdateAction (vpi_mod.f90:2491-2530): one warp per evaluation.  SM: both tables staged in shared memory and read
// through the zero tail exactly as the production sweep kernel (table_mode 2) does; else through L1/L2.
template <bool TRAP, bool SM>
__global__ void __launch_bounds__(128) k_update_action(int n, const double* Rsoa, const int* ip, const int* ib,
                                                      const double* xnew, const double* xold, double* dS) {
    if (SM) {
        extern __shared__ __align__(16) double pigs_smem_base[];
        const int ntab = tab_len(cP.Nmax);
        for (int i = threadIdx.x; i < ntab; i += blockDim.x) { pigs_smem_base[i] = cP.vtab[i]; pigs_smem_base[ntab + i] = cP.logwf[i]; }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int e = w; e < n; e += nw) {
        double xo[3] = {0, 0, 0}, xn[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 3; ++k) if (k < cP.dim) { xo[k] = xold[e * cP.dim + k]; xn[k] = xnew[e * cP.dim + k]; }
        const double* Rx = Rsoa + (size_t)e * 3 * cP.NpS;
        Partner first;
        first.x = first.y = first.z = 0.0;
        if (lane < cP.Np) first = load_partner(Rx, lane);
        // global-memory tables: the reference's roundings (XR); shared-memory tables: exactly the production (Philox) instance
        double t = bead_eval<TRAP, SM, SM, false, !SM>(Rx, ip[e] - 1, ib[e], lane, 32, lane == 0, xo, xn, lane, nullptr, first);
        if (lane == 0) dS[e] = t;
    }
}
cudaError_t launch_update_action(bool trap, bool smem_tables, const DevParams& P, int n, const double* Rsoa, const int* ip, const int* ib,
                                 const double* xnew, const double* xold, double* dS, cudaStream_t st) {
    cudaError_t e = upload(P, 128, st);
    if (e != cudaSuccess) return e;
    int blocks = (n + 3) / 4;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    if (trap) k_update_action<true, false><<<blocks, 128, 0, st>>>(n, Rsoa, ip, ib, xnew, xold, dS);
    else if (smem_tables) {
        const size_t sm = (size_t)2 * tab_len(P.Nmax) * sizeof(double);
        e = cudaFuncSetAttribute(k_update_action<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        if (blocks > 148) blocks = 148;
        k_update_action<false, true><<<blocks, 128, sm, st>>>(n, Rsoa, ip, ib, xnew, xold, dS);
    } else k_update_action<false, false><<<blocks, 128, 0, st>>>(n, Rsoa, ip, ib, xnew, xold, dS);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- layout transposes
// aos: [chain][ib][ip][dim]  <->  blocked soa: [chain][ib][NpS/32][3][32]  (pidx, pigs_device.cuh)
__global__ void k_aos_to_soa(const __grid_constant__ DevParams P, const double* aos, double* soa, long long nslice, int* flag) {
    const long long per = (long long)3 * P.NpS;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nslice * per; i += (long long)gridDim.x * blockDim.x) {
        long long s = i / per;
        int r = (int)(i - s * per), blk = r / PBLK, w = r - blk * PBLK, k = w >> 5, ip = blk * 32 + (w & 31);
        double v = 0.0;
        if (k < P.dim && ip < P.Np) {
            v = aos[(s * P.Np + ip) * P.dim + k];
            // the action kernel takes ONE periodic image per component (as the reference does, pbc_mod.f90:29-52):
            // coordinates must lie in the box [-L/2, L/2] (NaN is caught too)
            if (!P.trap && flag && !(fabs(v) <= P.Lh[k] * (1.0 + 1e-12))) atomicOr(flag, 1);
        }
        soa[i] = v;
    }
}
__global__ void k_soa_to_aos(const __grid_constant__ DevParams P, const double* soa, double* aos, long long nslice) {
    const long long per = (long long)P.Np * P.dim;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nslice * per; i += (long long)gridDim.x * blockDim.x) {
        long long s = i / per;
        int r = (int)(i - s * per), ip = r / P.dim, k = r - ip * P.dim;
        aos[i] = soa[s * 3 * P.NpS + pidx(ip) + 32 * k];
    }
}
cudaError_t launch_aos_to_soa(const DevParams& P, const double* aos, double* soa, int nchain, cudaStream_t st, int* flag) {
    long long nslice = (long long)nchain * P.S;
    k_aos_to_soa<<<148 * 8, 256, 0, st>>>(P, aos, soa, nslice, flag);
    return cudaGetLastError();
}
cudaError_t launch_soa_to_aos(const DevParams& P, const double* soa, double* aos, int nchain, cudaStream_t st) {
    long long nslice = (long long)nchain * P.S;
    k_soa_to_aos<<<148 * 8, 256, 0, st>>>(P, soa, aos, nslice);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- block accumulators
// vec[i] = sum over chains, fixed order (deterministic).  One warp per element:
// lanes stride the chains, then a shuffle tree.
__global__ void k_reduce_block(const __grid_constant__ DevParams P, double* vec) {
    const int nvec = NE + NCNT + (P.nacc - NE);
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int i = w; i < nvec; i += nw) {
        double s = 0.0;
        for (int c = lane; c < P.n_chains; c += 32) {
            if (i < NE) s += P.acc[(size_t)c * P.nacc + i];
            else if (i < NE + NCNT) s += (double)P.cnt[(size_t)c * NCNT + (i - NE)];
            else s += P.acc[(size_t)c * P.nacc + (i - NCNT)];
        }
        s = warp_sum(s);
        if (lane == 0) vec[i] = s;
    }
}
cudaError_t launch_reduce_block(const DevParams& P, double* vec, cudaStream_t st) {
    k_reduce_block<<<148, 256, 0, st>>>(P, vec);
    return cudaGetLastError();
}
__global__ void k_zero_block(const __grid_constant__ DevParams P) {
    const long long n = (long long)P.n_chains * P.nacc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) P.acc[i] = 0.0;
}
cudaError_t launch_zero_block(const DevParams& P, cudaStream_t st) {
    k_zero_block<<<148 * 2, 256, 0, st>>>(P);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- FP64 peak
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double* sink) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;
}
cudaError_t launch_dfma_peak(int blocks, int threads, int iters, double* sink, cudaStream_t st) {
    k_dfma_peak<<<blocks, threads, 0, st>>>(iters, sink);
    return cudaGetLastError();
}

}  // namespace pigs
