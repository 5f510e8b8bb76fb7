// Isolated benchmark of the pair loop (bead_eval) of libpigs_cuda: one warp per "chain", random slices
// from a big buffer (HBM-resident like production) or a small one (L2-resident).  Not part of the product.
#include "../pathintegralgroundstate_b200/csrc/pigs_device.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstring>
using namespace pigs;

#ifndef LB_THREADS
#define LB_THREADS 512
#endif
template <int KINDSEL>
__global__ void __launch_bounds__(LB_THREADS, 1) k_loop(const double* slices, int nslices, int iters, double* out) {
    extern __shared__ __align__(16) double pigs_smem_base[];
    const int ntab = tab_len(cP.Nmax);
    for (int i = threadIdx.x; i < ntab; i += blockDim.x) { pigs_smem_base[i] = cP.vtab[i]; pigs_smem_base[ntab + i] = cP.logwf[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned h = warp * 2654435761u + 12345u;
    double acc = 0.0;
    const size_t ss = (size_t)3 * cP.NpS;
    Partner first; first.x = first.y = first.z = 0.0;
    Carry cy; cy.a = first; cy.b = first; cy.next = nullptr; cy.K = nullptr;
    for (int it = 0; it < iters; ++it) {
        h = h * 1664525u + 1013904223u;
        int s = (h >> 8) % nslices;
#if defined(PF_BULK) || defined(PF_LINE)
        {   // prefetch the slice of the NEXT iteration (one bead ahead)
            unsigned h2 = h * 1664525u + 1013904223u;
            const double* Rn = slices + (size_t)((h2 >> 8) % nslices) * ss;
#ifdef PF_BULK
            if (lane == 0) prefetch_slice_L2(Rn);
#else
            for (int l = lane; l < (int)(ss * 8 / 128); l += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(Rn) + l * 128));
#endif
        }
#endif
        const double* Rx = slices + (size_t)s * ss;
        int ip0 = (h >> 3) % cP.Np;
        int ib = KINDSEL == 0 ? 2 : (KINDSEL == 1 ? 3 : (KINDSEL == 2 ? 0 : 1 + (int)((h >> 20) % 29)));
        double xo[3] = {Rx[pidx(ip0)], Rx[pidx(ip0) + PY], Rx[pidx(ip0) + PZ]};
        double xn[3] = {xo[0] + 0.05, xo[1] - 0.03, xo[2] + 0.02};
#if PIGS_LOOPV & 512
        if (lane < cP.Np) first = load_partner(Rx, lane);
        acc += bead_eval<false, true, true, false>(Rx, ip0, ib, lane, 32, lane == 0, xo, xn, lane, nullptr, first, nullptr, nullptr,
                                                   pigs_smem_base + 2 * ntab + (threadIdx.x >> 5) * 384);
#elif PIGS_LOOPV & 256
        if (it == 0) { cy.a = load_partner(Rx, lane); if (lane + 32 < cP.Np) cy.b = load_partner(Rx, lane + 32); }
        cy.next = slices + (size_t)(((h * 1664525u + 1013904223u) >> 8) % nslices) * ss;      // the slice of the next iteration
        acc += bead_eval<false, true, true, false>(Rx, ip0, ib, lane, 32, lane == 0, xo, xn, lane, nullptr, first, nullptr, &cy);
#else
        if (lane < cP.Np) first = load_partner(Rx, lane);
        acc += bead_eval<false, true, true, false>(Rx, ip0, ib, lane, 32, lane == 0, xo, xn, lane, nullptr, first);
#endif
    }
    if (lane == 0) out[warp] = acc;
}

int main(int argc, char** argv) {
    int Np = argc > 1 ? atoi(argv[1]) : 256;
    int nslices = argc > 2 ? atoi(argv[2]) : 65536;
    int iters = argc > 3 ? atoi(argv[3]) : 2000;
    int warps_per_sm = argc > 4 ? atoi(argv[4]) : 16;
    DevParams P; memset(&P, 0, sizeof P);
    P.dim = 3; P.Np = Np; P.Nb = 15; P.S = 31; P.NpS = (Np + 31) & ~31; P.Nmax = 10000;
    double L = cbrt(Np / 0.365);
    for (int k = 0; k < 3; ++k) { P.L[k] = L; P.Lh[k] = L / 2; P.invL[k] = 1 / L; unsigned long long b; memcpy(&b, &P.Lh[k], 8); unsigned hi = (unsigned)(b >> 32); memcpy(&P.LhF[k], &hi, 4); }
    P.tabW_off = 10006 * 8;
    double rcut = L / 2; P.rcut2 = rcut * rcut; P.dr = rcut / 9999.0; P.inv_dr = 1 / P.dr; P.half_inv_dr = 0.5 * P.inv_dr; P.rclamp2 = (P.Nmax + 3.5) * P.dr * (P.Nmax + 3.5) * P.dr; P.dt = 5e-3; P.wS[0] = 2 * P.dt / 3; P.wS[1] = 4 * P.dt / 3; P.wS[2] = P.dt / 3; P.cF = 4 * P.dt * P.dt * P.dt / 18;
    std::vector<double> tab(10006, 0.0), slices((size_t)nslices * 3 * P.NpS);
    for (int i = 0; i < 10002; ++i) { double r = (i + 1) * P.dr; tab[i] = 1.0 / (r * r * r + 0.1); }
    for (auto& v : slices) v = (rand() / (double)RAND_MAX - 0.5) * L;
    double *d_tab, *d_sl, *d_out;
    cudaMalloc(&d_tab, tab.size() * 8); cudaMalloc(&d_sl, slices.size() * 8); cudaMalloc(&d_out, 1 << 22);
    cudaMemcpy(d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_sl, slices.data(), slices.size() * 8, cudaMemcpyHostToDevice);
    P.vtab = d_tab; P.logwf = d_tab;
    cudaMemcpyToSymbol(cP, &P, sizeof P);
    size_t smem = 2 * 10006 * 8 + 16 * 384 * 8;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](auto kern, const char* name) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int block = warps_per_sm * 32, grid = 148;
        kern<<<grid, block, smem>>>(d_sl, nslices, 10, d_out);
        cudaEventRecord(a); kern<<<grid, block, smem>>>(d_sl, nslices, iters, d_out); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double upd = (double)grid * warps_per_sm * iters;
        std::vector<double> ho((size_t)grid * warps_per_sm);
        cudaMemcpy(ho.data(), d_out, ho.size() * 8, cudaMemcpyDeviceToHost);
        double cs = 0; for (double v : ho) cs += v;
        printf("V%-2d %-6s Np=%d slices=%d (%.0f MB) warps/SM=%d: %.1f M bead-updates/s  checksum %.15e (%s)\n", PIGS_LOOPV, name, Np, nslices, slices.size() * 8 / 1e6, warps_per_sm, upd / ms / 1e3, cs, cudaGetErrorString(cudaGetLastError()));
    };
    run(k_loop<0>, "even"); run(k_loop<1>, "odd"); run(k_loop<2>, "end"); run(k_loop<3>, "mixed");
    return 0;
}
