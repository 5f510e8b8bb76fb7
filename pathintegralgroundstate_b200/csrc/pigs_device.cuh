// pigs_device.cuh -- device-side building blocks of libpigs_cuda (sm_100a).
//
// Execution model.  One *chain group* of T threads (T = 32..512, a whole number
// of warps) owns one Markov chain; a CTA hosts G groups that share the
// shared-memory copies of the interpolation tables and synchronise with named
// barriers (bar.sync id,T) or __syncwarp when T == 32.  Every thread of a group
// follows the same control flow; decisions (segment choice, Metropolis) are
// computed redundantly or broadcast through shared memory, so all branching on
// Monte-Carlo state is group-uniform.  The pair sums of one bead-update
// (reference: UpdateAction, vpi_mod.f90:2491-2530) are spread over the lanes of
// a warp, partner j <-> lane, and reduced with a halving butterfly of warp
// shuffles; beads of one move are spread over the warps of the group.
//
// Data layout in HBM (per chain): path[ib][k][ip] -- structure-of-arrays per
// time slice so that lane<->partner loads are unit-stride and coalesced
// (the ABI layout Path(dim,Np,0:2Nb) is array-of-structures; the library
// transposes on upload/download).  Unused components (dim<3) are zero.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace pigs {

constexpr int NE = 12;      // energy sums
constexpr int NCNT = 24;    // int64 counters, same order as pigs_block_result
enum Cnt {
    C_IDIAG = 0, C_NGR, C_TRY_CM, C_TRY_STAG, C_TRY_CM_HALF, C_TRY_STAG_HALF,
    C_ACC_CM, C_ACC_BD, C_ACC_HEAD, C_ACC_TAIL, C_ACC_CM_HALF, C_ACC_BD_HALF, C_ACC_HEAD_HALF, C_ACC_TAIL_HALF,
    C_TRY_OPEN, C_ACC_OPEN, C_TRY_CLOSE, C_ACC_CLOSE, C_TRY_SWAP, C_ACC_SWAP,
    C_UPD_EVEN, C_UPD_ODD, C_UPD_END, C_NOPEN
};
// per-chain integer state
enum IState { IS_OPEN = 0, IS_IWORM, IS_IPERM, IS_NEWPC, IS_ENDPC, IS_MTI, IS_IK, IS_IDIAG_AUX, IS_N };

struct DevParams {
    int dim, Np, Nb, S, NpS, Nmax, Nbin, Nk, Npw;
    int trap, sampling, Lstag, Nlev, Nstag, Nobdm, swapping, CMFreq;
    int n_chains;
    double L[3], Lh[3], qbin[3], a_ho[3];
    double rcut2, dr, inv_dr, rbin, dt, delta_cm, CWorm, density, pi, logCd;
    unsigned long long seed;
    const double* logwf;      // (0:Nmax+1) in global memory
    const double* vtab;
    double* path;             // [chain][ib][k][NpS]
    double* xend;             // [chain][2][3]
    int* istate;              // [chain][IS_N]
    int* cyc;                 // [chain][Np]  Particles_in_perm_cycle
    int* hist;                // [chain][Np]  Perm_histogram
    unsigned* mt;             // [chain][624]
    unsigned long long* pctr; // [chain] Philox counter
    double* acc;              // [chain][nacc]
    long long* cnt;           // [chain][NCNT]
    size_t chain_stride;      // doubles per chain in path
    int nacc;                 // NE + Nbin + dim*Nk + Nbin*(Npw+1)
    int off_gr, off_sk, off_nr;
};

// ------------------------------------------------------------------ group
struct Grp {
    int tid, size, warp, lane, nwarps, bar;
    __device__ __forceinline__ void sync() const {
        if (size == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(size) : "memory");
    }
};

// per-group shared-memory scratch (carved from dynamic smem)
struct GrpSmem {
    double* seg_old;   // [3][S]   beads of the moved particle before the move
    double* seg_new;   // [3][S]   ... proposed
    double* part;      // [max(nwarps,4)*8] partial sums / reduction scratch
    double* bc;        // [8] broadcast slots
    double* pp;        // [Np] swap probabilities
    int*    ibc;       // [8] int broadcast
};
// layout: seg_old[3S] seg_new[3S] part[np*8] bc[8] pp[Np] ibc[8 ints] eacc[NE] cnt[NCNT]
__host__ __device__ inline size_t grp_smem_doubles(int S, int Np, int nwarps) {
    int np = nwarps < 4 ? 4 : nwarps;
    return (size_t)6 * S + (size_t)np * 8 + 8 + (size_t)Np + 4 /*8 ints*/ + NE + NCNT;
}

// ------------------------------------------------------------------ tables
// TABMODE 0: both tables through the read-only L1/L2 path; 1: VTable in shared
// memory, LogWF global; 2: both in shared memory.
struct Tabs {
    const double* V;
    const double* W;
};
template <bool SM>
__device__ __forceinline__ double tld(const double* t, int i) {
    if (SM) return t[i];
    return __ldg(t + i);
}

// Interpolate(opt,...) of interpolate.f90:1-45 with x/dx -> x*inv_dx (every
// interpolant is continuous in x, so an index flip at a grid point is harmless).
struct Lk {
    int i0;
    double a1, a2;
};
__device__ __forceinline__ Lk lk_prep(double r, double dr, double inv_dr, int Nmax) {
    Lk k;
    double t = r * inv_dr;
    int i0 = (int)t;                    // = ix-1
    i0 = max(1, min(i0, Nmax - 1));     // keep i0-1 .. i0+2 inside (0:Nmax+1); never active for r in (dr, rcut)
    k.i0 = i0;
    k.a1 = fma(-(double)i0, dr, r);     // aux1 = x-(ix-1)*dx
    k.a2 = dr - k.a1;
    return k;
}
template <bool SM>
__device__ __forceinline__ double lk_val(const double* F, const Lk& k, double inv_dr) {   // opt 0
    return (k.a1 * tld<SM>(F, k.i0 + 1) + k.a2 * tld<SM>(F, k.i0)) * inv_dr;
}
template <bool SM>
__device__ __forceinline__ void lk_val_d1(const double* F, const Lk& k, double inv_dr, double& v, double& d1) {   // opt 0 and 1
    double fm = tld<SM>(F, k.i0 - 1), f0 = tld<SM>(F, k.i0), f1 = tld<SM>(F, k.i0 + 1), f2 = tld<SM>(F, k.i0 + 2);
    double Fc = k.a1 * f1 + k.a2 * f0;
    double Fb = k.a1 * f0 + k.a2 * fm;
    double Fa = k.a1 * f2 + k.a2 * f1;
    v = Fc * inv_dr;
    d1 = 0.5 * (Fa - Fb) * inv_dr * inv_dr;
}
template <bool SM>
__device__ __forceinline__ void lk_d1_d2(const double* F, const Lk& k, double inv_dr, double& d1, double& d2) {   // opt 1 and 2
    double fm = tld<SM>(F, k.i0 - 1), f0 = tld<SM>(F, k.i0), f1 = tld<SM>(F, k.i0 + 1), f2 = tld<SM>(F, k.i0 + 2);
    double Fc = k.a1 * f1 + k.a2 * f0;
    double Fb = k.a1 * f0 + k.a2 * fm;
    double Fa = k.a1 * f2 + k.a2 * f1;
    d1 = 0.5 * (Fa - Fb) * inv_dr * inv_dr;
    d2 = (Fa - 2.0 * Fc + Fb) * inv_dr * inv_dr * inv_dr;
}

// Interpolate opt 1 and 2 in the reference's exact operation order, true
// divisions, no FMA contraction.  The second difference divides rounding noise
// by dx^2 (~1e7): only the same arithmetic reproduces the reference's value to
// 1e-10, so the mixed estimator (2 calls per MC step, not the hot loop) pays
// for IEEE divisions here.
template <bool SM>
__device__ __forceinline__ void lk_exact_d1_d2(const double* F, double x, double dx, int Nmax, double& d1, double& d2) {
    int ix = (int)__ddiv_rn(x, dx) + 1;
    ix = max(2, min(ix, Nmax));
    double aux1 = __dadd_rn(x, -__dmul_rn((double)(ix - 1), dx));
    double aux2 = __dadd_rn(dx, -aux1);
    double fm = tld<SM>(F, ix - 2), f0 = tld<SM>(F, ix - 1), f1 = tld<SM>(F, ix), f2 = tld<SM>(F, ix + 1);
    double Fb = __ddiv_rn(__dadd_rn(__dmul_rn(aux1, f0), __dmul_rn(aux2, fm)), dx);
    double Fc = __ddiv_rn(__dadd_rn(__dmul_rn(aux1, f1), __dmul_rn(aux2, f0)), dx);
    double Fa = __ddiv_rn(__dadd_rn(__dmul_rn(aux1, f2), __dmul_rn(aux2, f1)), dx);
    d1 = __ddiv_rn(__dmul_rn(0.5, __dadd_rn(Fa, -Fb)), dx);
    d2 = __ddiv_rn(__dadd_rn(__dadd_rn(Fa, -__dmul_rn(2.0, Fc)), Fb), __dmul_rn(dx, dx));
}
template <bool SM>
__device__ __forceinline__ double lk_exact_val(const double* F, double x, double dx, int Nmax) {
    int ix = (int)__ddiv_rn(x, dx) + 1;
    ix = max(2, min(ix, Nmax));
    double aux1 = __dadd_rn(x, -__dmul_rn((double)(ix - 1), dx));
    double aux2 = __dadd_rn(dx, -aux1);
    return __ddiv_rn(__dadd_rn(__dmul_rn(aux1, tld<SM>(F, ix)), __dmul_rn(aux2, tld<SM>(F, ix - 1))), dx);
}

// ------------------------------------------------------------------ geometry
// MinimumImage / BoundaryConditions (pbc_mod.f90:11-52): one shift, '>' first
__device__ __forceinline__ double mimg(double d, double L, double Lh) {
    if (d > Lh) d -= L;
    if (d < -Lh) d += L;
    return d;
}
// the in-line wrap of the bridge code (e.g. vpi_mod.f90:519-520): '<' first
__device__ __forceinline__ double mimg_lt_first(double d, double L, double Lh) {
    if (d < -Lh) d += L;
    if (d > Lh) d -= L;
    return d;
}

// ------------------------------------------------------------------ warp reductions
__device__ __forceinline__ double shx(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double warp_sum(double v) {
    v += shx(v, 16); v += shx(v, 8); v += shx(v, 4); v += shx(v, 2); v += shx(v, 1);
    return v;
}
// Halving butterfly over 8 per-lane values: afterwards every lane of quad q
// (lanes 4q..4q+3) holds the full 32-lane sum of a[q].  9 64-bit shuffles
// instead of 40.
__device__ __forceinline__ double warp_sum8(const double (&a)[8], int lane) {
    double b[4], c[2], d;
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double send = h16 ? a[i] : a[i + 4];
        double keep = h16 ? a[i + 4] : a[i];
        b[i] = keep + shx(send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        double send = h8 ? b[i] : b[i + 2];
        double keep = h8 ? b[i + 2] : b[i];
        c[i] = keep + shx(send, 8);
    }
    {
        double send = h4 ? c[0] : c[1];
        double keep = h4 ? c[1] : c[0];
        d = keep + shx(send, 4);
    }
    d += shx(d, 2);
    d += shx(d, 1);
    return d;      // value index = lane>>2
}

// ------------------------------------------------------------------ RNG
// MT19937 with the 1998 seeding, bit-faithful to random_mod.f90:5-115.  State
// lives in global memory (624 words per chain); only thread 0 of a group draws.
__device__ inline void mt_seed(unsigned* mt, int& mti, unsigned seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 69069u * mt[i - 1];
    mti = 624;
}
__device__ inline unsigned mt_next(unsigned* mt, int& mti) {
    if (mti >= 624) {
        if (mti == 625) mt_seed(mt, mti, 4357u);
        for (int kk = 0; kk < 624; ++kk) {
            unsigned y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
            mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        mti = 0;
    }
    unsigned y = mt[mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}
__device__ inline double mt_grnd(unsigned* mt, int& mti) {        // [0,1] inclusive
    return (double)mt_next(mt, mti) / 4294967295.0;
}
__device__ inline double mt_rangauss(unsigned* mt, int& mti) {    // random_mod.f90:195-219, first deviate
    double u1, u2, w;
    do {
        u1 = 2.0 * mt_grnd(mt, mti) - 1.0;
        u2 = 2.0 * mt_grnd(mt, mti) - 1.0;
        w = u1 * u1 + u2 * u2;
    } while (!(w <= 1.0));
    w = sqrt((-2.0 * log(w)) / w);
    return u1 * w;
}

// Philox4x32-10 (Salmon et al., SC'11)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

template <bool MT>
struct Rng {
    // MT replay
    unsigned* mt;
    int mti;
    // Philox
    unsigned long long ctr;
    uint2 key;
    unsigned chain;
    int slot;      // broadcast slot toggle (group-uniform)

    __device__ __forceinline__ uint4 philox_at(unsigned long long c) const {
        return philox4x32_10(make_uint4((unsigned)c, (unsigned)(c >> 32), chain, 0x50494753u), key);
    }
};
__device__ __forceinline__ double u01_from(unsigned lo, unsigned hi) {        // [0,1)
    unsigned long long x = ((unsigned long long)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

// one uniform, identical in every thread of the group
template <bool MT>
__device__ __forceinline__ double rng_uniform(const Grp& G, Rng<MT>& rng, const GrpSmem& sm) {
    if (MT) {
        rng.slot ^= 1;
        if (G.tid == 0) sm.bc[rng.slot] = mt_grnd(rng.mt, rng.mti);
        G.sync();
        return sm.bc[rng.slot];
    } else {
        uint4 r = rng.philox_at(rng.ctr);
        rng.ctr += 1;
        return u01_from(r.x, r.y);
    }
}
// unit Gaussians for beads b0 + j*bstride (j<nb), components k<dim, written into
// seg_new[k*S + bead] in the reference's draw order (bead-major, component-minor).
// Ends with a group sync.
template <bool MT>
__device__ __forceinline__ void rng_gauss_fill(const Grp& G, Rng<MT>& rng, const GrpSmem& sm, int S, int dim,
                                               int b0, int bstride, int nb) {
    const int n = nb * dim;
    if (MT) {
        if (G.tid == 0) {
            for (int i = 0; i < n; ++i) {
                int j = i / dim, k = i - j * dim;
                sm.seg_new[k * S + b0 + j * bstride] = mt_rangauss(rng.mt, rng.mti);
            }
        }
    } else {
        for (int i = G.tid; i < n; i += G.size) {
            int j = i / dim, k = i - j * dim;
            uint4 r = rng.philox_at(rng.ctr + (unsigned long long)i);
            double u1 = 1.0 - u01_from(r.x, r.y);       // (0,1]
            double u2 = u01_from(r.z, r.w);
            sm.seg_new[k * S + b0 + j * bstride] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        }
        rng.ctr += (unsigned long long)n;
    }
    G.sync();
}

// ------------------------------------------------------------------ the pair sums of one bead-update
// Values accumulated per lane for one displaced bead (UpdatePot, vpi_mod.f90:2660-2841,
// UpdateWf, :2534-2656), old and new position against partner j:
//   a[0] = PotNew-PotOld   a[1] = PsiNew-PsiOld
//   a[2..4] = Fnew(k)      a[5..7] = Fold(k)
// KIND: 0 interior even slice, 1 odd slice (Chin force term), 2 end slice (Jastrow).
template <int KIND, bool TRAP, bool VSM, bool WSM>
__device__ __forceinline__ void pair_one(const DevParams& P, const Tabs& T, double x0, double x1, double x2,
                                         double rx, double ry, double rz, bool is_new, double sgn, double (&a)[8]) {
    double d0 = x0 - rx, d1 = x1 - ry, d2 = x2 - rz;
    if (!TRAP) {
        d0 = mimg(d0, P.L[0], P.Lh[0]);
        d1 = mimg(d1, P.L[1], P.Lh[1]);
        d2 = mimg(d2, P.L[2], P.Lh[2]);
    }
    double r2 = d0 * d0 + d1 * d1 + d2 * d2;
    // PBC: both positions cut at rcut (Q24).  Trap: UpdatePot cuts only the OLD
    // position (Q12), UpdateWf cuts nothing.
    bool in_pot = TRAP ? (is_new || r2 <= P.rcut2) : (r2 <= P.rcut2);
    bool in_wf = TRAP ? true : in_pot;
    if (in_pot || (KIND == 2 && in_wf)) {
        double r = sqrt(r2);
        Lk k = lk_prep(r, P.dr, P.inv_dr, P.Nmax);
        if (in_pot) {
            if (KIND == 1) {
                double v, dv;
                lk_val_d1<VSM>(T.V, k, P.inv_dr, v, dv);
                a[0] += sgn * v;
                double s = dv / r;
                int o = is_new ? 2 : 5;
                a[o] += s * d0; a[o + 1] += s * d1; a[o + 2] += s * d2;
            } else {
                a[0] += sgn * lk_val<VSM>(T.V, k, P.inv_dr);
            }
        }
        if (KIND == 2) a[1] += sgn * lk_val<WSM>(T.W, k, P.inv_dr);
    }
}

template <int KIND, bool TRAP, bool VSM, bool WSM>
__device__ __forceinline__ void pair_loop(const DevParams& P, const Tabs& T, const double* Rx, int ip0, int j0, int jstride,
                                          const double (&xo)[3], const double (&xn)[3], double (&a)[8]) {
    const double* Ry = Rx + P.NpS;
    const double* Rz = Ry + P.NpS;
    for (int j = j0; j < P.Np; j += jstride) {
        if (j == ip0) continue;
        double rx = Rx[j], ry = Ry[j], rz = Rz[j];
        pair_one<KIND, TRAP, VSM, WSM>(P, T, xn[0], xn[1], xn[2], rx, ry, rz, true, 1.0, a);
        pair_one<KIND, TRAP, VSM, WSM>(P, T, xo[0], xo[1], xo[2], rx, ry, rz, false, -1.0, a);
    }
}

__device__ __forceinline__ int bead_kind(int ib, int Nb) { return (ib == 0 || ib == 2 * Nb) ? 2 : (ib & 1); }

// lane-partial sums of one bead against partners j0, j0+jstride, ...; the lane
// with add_self adds the one-body (trap) terms once.
template <bool TRAP, bool VSM, bool WSM>
__device__ __forceinline__ void bead_partial(const DevParams& P, const Tabs& T, const double* Rx, int ip0, int ib,
                                             int j0, int jstride, bool add_self, const double (&xo)[3],
                                             const double (&xn)[3], double (&a)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.0;
    const int kind = bead_kind(ib, P.Nb);
    if (TRAP && add_self) {
        for (int k = 0; k < P.dim; ++k) {          // system_mod.f90:213-252
            double ak = P.a_ho[k], a2 = ak * ak, a4 = a2 * a2;
            a[0] += 0.5 * xn[k] * xn[k] / a4 - 0.5 * xo[k] * xo[k] / a4;
            if (kind == 1) { a[2 + k] += xn[k] / a4; a[5 + k] += xo[k] / a4; }
            if (kind == 2) a[1] += -0.5 * (xn[k] / ak) * (xn[k] / ak) + 0.5 * (xo[k] / ak) * (xo[k] / ak);
        }
    }
    if (kind == 0) pair_loop<0, TRAP, VSM, WSM>(P, T, Rx, ip0, j0, jstride, xo, xn, a);
    else if (kind == 1) pair_loop<1, TRAP, VSM, WSM>(P, T, Rx, ip0, j0, jstride, xo, xn, a);
    else pair_loop<2, TRAP, VSM, WSM>(P, T, Rx, ip0, j0, jstride, xo, xn, a);
}

// DeltaS of UpdateAction from the eight reduced values
// (GreenFunction opt 0, global_mod.f90:29-46).
__device__ __forceinline__ double assemble_dS(const DevParams& P, int ib, const double (&v)[8]) {
    const int kind = bead_kind(ib, P.Nb);
    double dt = P.dt;
    if (kind == 2) return -v[1] + dt * v[0] / 3.0;
    if (kind == 0) return 2.0 * dt * v[0] / 3.0;
    double f2 = (v[2] * v[2] + v[3] * v[3] + v[4] * v[4]) - (v[5] * v[5] + v[6] * v[6] + v[7] * v[7]);
    return 4.0 * dt * (v[0] + dt * dt * f2 / 6.0) / 3.0;
}
// the same, distributed: lane holds value index q = lane>>2 (after warp_sum8);
// returns the lane's additive term of DeltaS.
__device__ __forceinline__ double dS_term(const DevParams& P, int ib, int q, double val) {
    const int kind = bead_kind(ib, P.Nb);
    double dt = P.dt;
    if (q == 0) return (kind == 2 ? dt / 3.0 : (kind == 0 ? 2.0 * dt / 3.0 : 4.0 * dt / 3.0)) * val;
    if (q == 1) return kind == 2 ? -val : 0.0;
    if (kind != 1) return 0.0;
    double c = 4.0 * dt * dt * dt / 18.0;
    return (q < 5 ? c : -c) * val * val;
}

}  // namespace pigs
