/*
 * pigs_oracle.cpp -- CPU ORACLE: serial C++ restatement of the Fortran 90
 * reference amaciarey/PathIntegralGroundState.   TEST INFRASTRUCTURE ONLY
 * (see pigs_oracle.h: who may call it, and why parity is "unpinned").
 *
 * Every function cites the reference file:line it follows.  The statement
 * order, the order of random draws, the float32 casts and the quirks listed in
 * SURVEY.md Appendix B are reproduced on purpose; nothing is "improved".
 * Compile with -O2 -ffp-contract=off (x86-64 gfortran -O2 emits no FMA).
 */
#include "pigs_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

/* ---------------------------------------------------------------- random_mod.f90 */
struct Mt19937_1998 {
    uint32_t mt[624];
    int mti = 625;                       /* data mti/N1/ : random_mod.f90:59 */

    void sgrnd(uint32_t seed) {          /* random_mod.f90:5-31 */
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 69069u * mt[i - 1];
        mti = 624;                       /* do-loop variable ends at N */
    }
    uint32_t raw() {                     /* random_mod.f90:35-106 */
        const uint32_t MATA = 0x9908b0dfu, UMASK = 0x80000000u, LMASK = 0x7fffffffu;
        if (mti >= 624) {
            if (mti == 625) sgrnd(4357u);
            int kk;
            for (kk = 0; kk < 624 - 397; ++kk) {
                uint32_t y = (mt[kk] & UMASK) | (mt[kk + 1] & LMASK);
                mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? MATA : 0u);
            }
            for (; kk < 623; ++kk) {
                uint32_t y = (mt[kk] & UMASK) | (mt[kk + 1] & LMASK);
                mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MATA : 0u);
            }
            uint32_t y = (mt[623] & UMASK) | (mt[0] & LMASK);
            mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MATA : 0u);
            mti = 0;
        }
        uint32_t y = mt[mti++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double grnd() {                      /* random_mod.f90:108-112 : [0,1] inclusive */
        return (double)raw() / (4294967296.0 - 1.0);
    }
    double rangauss() {                  /* random_mod.f90:195-219, sigma=1 mu=0, x1 only */
        double u1, u2, w;
        for (;;) {
            u1 = 2.0 * grnd() - 1.0;
            u2 = 2.0 * grnd() - 1.0;
            w = u1 * u1 + u2 * u2;
            if (w <= 1.0) break;
        }
        w = std::sqrt((-2.0 * std::log(w)) / w);
        return 0.0 + 1.0 * u1 * w;
    }
};

inline double ipow(double x, int n) {    /* gfortran x**n for small integer n: repeated products */
    double r = 1.0;
    for (int i = 0; i < n; ++i) r *= x;
    return r;
}

/* ---------------------------------------------------------------- interpolate.f90:1-45 */
double Interpolate(int opt, int N, double dx, const double* F, double x) {
    (void)N;
    int ix = (int)(x / dx) + 1;
    double aux1 = x - (ix - 1) * dx;
    double aux2 = dx - aux1;
    double r = 0.0;
    if (opt == 0) {
        r = (aux1 * F[ix] + aux2 * F[ix - 1]) / dx;
    } else if (opt == 1) {
        double Fbefore = (aux1 * F[ix - 1] + aux2 * F[ix - 2]) / dx;
        double Fafter  = (aux1 * F[ix + 1] + aux2 * F[ix]) / dx;
        r = 0.5 * (Fafter - Fbefore) / dx;
    } else if (opt == 2) {
        double Fbefore = (aux1 * F[ix - 1] + aux2 * F[ix - 2]) / dx;
        double Fcurr   = (aux1 * F[ix] + aux2 * F[ix - 1]) / dx;
        double Fafter  = (aux1 * F[ix + 1] + aux2 * F[ix]) / dx;
        r = (Fafter - 2.0 * Fcurr + Fbefore) / (dx * dx);
    }
    return r;
}

/* ---------------------------------------------------------------- system_mod.f90 */
double LogPsi(int opt, double Rm, double rij) {           /* system_mod.f90:38-66 */
    /* (Rm/rij)**5 as gfortran/GCC expand a constant integer power (the powi table): x^5 = (x * x^2) * x^2 --
     * pinned by the machine translation of system_mod.f90 (oracle/_ref, tests/test_ref_pin.py) */
    const double q = Rm / rij, q2 = q * q, q3 = q * q2;
    double q5 = q3 * q2;
    if (opt == 0) return -0.5 * q5;
    if (opt == 1) return 2.5 * q5 / rij;
    return -15.0 * q5 / (rij * rij);
}

double Potential(double rij) {                            /* system_mod.f90:136-182, Aziz II HFD-B(HE) */
    const double E_0 = 10.948, rm = 2.963, A = 1.8443101e5, alpha = 10.43329537,
                 beta = -2.27965105, C6 = 1.36745214, C8 = 0.42123807, C10 = 0.17473318,
                 D = 1.4826;
    const double V0 = E_0 / 1.85505153154686;
    double dij = rij * 2.556 / rm;
    double dij2 = dij * dij, dij4 = dij2 * dij2, dij6 = dij4 * dij2;
    double Hx;
    if (dij <= D) {
        double t = D / dij - 1.0;
        Hx = std::exp(-(t * t));
    } else {
        Hx = 1.0;
    }
    return V0 * (A * std::exp(-alpha * dij + beta * dij2) - (C6 + C8 / dij2 + C10 / dij4) * Hx / dij6);
}

double TrapPsi(int opt, double a, double x) {             /* system_mod.f90:213-234 */
    if (opt == 0) return -0.5 * ((x / a) * (x / a));
    if (opt == 1) return -(x / (a * a));
    return -1.0 / (a * a);
}
double TrapPot(int opt, double a, double x) {             /* system_mod.f90:238-252 */
    double a4 = (a * a) * (a * a);
    if (opt == 0) return 0.5 * (x * x) / a4;
    return x / a4;
}

} // namespace

/* ================================================================ the simulation object
 * = global_mod + system_mod module variables + the driver's locals */
struct orc_sim {
    /* global_mod.f90:5-12 */
    bool wf_table, v_table;
    double pi, rbin, dr, rcut, rcut2, CWorm;
    int dim, Np, Nbin, Nb, Nmax, Npw;
    double Lbox[3], LboxHalf[3], qbin[3];
    /* system_mod.f90:8-9 */
    double Rm, a_ho[3];
    /* driver scalars (vpi.f90) */
    bool trap, crystal, swapping;
    int sampling, seed, CMFreq, Lstag, Nlev, Nstag, Nobdm, Nk, action = 0;
    double density, dt, delta_cm;
    std::vector<double> Path, LogWF, VTable;
    double xend[2][3];                       /* xend(k,j) -> xend[j-1][k-1] */
    bool isopen = false;
    int iworm = 0;
    bool new_perm_cycle = false, end_perm_cycle = false, swap_accepted = false;
    int iperm = 0, ik = 0;
    std::vector<int32_t> Particles_in_perm_cycle, Perm_histogram;
    int idiag = 0, idiag_aux = 0;
    Mt19937_1998 rng;
    uint64_t nupd[3] = {0, 0, 0};

    inline double& P(int k, int ip, int ib) { return Path[(size_t)(k - 1) + (size_t)dim * ((size_t)(ip - 1) + (size_t)Np * (size_t)ib)]; }
    inline double* slice(int ib) { return &Path[(size_t)dim * (size_t)Np * (size_t)ib]; }

    /* pbc_mod.f90:11-25 */
    inline void BoundaryConditions(int k, double& x) const {
        if (x > LboxHalf[k - 1]) x = x - Lbox[k - 1];
        if (x < -LboxHalf[k - 1]) x = x + Lbox[k - 1];
    }
    /* pbc_mod.f90:29-52 */
    inline void MinimumImage(double* xij, double& rij2) const {
        rij2 = 0.0;
        for (int k = 0; k < dim; ++k) {
            if (xij[k] > LboxHalf[k]) xij[k] = xij[k] - Lbox[k];
            if (xij[k] < -LboxHalf[k]) xij[k] = xij[k] + Lbox[k];
            rij2 = rij2 + xij[k] * xij[k];
        }
    }
    /* global_mod.f90:19-72 */
    double GreenFunction(int opt, int ib, double dt_, double Pot, double F2) const {
        double g = 0.0;
        if (action == 1) return opt == 0 ? dt_ * Pot : Pot;      /* global_mod.f90:48,67 (commented alternative) */
        if (opt == 0) {
            double Ve = Pot, Vc = Pot + dt_ * dt_ * F2 / 6.0;
            if (ib == 0) g = dt_ * Ve / 3.0;
            else if (ib == 2 * Nb) g = dt_ * Ve / 3.0;
            else if (ib % 2 == 0) g = 2.0 * dt_ * Ve / 3.0;
            else g = 4.0 * dt_ * Vc / 3.0;
        } else if (opt == 1) {
            double dVe = Pot, dVc = Pot + dt_ * dt_ * F2 / 2.0;
            if (ib == 0) g = dVe / 3.0;
            else if (ib == 2 * Nb) g = dVe / 3.0;
            else if (ib % 2 == 0) g = 2.0 * dVe / 3.0;
            else g = 4.0 * dVc / 3.0;
        }
        return g;
    }

    /* vpi_mod.f90:2534-2656 */
    void UpdateWf(int ip, const double* R, const double* xnew, const double* xold, double& DeltaPsi) const {
        double PsiOld = 0.0, PsiNew = 0.0;
        double xijold[3], xijnew[3], rijold2, rijnew2, rijold, rijnew, urold, urnew;
        if (trap) {
            for (int k = 0; k < dim; ++k) {
                PsiOld = PsiOld + TrapPsi(0, a_ho[k], xold[k]);
                PsiNew = PsiNew + TrapPsi(0, a_ho[k], xnew[k]);
            }
        }
        for (int jp = 1; jp <= Np; ++jp) {
            if (jp != ip) {
                rijold2 = 0.0; rijnew2 = 0.0;
                for (int k = 0; k < dim; ++k) {
                    xijold[k] = xold[k] - R[k + dim * (jp - 1)];
                    xijnew[k] = xnew[k] - R[k + dim * (jp - 1)];
                }
                if (trap) {
                    for (int k = 0; k < dim; ++k) {
                        rijold2 = rijold2 + xijold[k] * xijold[k];
                        rijnew2 = rijnew2 + xijnew[k] * xijnew[k];
                    }
                } else {
                    MinimumImage(xijnew, rijnew2);
                    MinimumImage(xijold, rijold2);
                }
                if (trap) {
                    rijold = std::sqrt(rijold2);
                    urold = wf_table ? Interpolate(0, Nmax, dr, LogWF.data(), rijold) : LogPsi(0, Rm, rijold);
                    PsiOld = PsiOld + urold;
                    rijnew = std::sqrt(rijnew2);
                    urnew = wf_table ? Interpolate(0, Nmax, dr, LogWF.data(), rijnew) : LogPsi(0, Rm, rijnew);
                    PsiNew = PsiNew + urnew;
                } else {
                    if (rijold2 <= rcut2) {
                        rijold = std::sqrt(rijold2);
                        urold = wf_table ? Interpolate(0, Nmax, dr, LogWF.data(), rijold) : LogPsi(0, Rm, rijold);
                        PsiOld = PsiOld + urold;
                    }
                    if (rijnew2 <= rcut2) {
                        rijnew = std::sqrt(rijnew2);
                        urnew = wf_table ? Interpolate(0, Nmax, dr, LogWF.data(), rijnew) : LogPsi(0, Rm, rijnew);
                        PsiNew = PsiNew + urnew;
                    }
                }
            }
        }
        DeltaPsi = PsiNew - PsiOld;
    }

    /* system_mod.f90:186-209: Force() has an EMPTY body in the reference (returns
     * garbage); v_table=F is therefore unusable on odd beads.  The oracle defines
     * it as 0 so that v_table=F stays deterministic; no test relies on it. */
    static double Force(int, const double*, double) { return 0.0; }

    /* vpi_mod.f90:2660-2841 */
    void UpdatePot(int ip, const double* R, const double* xnew, const double* xold,
                   double& DeltaPot, double* DeltaF2) const {
        double PotNew = 0.0, PotOld = 0.0;
        double Fnew[3] = {0, 0, 0}, Fold[3] = {0, 0, 0};
        double xijnew[3], xijold[3], rijnew2, rijold2, rijnew, rijold;
        const double* VT = VTable.data();
        if (trap) {
            for (int k = 0; k < dim; ++k) {
                PotNew = PotNew + TrapPot(0, a_ho[k], xnew[k]);
                PotOld = PotOld + TrapPot(0, a_ho[k], xold[k]);
                Fold[k] = TrapPot(1, a_ho[k], xold[k]);
                Fnew[k] = TrapPot(1, a_ho[k], xnew[k]);
            }
        }
        for (int jp = 1; jp <= Np; ++jp) {
            if (jp != ip) {
                rijnew2 = 0.0; rijold2 = 0.0;
                for (int k = 0; k < dim; ++k) {
                    xijnew[k] = xnew[k] - R[k + dim * (jp - 1)];
                    xijold[k] = xold[k] - R[k + dim * (jp - 1)];
                }
                if (trap) {
                    for (int k = 0; k < dim; ++k) {
                        rijold2 = rijold2 + xijold[k] * xijold[k];
                        rijnew2 = rijnew2 + xijnew[k] * xijnew[k];
                    }
                } else {
                    MinimumImage(xijold, rijold2);
                    MinimumImage(xijnew, rijnew2);
                }
                if (trap) {
                    rijnew = std::sqrt(rijnew2);                       /* no cutoff on NEW in trap mode (Q12) */
                    if (v_table) PotNew = PotNew + Interpolate(0, Nmax, dr, VT, rijnew);
                    else PotNew = PotNew + Potential(rijnew);
                    if (DeltaF2) {
                        if (v_table) for (int k = 0; k < dim; ++k) Fnew[k] = Fnew[k] + Interpolate(1, Nmax, dr, VT, rijnew) * xijnew[k] / rijnew;
                        else for (int k = 0; k < dim; ++k) Fnew[k] = Fnew[k] + Force(k, xijnew, rijnew);
                    }
                    if (rijold2 <= rcut2) {
                        rijold = std::sqrt(rijold2);
                        if (v_table) PotOld = PotOld + Interpolate(0, Nmax, dr, VT, rijold);
                        else PotOld = PotOld + Potential(rijold);
                        if (DeltaF2) {
                            if (v_table) for (int k = 0; k < dim; ++k) Fold[k] = Fold[k] + Interpolate(1, Nmax, dr, VT, rijold) * xijold[k] / rijold;
                            else for (int k = 0; k < dim; ++k) Fold[k] = Fold[k] + Force(k, xijold, rijold);
                        }
                    }
                } else {
                    if (rijnew2 <= rcut2) {
                        rijnew = std::sqrt(rijnew2);
                        if (v_table) PotNew = PotNew + Interpolate(0, Nmax, dr, VT, rijnew);
                        else PotNew = PotNew + Potential(rijnew);
                        if (DeltaF2) {
                            if (v_table) for (int k = 0; k < dim; ++k) Fnew[k] = Fnew[k] + Interpolate(1, Nmax, dr, VT, rijnew) * xijnew[k] / rijnew;
                            else for (int k = 0; k < dim; ++k) Fnew[k] = Fnew[k] + Force(k, xijnew, rijnew);
                        }
                    }
                    if (rijold2 <= rcut2) {
                        rijold = std::sqrt(rijold2);
                        if (v_table) PotOld = PotOld + Interpolate(0, Nmax, dr, VT, rijold);
                        else PotOld = PotOld + Potential(rijold);
                        if (DeltaF2) {
                            if (v_table) for (int k = 0; k < dim; ++k) Fold[k] = Fold[k] + Interpolate(1, Nmax, dr, VT, rijold) * xijold[k] / rijold;
                            else for (int k = 0; k < dim; ++k) Fold[k] = Fold[k] + Force(k, xijold, rijold);
                        }
                    }
                }
            }
        }
        double Fnew2 = 0.0, Fold2 = 0.0;
        if (DeltaF2) {
            for (int k = 0; k < dim; ++k) {
                Fnew2 = Fnew2 + Fnew[k] * Fnew[k];
                Fold2 = Fold2 + Fold[k] * Fold[k];
            }
            *DeltaF2 = Fnew2 - Fold2;
        }
        DeltaPot = PotNew - PotOld;
    }

    /* vpi_mod.f90:2491-2530 ; R = Path(:,:,ib) */
    void UpdateActionR(const double* R, int ip, int ib, const double* xnew, const double* xold, double dt_, double& DeltaS) {
        double DeltaPot, DeltaF2, DeltaLogPsi;
        if (ib % 2 == 0) {
            UpdatePot(ip, R, xnew, xold, DeltaPot, nullptr);
            DeltaF2 = 0.0;
        } else {
            UpdatePot(ip, R, xnew, xold, DeltaPot, &DeltaF2);
        }
        if (ib == 0) UpdateWf(ip, R, xnew, xold, DeltaLogPsi);
        else if (ib == 2 * Nb) UpdateWf(ip, R, xnew, xold, DeltaLogPsi);
        else DeltaLogPsi = 0.0;
        DeltaS = -DeltaLogPsi + GreenFunction(0, ib, dt_, DeltaPot, DeltaF2);
        /* the metric unit: one bead-update */
        if (ib == 0 || ib == 2 * Nb) ++nupd[2];
        else if (ib % 2 == 0) ++nupd[0];
        else ++nupd[1];
    }
    void UpdateAction(int ip, int ib, const double* xnew, const double* xold, double dt_, double& DeltaS) {
        UpdateActionR(slice(ib), ip, ib, xnew, xold, dt_, DeltaS);
    }

    /* the Metropolis question, 19 identical sites, e.g. vpi_mod.f90:356-364 */
    bool Metropolis(double S) {
        bool accept;
        if (std::exp(-S) >= 1.0) accept = true;
        else {
            if (std::exp(-S) >= rng.grnd()) accept = true;
            else accept = false;
        }
        return accept;
    }

    /* int(n*grnd()) as written at every site of Appendix A.  grnd() is [0,1] INCLUSIVE
     * (Q6): u == 1 (p = 2^-32 per draw) indexes one past the range in the reference,
     * which is an out-of-bounds access there; the oracle (and the CUDA path, identically)
     * clamps that single case so long baseline runs cannot corrupt memory. */
    inline int draw_int(int n) {
        int v = (int)(n * rng.grnd());
        if (v >= n) v = n > 0 ? n - 1 : 0;
        return v;
    }

    /* ---- bridge primitives: identical text instantiated ~20 times in vpi_mod.f90 ---- */
    /* xprev(k) = Path(k,ip,iprev)-xold(k); wrap; xprev = xold+xprev   (e.g. vpi_mod.f90:517-522) */
    inline double unwrap_prev(int k, double xold, double pprev) const {
        double d = pprev - xold;
        if (!trap) {
            if (d < -LboxHalf[k - 1]) d = d + Lbox[k - 1];
            if (d > LboxHalf[k - 1]) d = d - Lbox[k - 1];
        }
        return xold + d;
    }
    /* xnext(k) = xold(k)-Path(k,ip,inext); wrap; xnext = xold-xnext   (e.g. vpi_mod.f90:524-529) */
    inline double unwrap_next(int k, double xold, double pnext) const {
        double d = xold - pnext;
        if (!trap) {
            if (d < -LboxHalf[k - 1]) d = d + Lbox[k - 1];
            if (d > LboxHalf[k - 1]) d = d - Lbox[k - 1];
        }
        return xold - d;
    }
    /* one staging bead: vpi_mod.f90:509-549 (and 651-691, 791-831, 1416-1456, 1581-1621,
     * 1742-1782, 1985-2025, 2160-2200, 2396-2436) */
    double stage_bead(int ip, int ii, int L, int j, int ie) {
        double xnew[3], xold[3], DeltaS;
        for (int k = 1; k <= dim; ++k) {
            xold[k - 1] = P(k, ip, ii + j);
            double gauss1 = rng.rangauss();
            double xprev = unwrap_prev(k, xold[k - 1], P(k, ip, ii + j - 1));
            double xnext = unwrap_next(k, xold[k - 1], P(k, ip, ie));
            double sigma = std::sqrt((double)((float)(L - j) / (float)(L - j + 1)) * dt);     /* Q15: float32 ratio */
            double xmid = (xnext + xprev * (double)(L - j)) / (double)(float)(L - j + 1);
            xnew[k - 1] = xmid + sigma * gauss1;
            if (!trap) BoundaryConditions(k, xnew[k - 1]);
            P(k, ip, ii + j) = xnew[k - 1];
        }
        UpdateAction(ip, ii + j, xnew, xold, dt, DeltaS);
        return DeltaS;
    }
    /* free end anchored on the NEXT side (head-like): vpi_mod.f90:619-645, 1039-1066, 1544-1571, 1950-1977 */
    double free_end_next(int ip, int iend, int ianchor, double sigma) {
        double xnew[3], xold[3], DeltaS;
        for (int k = 1; k <= dim; ++k) {
            xold[k - 1] = P(k, ip, iend);
            double gauss1 = rng.rangauss();
            double xnext = unwrap_next(k, xold[k - 1], P(k, ip, ianchor));
            double xmid = xnext;
            xnew[k - 1] = xmid + sigma * gauss1;
            if (!trap) BoundaryConditions(k, xnew[k - 1]);
            P(k, ip, iend) = xnew[k - 1];
        }
        UpdateAction(ip, iend, xnew, xold, dt, DeltaS);
        return DeltaS;
    }
    /* free end anchored on the PREV side (tail-like): vpi_mod.f90:758-785, 1227-1254, 1705-1732, 1886-1913 */
    double free_end_prev(int ip, int iend, int ianchor, double sigma) {
        double xnew[3], xold[3], DeltaS;
        for (int k = 1; k <= dim; ++k) {
            xold[k - 1] = P(k, ip, iend);
            double gauss1 = rng.rangauss();
            double xprev = unwrap_prev(k, xold[k - 1], P(k, ip, ianchor));
            double xmid = xprev;
            xnew[k - 1] = xmid + sigma * gauss1;
            if (!trap) BoundaryConditions(k, xnew[k - 1]);
            P(k, ip, iend) = xnew[k - 1];
        }
        UpdateAction(ip, iend, xnew, xold, dt, DeltaS);
        return DeltaS;
    }
    /* multilevel part: vpi_mod.f90:903-971 (and 1083-1151, 1271-1339) */
    bool bisect_levels(int ip, int ii, int Nl) {
        bool accept = false;
        for (int ilev = 1; ilev <= Nl; ++ilev) {
            int delta_ib = 1 << (Nl - ilev + 1);
            double dt_bis = 0.5 * (double)(float)delta_ib * dt;
            double sigma = std::sqrt(0.5 * dt_bis);
            double LevelDeltaS = 0.0;
            for (int j = 1; j <= (1 << (ilev - 1)); ++j) {
                int iprev = ii + (j - 1) * delta_ib;
                int inext = ii + j * delta_ib;
                int icurr = (iprev + inext) / 2;
                double xnew[3], xold[3], DeltaS;
                for (int k = 1; k <= dim; ++k) {
                    xold[k - 1] = P(k, ip, icurr);
                    double gauss1 = rng.rangauss();
                    double xprev = unwrap_prev(k, xold[k - 1], P(k, ip, iprev));
                    double xnext = unwrap_next(k, xold[k - 1], P(k, ip, inext));
                    double xmid = 0.5 * (xprev + xnext);
                    xnew[k - 1] = xmid + sigma * gauss1;
                    if (!trap) BoundaryConditions(k, xnew[k - 1]);
                    P(k, ip, icurr) = xnew[k - 1];
                }
                UpdateAction(ip, icurr, xnew, xold, dt, DeltaS);
                LevelDeltaS = LevelDeltaS + DeltaS;
            }
            if (std::exp(-LevelDeltaS) >= 1.0) accept = true;
            else {
                if (std::exp(-LevelDeltaS) >= rng.grnd()) accept = true;
                else { accept = false; break; }
            }
        }
        return accept;
    }
    void save_chain(std::vector<double>& Old, int ip, int ii, int ie) {
        Old.resize((size_t)dim * (2 * Nb + 1));
        for (int ib = ii; ib <= ie; ++ib) for (int k = 1; k <= dim; ++k) Old[(k - 1) + dim * ib] = P(k, ip, ib);
    }
    void restore_chain(const std::vector<double>& Old, int ip, int ii, int ie) {
        for (int ib = ii; ib <= ie; ++ib) for (int k = 1; k <= dim; ++k) P(k, ip, ib) = Old[(k - 1) + dim * ib];
    }

    /* ---- the 14 moves ---- */
    /* vpi_mod.f90:313-379 */
    void TranslateChain(double delta, int ip, int& accepted) {
        double dx[3], xold[3], xnew[3], DeltaS, SumDeltaS;
        std::vector<double> NewChain((size_t)dim * (2 * Nb + 1));
        for (int k = 0; k < dim; ++k) dx[k] = delta * (2.0 * rng.grnd() - 1.0);
        SumDeltaS = 0.0;
        for (int ib = 0; ib <= 2 * Nb; ++ib) {
            for (int k = 1; k <= dim; ++k) {
                xold[k - 1] = P(k, ip, ib);
                xnew[k - 1] = xold[k - 1] + dx[k - 1];
                if (!trap) BoundaryConditions(k, xnew[k - 1]);
                NewChain[(k - 1) + dim * ib] = xnew[k - 1];
            }
            UpdateAction(ip, ib, xnew, xold, dt, DeltaS);
            SumDeltaS = SumDeltaS + DeltaS;
        }
        if (Metropolis(SumDeltaS)) {
            accepted = accepted + 1;
            for (int ib = 0; ib <= 2 * Nb; ++ib) for (int k = 1; k <= dim; ++k) P(k, ip, ib) = NewChain[(k - 1) + dim * ib];
        }
    }
    /* vpi_mod.f90:383-476 */
    void TranslateHalfChain(int half, double delta, int ip, int& accepted) {
        double dx[3], xold[3], xnew[3], DeltaS, SumDeltaS;
        std::vector<double> OldChain;
        for (int k = 1; k <= dim; ++k) P(k, ip, Nb) = xend[half - 1][k - 1];
        for (int k = 0; k < dim; ++k) dx[k] = delta * (2.0 * rng.grnd() - 1.0);
        SumDeltaS = 0.0;
        int ibi, ibf;
        if (half == 1) { ibi = 0; ibf = Nb; } else { ibi = Nb; ibf = 2 * Nb; }
        save_chain(OldChain, ip, ibi, ibf);
        for (int ib = ibi; ib <= ibf; ++ib) {
            for (int k = 1; k <= dim; ++k) {
                xold[k - 1] = P(k, ip, ib);
                xnew[k - 1] = xold[k - 1] + dx[k - 1];
                if (!trap) BoundaryConditions(k, xnew[k - 1]);
                P(k, ip, ib) = xnew[k - 1];
            }
            UpdateAction(ip, ib, xnew, xold, dt, DeltaS);
            SumDeltaS = SumDeltaS + DeltaS;
        }
        if (Metropolis(SumDeltaS)) {
            accepted = accepted + 1;
            for (int k = 1; k <= dim; ++k) xend[half - 1][k - 1] = P(k, ip, Nb);
        } else {
            restore_chain(OldChain, ip, ibi, ibf);
        }
    }
    /* vpi_mod.f90:480-578 */
    void Staging(int Ls, int ip, int& accepted) {
        std::vector<double> OldChain;
        int ii = draw_int(2 * Nb - Ls + 1);
        int ie = ii + Ls;
        save_chain(OldChain, ip, ii, ie);
        double SumDeltaS = 0.0;
        for (int j = 1; j <= Ls - 1; ++j) SumDeltaS = SumDeltaS + stage_bead(ip, ii, Ls, j, ie);
        if (Metropolis(SumDeltaS)) accepted = accepted + 1;
        else restore_chain(OldChain, ip, ii, ie);
    }
    /* vpi_mod.f90:582-720 */
    void MoveHead(int Lmax, int ip, int& accepted) {
        std::vector<double> OldChain;
        int Ls = draw_int(Lmax - 1) + 2;
        int ii = 0, ie = ii + Ls;
        save_chain(OldChain, ip, ii, ie);
        double SumDeltaS = 0.0;
        SumDeltaS = SumDeltaS + free_end_next(ip, ii, ie, std::sqrt((double)(float)Ls * dt));
        for (int j = 1; j <= Ls - 1; ++j) SumDeltaS = SumDeltaS + stage_bead(ip, ii, Ls, j, ie);
        if (Metropolis(SumDeltaS)) accepted = accepted + 1;
        else restore_chain(OldChain, ip, ii, ie);
    }
    /* vpi_mod.f90:724-860 */
    void MoveTail(int Lmax, int ip, int& accepted) {
        std::vector<double> OldChain;
        int Ls = draw_int(Lmax - 1) + 2;
        int ii = 2 * Nb - Ls, ie = 2 * Nb;
        save_chain(OldChain, ip, ii, ie);
        double SumDeltaS = 0.0;
        SumDeltaS = SumDeltaS + free_end_prev(ip, ie, ii, std::sqrt((double)(float)Ls * dt));
        for (int j = 1; j <= Ls - 1; ++j) SumDeltaS = SumDeltaS + stage_bead(ip, ii, Ls, j, ie);
        if (Metropolis(SumDeltaS)) accepted = accepted + 1;
        else restore_chain(OldChain, ip, ii, ie);
    }
    /* vpi_mod.f90:864-998 */
    void Bisection(int level, int ip, int& accepted) {
        std::vector<double> OldChain;
        int Nl = level;
        int ii = draw_int(2 * Nb - (1 << Nl) + 1);
        int ie = ii + (1 << Nl);
        save_chain(OldChain, ip, ii, ie);
        bool accept = bisect_levels(ip, ii, Nl);
        if (accept) accepted = accepted + 1;
        else restore_chain(OldChain, ip, ii, ie);
    }
    /* vpi_mod.f90:1002-1184 */
    void MoveHeadBisection(int level, int ip, int& accepted) {
        std::vector<double> OldChain;
        int Nl = draw_int(level - 1) + 2;
        int ii = 0, ie = ii + (1 << Nl);
        save_chain(OldChain, ip, ii, ie);
        double DeltaS = free_end_next(ip, ii, ie, std::sqrt((double)(1 << Nl) * dt));
        bool continue_bisecting = Metropolis(DeltaS);
        bool accept;
        if (continue_bisecting) accept = bisect_levels(ip, ii, Nl);
        else accept = false;
        if (accept) accepted = accepted + 1;
        else restore_chain(OldChain, ip, ii, ie);
    }
    /* vpi_mod.f90:1188-1372 */
    void MoveTailBisection(int level, int ip, int& accepted) {
        std::vector<double> OldChain;
        int Nl = draw_int(level - 1) + 2;
        int ii = 2 * Nb - (1 << Nl), ie = 2 * Nb;
        save_chain(OldChain, ip, ii, ie);
        double DeltaS = free_end_prev(ip, ie, ii, std::sqrt((double)(1 << Nl) * dt));
        bool continue_bisecting = Metropolis(DeltaS);
        bool accept;
        if (continue_bisecting) accept = bisect_levels(ip, ii, Nl);
        else accept = false;
        if (accept) accepted = accepted + 1;
        else restore_chain(OldChain, ip, ii, ie);
    }
    /* vpi_mod.f90:1376-1491 */
    void StagingHalfChain(int half, int Ls, int ip, int& accepted) {
        std::vector<double> OldChain;
        for (int k = 1; k <= dim; ++k) P(k, ip, Nb) = xend[half - 1][k - 1];
        int ii, ie;
        if (half == 1) { ii = draw_int(Nb - Ls + 1); ie = ii + Ls; }
        else { ii = draw_int(Nb - Ls + 1) + Nb; ie = ii + Ls; }
        save_chain(OldChain, ip, ii, ie);
        double SumDeltaS = 0.0;
        for (int j = 1; j <= Ls - 1; ++j) SumDeltaS = SumDeltaS + stage_bead(ip, ii, Ls, j, ie);
        if (Metropolis(SumDeltaS)) {
            accepted = accepted + 1;
            for (int k = 1; k <= dim; ++k) xend[half - 1][k - 1] = P(k, ip, Nb);
        } else {
            restore_chain(OldChain, ip, ii, ie);
        }
    }
    /* vpi_mod.f90:1495-1656 */
    void MoveHeadHalfChain(int half, int Lmax, int ip, int& accepted) {
        std::vector<double> OldChain;
        int Ls = draw_int(Lmax - 1) + 2;
        for (int k = 1; k <= dim; ++k) P(k, ip, Nb) = xend[half - 1][k - 1];
        int ii, ie;
        if (half == 1) { ii = 0; ie = ii + Ls; } else { ii = Nb; ie = ii + Ls; }
        save_chain(OldChain, ip, ii, ie);
        double SumDeltaS = 0.0;
        double DeltaS = free_end_next(ip, ii, ie, std::sqrt((double)(float)Ls * dt));
        if (half == 1) SumDeltaS = SumDeltaS + DeltaS;
        else SumDeltaS = SumDeltaS + 0.5 * DeltaS;
        for (int j = 1; j <= Ls - 1; ++j) SumDeltaS = SumDeltaS + stage_bead(ip, ii, Ls, j, ie);
        if (Metropolis(SumDeltaS)) {
            accepted = accepted + 1;
            for (int k = 1; k <= dim; ++k) xend[half - 1][k - 1] = P(k, ip, Nb);
        } else {
            restore_chain(OldChain, ip, ii, ie);
        }
    }
    /* vpi_mod.f90:1660-1817 */
    void MoveTailHalfChain(int half, int Lmax, int ip, int& accepted) {
        std::vector<double> OldChain;
        int Ls = draw_int(Lmax - 1) + 2;
        for (int k = 1; k <= dim; ++k) P(k, ip, Nb) = xend[half - 1][k - 1];
        int ii, ie;
        if (half == 1) { ii = Nb - Ls; ie = Nb; } else { ii = 2 * Nb - Ls; ie = 2 * Nb; }
        save_chain(OldChain, ip, ii, ie);
        double SumDeltaS = 0.0;
        double DeltaS = free_end_prev(ip, ii + Ls, ii, std::sqrt((double)(float)Ls * dt));
        if (half == 1) SumDeltaS = SumDeltaS + 0.5 * DeltaS;
        else SumDeltaS = SumDeltaS + DeltaS;
        for (int j = 1; j <= Ls - 1; ++j) SumDeltaS = SumDeltaS + stage_bead(ip, ii, Ls, j, ie);
        if (Metropolis(SumDeltaS)) {
            accepted = accepted + 1;
            for (int k = 1; k <= dim; ++k) xend[half - 1][k - 1] = P(k, ip, Nb);
        } else {
            restore_chain(OldChain, ip, ii, ie);
        }
    }
    double link_DeltaK(int ip, int ii, int ie, int Ls) {      /* vpi_mod.f90:1859-1873 etc. */
        double xij[3], rij2;
        for (int k = 1; k <= dim; ++k) xij[k - 1] = P(k, ip, ii) - P(k, ip, ie);
        if (trap) { rij2 = 0.0; for (int k = 0; k < dim; ++k) rij2 = rij2 + xij[k] * xij[k]; }
        else MinimumImage(xij, rij2);
        return -0.5 * rij2 / ((double)(float)Ls * dt) - 0.5 * (double)(float)dim * std::log(2.0 * pi * (double)(float)Ls * dt);
    }
    /* vpi_mod.f90:1821-2076 */
    void OpenChain(int Lmax, int ip, int& accepted) {
        std::vector<double> OldChain;
        int Ls = 2 * draw_int((Lmax - 2) / 2) + 2;
        int half = (int)(rng.grnd() * 2) + 1;
        if (half > 2) half = 2;
        double SumDeltaS = -std::log(CWorm * density);
        double DeltaS, DeltaK;
        int ii, ie;
        if (half == 1) {
            ii = Nb - Ls; ie = Nb;
            DeltaK = link_DeltaK(ip, ii, ie, Ls);
            save_chain(OldChain, ip, ii, ie);
            DeltaS = free_end_prev(ip, ie, ii, std::sqrt((double)(float)Ls * dt));
        } else {
            ii = Nb; ie = Nb + Ls;
            DeltaK = link_DeltaK(ip, ii, ie, Ls);
            save_chain(OldChain, ip, ii, ie);
            DeltaS = free_end_next(ip, ii, ie, std::sqrt((double)(float)Ls * dt));
        }
        SumDeltaS = SumDeltaS + 0.5 * DeltaS;
        for (int j = 1; j <= Ls - 1; ++j) SumDeltaS = SumDeltaS + stage_bead(ip, ii, Ls, j, ie);
        bool accept;
        if (std::exp(-SumDeltaS - DeltaK) >= 1.0) accept = true;
        else {
            if (std::exp(-SumDeltaS - DeltaK) >= rng.grnd()) accept = true;
            else accept = false;
        }
        if (accept) {
            isopen = true;
            accepted = accepted + 1;
            if (half == 1) {
                for (int k = 1; k <= dim; ++k) { xend[0][k - 1] = P(k, ip, Nb); xend[1][k - 1] = OldChain[(k - 1) + dim * Nb]; }
            } else {
                for (int k = 1; k <= dim; ++k) { xend[0][k - 1] = OldChain[(k - 1) + dim * Nb]; xend[1][k - 1] = P(k, ip, Nb); }
            }
            new_perm_cycle = true;
        } else {
            restore_chain(OldChain, ip, ii, ie);
            for (int k = 1; k <= dim; ++k) { xend[0][k - 1] = P(k, ip, Nb); xend[1][k - 1] = xend[0][k - 1]; }
            new_perm_cycle = false;
        }
    }
    /* vpi_mod.f90:2080-2266 */
    void CloseChain(int Lmax, int ip, int& accepted) {
        std::vector<double> OldChain;
        int Ls = 2 * draw_int((Lmax - 2) / 2) + 2;
        int half = (int)(rng.grnd() * 2) + 1;
        if (half > 2) half = 2;
        double SumDeltaS = std::log(CWorm * density);
        double DeltaS, xnew[3], xold[3];
        int ii, ie;
        if (half == 1) {
            ii = Nb - Ls; ie = Nb;
            save_chain(OldChain, ip, ii, ie);
            for (int k = 1; k <= dim; ++k) {
                P(k, ip, ie) = xend[1][k - 1];
                xold[k - 1] = OldChain[(k - 1) + dim * ie];
                xnew[k - 1] = P(k, ip, ie);
            }
            UpdateAction(ip, ie, xnew, xold, dt, DeltaS);
        } else {
            ii = Nb; ie = Nb + Ls;
            save_chain(OldChain, ip, ii, ie);
            for (int k = 1; k <= dim; ++k) {
                P(k, ip, ii) = xend[0][k - 1];
                xold[k - 1] = OldChain[(k - 1) + dim * ii];
                xnew[k - 1] = P(k, ip, ii);
            }
            UpdateAction(ip, ii, xnew, xold, dt, DeltaS);
        }
        SumDeltaS = SumDeltaS + 0.5 * DeltaS;
        for (int j = 1; j <= Ls - 1; ++j) SumDeltaS = SumDeltaS + stage_bead(ip, ii, Ls, j, ie);
        double DeltaK = link_DeltaK(ip, ii, ie, Ls);
        bool accept;
        if (std::exp(-SumDeltaS + DeltaK) >= 1.0) accept = true;
        else {
            if (std::exp(-SumDeltaS + DeltaK) >= rng.grnd()) accept = true;
            else accept = false;
        }
        if (accept) {
            isopen = false;
            accepted = accepted + 1;
            for (int k = 1; k <= dim; ++k) { xend[0][k - 1] = P(k, ip, Nb); xend[1][k - 1] = xend[0][k - 1]; }
            end_perm_cycle = true;
        } else {
            restore_chain(OldChain, ip, ii, ie);
            end_perm_cycle = false;
        }
    }
    /* vpi_mod.f90:2270-2487 */
    void Swap(int Lmax, int iw, int& accepted, int& ipar, bool& swap_acc) {
        std::vector<double> Pp((size_t)Np, 0.0), OldChain((size_t)dim * (2 * Nb + 1)), OldWorm((size_t)dim * (2 * Nb + 1));
        double xij[3], rij2, Sk, Sw, uran, sum;
        swap_acc = false;
        int Ls = 2 * draw_int((Lmax - 2) / 2) + 2;
        int ii = Nb - Ls, ie = Nb;
        int ikl = 0;
        Sw = 0.0;
        for (int ip = 1; ip <= Np; ++ip) {
            for (int k = 1; k <= dim; ++k) xij[k - 1] = P(k, ip, ii) - xend[1][k - 1];
            if (trap) { rij2 = 0.0; for (int k = 0; k < dim; ++k) rij2 = rij2 + xij[k] * xij[k]; }
            else MinimumImage(xij, rij2);
            Pp[ip - 1] = std::exp(-0.5 * rij2 / ((double)(float)Ls * dt));
            Sw = Sw + Pp[ip - 1];
        }
        uran = rng.grnd();
        int ip = 0;
        sum = 0.0;
        for (;;) {
            ip = ip + 1;
            if (ip > Np) { ikl = Np; break; }            /* Q18: the reference would run out of bounds; clamp */
            sum = sum + Pp[ip - 1] / Sw;
            if (uran <= sum) { ikl = ip; break; }
        }
        if (ikl != iw) {
            Sk = 0.0;
            for (int jp = 1; jp <= Np; ++jp) {
                for (int k = 1; k <= dim; ++k) xij[k - 1] = P(k, jp, ii) - P(k, ikl, ie);
                if (trap) { rij2 = 0.0; for (int k = 0; k < dim; ++k) rij2 = rij2 + xij[k] * xij[k]; }
                else MinimumImage(xij, rij2);
                Sk = Sk + std::exp(-0.5 * rij2 / ((double)(float)Ls * dt));
            }
            if (rng.grnd() <= Sw / Sk) {
                for (int ib = 0; ib <= 2 * Nb; ++ib) for (int k = 1; k <= dim; ++k) {
                    OldChain[(k - 1) + dim * ib] = P(k, ikl, ib);
                    OldWorm[(k - 1) + dim * ib] = P(k, iw, ib);
                }
                for (int k = 1; k <= dim; ++k) P(k, ikl, ie) = xend[1][k - 1];
                double SumDeltaS = 0.0;
                for (int j = 1; j <= Ls - 1; ++j) SumDeltaS = SumDeltaS + stage_bead(ikl, ii, Ls, j, ie);
                if (Metropolis(SumDeltaS)) {
                    accepted = accepted + 1;
                    for (int ib = Nb; ib <= 2 * Nb; ++ib) for (int k = 1; k <= dim; ++k) {
                        P(k, iw, ib) = P(k, ikl, ib);
                        P(k, ikl, ib) = OldWorm[(k - 1) + dim * ib];
                    }
                    for (int k = 1; k <= dim; ++k) {
                        xend[1][k - 1] = OldChain[(k - 1) + dim * Nb];
                        P(k, iw, Nb) = xend[1][k - 1];
                    }
                    swap_acc = true;
                    ipar = ikl;
                } else {
                    for (int ib = 0; ib <= 2 * Nb; ++ib) for (int k = 1; k <= dim; ++k) {
                        P(k, ikl, ib) = OldChain[(k - 1) + dim * ib];
                        P(k, iw, ib) = OldWorm[(k - 1) + dim * ib];
                    }
                    swap_acc = false;
                }
            }
        }
    }

    /* ---- estimators: sample_mod.f90 ---- */
    /* sample_mod.f90:13-150 */
    void PotentialEnergy(const double* R, double& Pot, double* F2) const {
        std::vector<double> F((size_t)dim * Np);
        double xij[3], rij2, rij, fij;
        const double* VT = VTable.data();
        Pot = 0.0;
        for (int ip = 0; ip < Np; ++ip) for (int k = 0; k < dim; ++k) {
            if (trap) { F[k + dim * ip] = TrapPot(1, a_ho[k], R[k + dim * ip]); Pot = Pot + TrapPot(0, a_ho[k], R[k + dim * ip]); }
            else F[k + dim * ip] = 0.0;
        }
        for (int ip = 0; ip < Np - 1; ++ip) for (int jp = ip + 1; jp < Np; ++jp) {
            for (int k = 0; k < dim; ++k) xij[k] = R[k + dim * ip] - R[k + dim * jp];
            if (trap) { rij2 = 0.0; for (int k = 0; k < dim; ++k) rij2 = rij2 + xij[k] * xij[k]; }
            else MinimumImage(xij, rij2);
            if (trap || rij2 <= rcut2) {
                rij = std::sqrt(rij2);
                if (v_table) Pot = Pot + Interpolate(0, Nmax, dr, VT, rij);
                else Pot = Pot + Potential(rij);
                if (F2) {
                    for (int k = 0; k < dim; ++k) {
                        if (v_table) fij = Interpolate(1, Nmax, dr, VT, rij) * xij[k] / rij;
                        else fij = Force(k, xij, rij);
                        F[k + dim * ip] = F[k + dim * ip] + fij;
                        F[k + dim * jp] = F[k + dim * jp] - fij;
                    }
                }
            }
        }
        if (F2) {
            *F2 = 0.0;
            for (int ip = 0; ip < Np; ++ip) for (int k = 0; k < dim; ++k) *F2 = *F2 + F[k + dim * ip] * F[k + dim * ip];
        }
    }
    /* sample_mod.f90:154-319 */
    void LocalEnergy(const double* R, double& E, double& Kin, double& Pot) const {
        std::vector<double> F((size_t)dim * Np);
        double xij[3], rij2, rij, fij, dudr, d2udr2, LapLogPsi;
        Kin = 0.0; Pot = 0.0; LapLogPsi = 0.0;
        for (int i = 0; i < Np; ++i) for (int k = 0; k < dim; ++k) {
            if (trap) {
                F[k + dim * i] = TrapPsi(1, a_ho[k], R[k + dim * i]);
                Pot = Pot + TrapPot(0, a_ho[k], R[k + dim * i]);
                LapLogPsi = LapLogPsi + TrapPsi(2, a_ho[k], R[k + dim * i]);
            } else F[k + dim * i] = 0.0;
        }
        LapLogPsi = 0.5 * LapLogPsi;
        for (int i = 0; i < Np - 1; ++i) for (int j = i + 1; j < Np; ++j) {
            for (int k = 0; k < dim; ++k) xij[k] = R[k + dim * i] - R[k + dim * j];
            if (trap) { rij2 = 0.0; for (int k = 0; k < dim; ++k) rij2 = rij2 + xij[k] * xij[k]; }
            else MinimumImage(xij, rij2);
            if (trap || rij2 <= rcut2) {
                rij = std::sqrt(rij2);
                if (wf_table) {
                    dudr = Interpolate(1, Nmax, dr, LogWF.data(), rij);
                    d2udr2 = Interpolate(2, Nmax, dr, LogWF.data(), rij);
                } else {
                    dudr = LogPsi(1, Rm, rij);
                    d2udr2 = LogPsi(2, Rm, rij);
                }
                LapLogPsi = LapLogPsi + (((double)(float)dim - 1) * dudr / rij + d2udr2);
                for (int k = 0; k < dim; ++k) {
                    fij = dudr * xij[k] / rij;
                    F[k + dim * i] = F[k + dim * i] + fij;
                    F[k + dim * j] = F[k + dim * j] - fij;
                }
                if (v_table) Pot = Pot + Interpolate(0, Nmax, dr, VTable.data(), rij);
                else Pot = Pot + Potential(rij);
            }
        }
        Kin = 2.0 * LapLogPsi;
        for (int i = 0; i < Np; ++i) for (int k = 0; k < dim; ++k) Kin = Kin + F[k + dim * i] * F[k + dim * i];
        Kin = -0.5 * Kin;
        E = Kin + Pot;
    }
    /* sample_mod.f90:323-388 */
    void ThermEnergy(const double* Pth, double dt_, double& E, double& Ec, double& Ep) const {
        double Pot, F2, xij[3], rij2;
        Ep = 0.0; Ec = 0.0; E = 0.0;
        const size_t ss = (size_t)dim * Np;
        for (int ib = 0; ib <= 2 * Nb - 1; ++ib) {
            if (ib % 2 == 0) { PotentialEnergy(Pth + ss * ib, Pot, nullptr); F2 = 0.0; }
            else PotentialEnergy(Pth + ss * ib, Pot, &F2);
            if (ib == Nb) Ep = Pot;
            E = E + GreenFunction(1, ib, dt_, Pot, F2);
            for (int ip = 0; ip < Np; ++ip) {
                for (int k = 0; k < dim; ++k) xij[k] = Pth[ss * ib + k + dim * ip] - Pth[ss * (ib + 1) + k + dim * ip];
                if (trap) {
                    rij2 = 0.0;
                    for (int k = 0; k < dim; ++k) rij2 = rij2 + xij[k] * xij[k];
                    E = E - 0.5 * rij2 / (dt_ * dt_);
                } else {
                    MinimumImage(xij, rij2);
                    if (rij2 <= rcut2) E = E - 0.5 * rij2 / (dt_ * dt_);
                }
            }
        }
        E = 0.5 * (E / (double)(float)Nb + (double)(float)(dim * Np) / dt_);
        Ec = E - Ep;
    }
    /* sample_mod.f90:392-431 */
    void PairCorrelation(const double* R, double* gr) const {
        double xij[3], rij2, rij;
        for (int ip = 0; ip < Np - 1; ++ip) for (int jp = ip + 1; jp < Np; ++jp) {
            for (int k = 0; k < dim; ++k) xij[k] = R[k + dim * ip] - R[k + dim * jp];
            MinimumImage(xij, rij2);
            if (rij2 <= rcut2) {
                rij = std::sqrt(rij2);
                int ibin = (int)(rij / rbin) + 1;
                if (ibin <= Nbin) gr[ibin - 1] = gr[ibin - 1] + 2.0;   /* reference: bounds-check abort if rij==rcut exactly */
            }
        }
    }
    /* sample_mod.f90:435-476 */
    void StructureFactor(int Nk_, const double* R, double* Sk) const {
        for (int iq = 1; iq <= Nk_; ++iq) for (int k = 0; k < dim; ++k) {
            double SumCos = 0.0, SumSin = 0.0;
            for (int ip = 0; ip < Np; ++ip) {
                double x = R[k + dim * ip];
                double qr = (double)(float)iq * qbin[k] * x;
                SumCos = SumCos + std::cos(qr);
                SumSin = SumSin + std::sin(qr);
            }
            Sk[k + dim * (iq - 1)] = Sk[k + dim * (iq - 1)] + (SumCos * SumCos + SumSin * SumSin);
        }
    }
    /* sample_mod.f90:480-526 ; xend(dim,2) column-major */
    void OBDM(const double* xe, double* nrho) const {
        double xij[3] = {0, 0, 0}, rij2, rij;
        for (int k = 0; k < dim; ++k) xij[k] = xe[k] - xe[k + dim];
        MinimumImage(xij, rij2);
        if (rij2 <= rcut2) {
            rij = std::sqrt(rij2);
            int ibin = (int)(rij / rbin) + 1;
            if (ibin > Nbin) return;
            double sintheta = (dim >= 2 ? xij[1] : 0.0) / rij;
            double costheta = xij[0] / rij;
            double e2r = costheta * costheta - sintheta * sintheta;      /* exptheta*exptheta */
            double e2i = costheta * sintheta + sintheta * costheta;
            double mr = 1.0, mi = 0.0;
            for (int m = 0; m <= Npw; ++m) {
                nrho[m + (Npw + 1) * (ibin - 1)] = nrho[m + (Npw + 1) * (ibin - 1)] + mr;
                double nr = mr * e2r - mi * e2i, ni = mr * e2i + mi * e2r;
                mr = nr; mi = ni;
            }
        }
    }
    /* sample_mod.f90:530-594 */
    void PermutationSampling(int iw, bool have_swap, int ikk, bool swap_acc) {
        if (!swapping) return;     /* reference touches unallocated arrays here (Q22); configs run swapping=T */
        bool particle_already_in_cycle = false;
        if (new_perm_cycle) {
            std::fill(Particles_in_perm_cycle.begin(), Particles_in_perm_cycle.end(), 0);
            Particles_in_perm_cycle[0] = iw;
            iperm = 1;
            new_perm_cycle = false;
        }
        if (have_swap) {
            if (swap_acc) {
                for (int ip = 0; ip < Np; ++ip) {
                    if (Particles_in_perm_cycle[ip] == ikk) { particle_already_in_cycle = true; break; }
                    else particle_already_in_cycle = false;
                }
                if (end_perm_cycle == false) {
                    if (!particle_already_in_cycle && iperm < Np) {      /* (iperm < Np always holds for a consistent state) */
                        iperm = iperm < 0 ? 1 : iperm + 1;
                        Particles_in_perm_cycle[iperm - 1] = ikk;
                    }
                }
            }
        }
        if (end_perm_cycle) {
            /* an open worm set from outside without its permutation record has iperm = 0: the reference would
               index Perm_histogram(0) (out of bounds); both this oracle and the CUDA path count it as length 1 */
            int len = iperm < 1 ? 1 : (iperm > Np ? Np : iperm);
            Perm_histogram[len - 1] = Perm_histogram[len - 1] + 1;
            if (isopen) {
                std::fill(Particles_in_perm_cycle.begin(), Particles_in_perm_cycle.end(), 0);
                Particles_in_perm_cycle[0] = iw;
                iperm = 1;
            }
            end_perm_cycle = false;
        }
    }

    /* one block of the step loop: vpi.f90:250-475 */
    void RunBlock(int Nstep, orc_block& b, double* gr, double* Sk, double* nrho) {
        std::memset(&b, 0, sizeof b);
        uint64_t n0[3] = {nupd[0], nupd[1], nupd[2]};
        for (int i = 0; i < Nbin; ++i) gr[i] = 0.0;
        for (int i = 0; i < dim * Nk; ++i) Sk[i] = 0.0;
        int acc_cm = 0, acc_bd = 0, acc_head = 0, acc_tail = 0, acc_cm_half = 0, acc_bd_half = 0, acc_head_half = 0,
            acc_tail_half = 0, acc_open = 0, acc_close = 0, acc_swap = 0;
        const size_t ss = (size_t)dim * Np;
        for (int istep = 1; istep <= Nstep; ++istep) {
            int iupdate = (int)(rng.grnd() * 2);
            if (isopen) {
                if (iupdate == 0) {
                    CloseChain(Lstag, iworm, acc_close);
                    b.try_close = b.try_close + 1;
                    PermutationSampling(iworm, false, 0, false);
                }
            } else {
                if (iupdate == 1) {
                    iworm = draw_int(Np) + 1;           /* int(grnd()*Np)+1, vpi.f90:315 */
                    OpenChain(Lstag, iworm, acc_open);
                    b.try_open = b.try_open + 1;
                    PermutationSampling(iworm, false, 0, false);
                }
            }
            if (isopen) {
                if (istep % CMFreq == 0) {
                    for (int ip = 1; ip <= Np; ++ip) if (ip != iworm) { b.try_cm = b.try_cm + 1; TranslateChain(delta_cm, ip, acc_cm); }
                }
                for (int istag = 1; istag <= Nstag; ++istag) for (int ip = 1; ip <= Np; ++ip) if (ip != iworm) {
                    b.try_stag = b.try_stag + 1;
                    if (sampling == 0) { MoveHead(Lstag, ip, acc_head); MoveTail(Lstag, ip, acc_tail); Staging(Lstag, ip, acc_bd); }
                    else { MoveHeadBisection(Nlev, ip, acc_head); MoveTailBisection(Nlev, ip, acc_tail); Bisection(Nlev, ip, acc_bd); }
                }
                for (int iobdm = 1; iobdm <= Nobdm; ++iobdm) {
                    int ip = iworm;
                    for (int j = 1; j <= 2; ++j) { b.try_cm_half = b.try_cm_half + 1; TranslateHalfChain(j, delta_cm, ip, acc_cm_half); }
                    for (int j = 1; j <= 2; ++j) {
                        b.try_stag_half = b.try_stag_half + 1;
                        MoveHeadHalfChain(j, Lstag, ip, acc_head_half);
                        MoveTailHalfChain(j, Lstag, ip, acc_tail_half);
                        StagingHalfChain(j, Lstag, ip, acc_bd_half);
                    }
                    if (swapping) {
                        b.try_swap = b.try_swap + 1;
                        Swap(Lstag, ip, acc_swap, ik, swap_accepted);
                        PermutationSampling(iworm, true, ik, swap_accepted);
                    }
                    if (!trap) {
                        double xe[6];
                        for (int j = 0; j < 2; ++j) for (int k = 0; k < dim; ++k) xe[k + dim * j] = xend[j][k];
                        OBDM(xe, nrho);
                    }
                }
            } else {
                idiag = idiag + 1; idiag_aux = idiag_aux + 1; b.idiag_block = b.idiag_block + 1;
                if (istep % CMFreq == 0) {
                    for (int ip = 1; ip <= Np; ++ip) { b.try_cm = b.try_cm + 1; TranslateChain(delta_cm, ip, acc_cm); }
                }
                for (int istag = 1; istag <= Nstag; ++istag) for (int ip = 1; ip <= Np; ++ip) {
                    b.try_stag = b.try_stag + 1;
                    if (sampling == 0) { MoveHead(Lstag, ip, acc_head); MoveTail(Lstag, ip, acc_tail); Staging(Lstag, ip, acc_bd); }
                    else { MoveHeadBisection(Nlev, ip, acc_head); MoveTailBisection(Nlev, ip, acc_tail); Bisection(Nlev, ip, acc_bd); }
                }
                double E1, E2, E, Kin, Pot, Et, Kt;
                LocalEnergy(Path.data(), E1, Kin, Pot);
                LocalEnergy(Path.data() + ss * (2 * Nb), E2, Kin, Pot);
                E = 0.5 * (E1 + E2);
                ThermEnergy(Path.data(), dt, Et, Kt, Pot);
                Kin = E - Pot;
                b.sumE += E; b.sumK += Kin; b.sumV += Pot;
                b.sumEt += Et; b.sumKt += Kt; b.sumVt += Pot;
                b.sumE2 += E * E; b.sumK2 += Kin * Kin; b.sumV2 += Pot * Pot;
                b.sumEt2 += Et * Et; b.sumKt2 += Kt * Kt; b.sumVt2 += Pot * Pot;
                b.ngr = b.ngr + 1;
                if (!trap) {
                    PairCorrelation(Path.data() + ss * Nb, gr);
                    StructureFactor(Nk, Path.data() + ss * Nb, Sk);
                }
            }
        }
        b.acc_cm = acc_cm; b.acc_bd = acc_bd; b.acc_head = acc_head; b.acc_tail = acc_tail;
        b.acc_cm_half = acc_cm_half; b.acc_bd_half = acc_bd_half; b.acc_head_half = acc_head_half; b.acc_tail_half = acc_tail_half;
        b.acc_open = acc_open; b.acc_close = acc_close; b.acc_swap = acc_swap;
        b.idiag_aux = idiag_aux;
        for (int c = 0; c < 3; ++c) b.bead_updates[c] = nupd[c] - n0[c];
    }
};

/* ================================================================ C interface */
extern "C" {

orc_sim* orc_create(const orc_params* p) {
    orc_sim* s = new orc_sim();
    s->dim = p->dim; s->Np = p->Np; s->density = p->density;
    s->crystal = p->crystal != 0; s->trap = p->trap != 0;
    s->dt = p->dt; s->Nb = p->Nb; s->seed = p->seed; s->delta_cm = p->delta_cm; s->CMFreq = p->CMFreq;
    s->sampling = p->sampling; s->Lstag = p->Lstag; s->Nlev = p->Nlev; s->Nstag = p->Nstag;
    s->Nbin = p->Nbin; s->Nk = p->Nk; s->swapping = p->swapping != 0; s->CWorm = p->CWorm;
    s->Nobdm = p->Nobdm; s->Npw = p->Npw; s->Nmax = p->Nmax;
    s->wf_table = p->wf_table != 0; s->v_table = p->v_table != 0; s->Rm = p->Rm; s->action = p->action;
    for (int k = 0; k < 3; ++k) { s->a_ho[k] = p->a_ho[k]; s->Lbox[k] = s->LboxHalf[k] = s->qbin[k] = 0.0; }
    const int dim = s->dim;
    /* vpi.f90:80-128 */
    s->pi = std::acos(-1.0);
    if (s->trap) {
        s->rcut = 1.0;
        for (int k = 0; k < dim; ++k) s->rcut = 3.0 * s->rcut * s->a_ho[k];
        s->density = (double)(float)s->Np / (std::pow(s->pi, 0.5 * dim) * s->rcut / std::tgamma(0.5 * dim + 1.0));
        s->rcut = std::pow(s->rcut, 1.0 / (double)(float)dim);
        s->rcut = 10.0 * s->rcut;
        double amin = s->a_ho[0];
        for (int k = 1; k < dim; ++k) amin = std::fmin(amin, s->a_ho[k]);
        s->delta_cm = s->delta_cm * amin;
        /* Lbox is unallocated in the reference's trap mode; keep a huge box so wraps never fire */
        for (int k = 0; k < 3; ++k) { s->Lbox[k] = 1e300; s->LboxHalf[k] = 0.5e300; s->qbin[k] = 0.0; }
    } else {
        if (s->crystal) {
            for (int k = 0; k < dim; ++k) s->Lbox[k] = p->Lbox_crystal[k];
        } else {
            for (int k = 0; k < dim; ++k) s->Lbox[k] = std::pow((double)(float)s->Np / s->density, 1.0 / (double)(float)dim);
        }
        for (int k = 0; k < dim; ++k) { s->LboxHalf[k] = 0.5 * s->Lbox[k]; s->qbin[k] = 2.0 * s->pi / s->Lbox[k]; }
        s->rcut = s->LboxHalf[0];
        for (int k = 1; k < dim; ++k) s->rcut = std::fmin(s->rcut, s->LboxHalf[k]);
        s->delta_cm = s->delta_cm / std::pow(s->density, 1.0 / (double)(float)dim);
    }
    s->rcut2 = s->rcut * s->rcut;
    s->rbin = s->rcut / (double)(float)s->Nbin;
    s->dr = s->rcut / (double)(float)(s->Nmax - 1);          /* vpi_mod.f90:94,127 */
    s->isopen = false; s->iworm = 0;
    s->Path.assign((size_t)dim * s->Np * (2 * s->Nb + 1), 0.0);
    s->LogWF.assign((size_t)s->Nmax + 2, 0.0);
    s->VTable.assign((size_t)s->Nmax + 2, 0.0);
    s->Particles_in_perm_cycle.assign((size_t)s->Np, 0);
    s->Perm_histogram.assign((size_t)s->Np, 0);
    for (int j = 0; j < 2; ++j) for (int k = 0; k < 3; ++k) s->xend[j][k] = 0.0;
    return s;
}
void orc_destroy(orc_sim* s) { delete s; }

void orc_get_geometry(const orc_sim* s, double* Lbox3, double* rcut, double* dr, double* rbin, double* density, double* dcm) {
    for (int k = 0; k < 3; ++k) Lbox3[k] = s->Lbox[k];
    *rcut = s->rcut; *dr = s->dr; *rbin = s->rbin; *density = s->density; *dcm = s->delta_cm;
}

void orc_fill_tables(orc_sim* s) {
    /* vpi_mod.f90:84-112 */
    for (int i = 1; i <= s->Nmax; ++i) { double r = (i - 1) * s->dr; s->LogWF[i] = LogPsi(0, s->Rm, r); }
    s->LogWF[0] = s->LogWF[2]; s->LogWF[s->Nmax + 1] = s->LogWF[s->Nmax];
    /* vpi_mod.f90:116-145 */
    for (int i = 1; i <= s->Nmax; ++i) { double r = (i - 1) * s->dr; s->VTable[i] = Potential(r); }
    s->VTable[0] = s->VTable[2]; s->VTable[s->Nmax + 1] = s->VTable[s->Nmax];
}
void orc_set_tables(orc_sim* s, const double* W, const double* V) {
    std::memcpy(s->LogWF.data(), W, sizeof(double) * (s->Nmax + 2));
    std::memcpy(s->VTable.data(), V, sizeof(double) * (s->Nmax + 2));
}
void orc_get_tables(const orc_sim* s, double* W, double* V) {
    std::memcpy(W, s->LogWF.data(), sizeof(double) * (s->Nmax + 2));
    std::memcpy(V, s->VTable.data(), sizeof(double) * (s->Nmax + 2));
}

void orc_init(orc_sim* s, const double* R0) {               /* vpi_mod.f90:187-256 */
    const int dim = s->dim, Np = s->Np;
    std::vector<double> R((size_t)dim * Np);
    s->rng.sgrnd((uint32_t)s->seed);
    if (s->trap) {
        for (int ip = 0; ip < Np; ++ip) for (int k = 0; k < dim; ++k) R[k + dim * ip] = 2.0 * s->a_ho[k] * (s->rng.grnd() - 0.5);
    } else if (s->crystal) {
        for (size_t i = 0; i < R.size(); ++i) R[i] = R0[i];
    } else {
        for (int ip = 0; ip < Np; ++ip) for (int k = 0; k < dim; ++k) R[k + dim * ip] = s->Lbox[k] * (s->rng.grnd() - 0.5);
    }
    for (int ib = 0; ib <= 2 * s->Nb; ++ib) for (int ip = 1; ip <= Np; ++ip) for (int k = 1; k <= dim; ++k) s->P(k, ip, ib) = R[(k - 1) + dim * (ip - 1)];
    for (int j = 0; j < 2; ++j) for (int k = 1; k <= dim; ++k) s->xend[j][k - 1] = s->P(k, Np, s->Nb);
}
void orc_set_state(orc_sim* s, const double* Path, const double* xend, int isopen, int iworm) {
    std::memcpy(s->Path.data(), Path, sizeof(double) * s->Path.size());
    for (int j = 0; j < 2; ++j) for (int k = 0; k < s->dim; ++k) s->xend[j][k] = xend[k + s->dim * j];
    s->isopen = isopen != 0; s->iworm = iworm;
}
void orc_get_state(const orc_sim* s, double* Path, double* xend, int* isopen, int* iworm) {
    std::memcpy(Path, s->Path.data(), sizeof(double) * s->Path.size());
    for (int j = 0; j < 2; ++j) for (int k = 0; k < s->dim; ++k) xend[k + s->dim * j] = s->xend[j][k];
    *isopen = s->isopen ? 1 : 0; *iworm = s->iworm;
}
void orc_set_perm(orc_sim* s, int iperm, const int32_t* cyc, const int32_t* hist, int npc, int epc) {
    s->iperm = iperm; s->new_perm_cycle = npc != 0; s->end_perm_cycle = epc != 0;
    for (int i = 0; i < s->Np; ++i) { s->Particles_in_perm_cycle[i] = cyc[i]; s->Perm_histogram[i] = hist[i]; }
}
void orc_get_perm(const orc_sim* s, int* iperm, int32_t* cyc, int32_t* hist, int* npc, int* epc) {
    *iperm = s->iperm; *npc = s->new_perm_cycle; *epc = s->end_perm_cycle;
    for (int i = 0; i < s->Np; ++i) { cyc[i] = s->Particles_in_perm_cycle[i]; hist[i] = s->Perm_histogram[i]; }
}

void   orc_sgrnd(orc_sim* s, int32_t seed) { s->rng.sgrnd((uint32_t)seed); }
double orc_grnd(orc_sim* s) { return s->rng.grnd(); }
double orc_rangauss(orc_sim* s) { return s->rng.rangauss(); }
uint32_t orc_mt_raw(orc_sim* s) { return s->rng.raw(); }
void orc_get_mt(const orc_sim* s, uint32_t* mt, int32_t* mti) { std::memcpy(mt, s->rng.mt, sizeof s->rng.mt); *mti = s->rng.mti; }
void orc_set_mt(orc_sim* s, const uint32_t* mt, int32_t mti) { std::memcpy(s->rng.mt, mt, sizeof s->rng.mt); s->rng.mti = mti; }

double orc_interpolate(int opt, int N, double dx, const double* F, double x) { return Interpolate(opt, N, dx, F, x); }
double orc_potential(double r) { return Potential(r); }
double orc_logpsi(int opt, double Rm, double r) { return LogPsi(opt, Rm, r); }
double orc_green_function(const orc_sim* s, int opt, int ib, double dt, double Pot, double F2) { return s->GreenFunction(opt, ib, dt, Pot, F2); }
void   orc_minimum_image(const orc_sim* s, double* xij, double* rij2) { s->MinimumImage(xij, *rij2); }
double orc_boundary_conditions(const orc_sim* s, int k, double x) { s->BoundaryConditions(k, x); return x; }

double orc_update_action(orc_sim* s, int ip, int ib, const double* xnew, const double* xold) {
    double dS; s->UpdateAction(ip, ib, xnew, xold, s->dt, dS); return dS;
}
double orc_update_action_R(orc_sim* s, const double* R, int ip, int ib, const double* xnew, const double* xold) {
    double dS; s->UpdateActionR(R, ip, ib, xnew, xold, s->dt, dS); return dS;
}

int orc_move(orc_sim* s, int move, int ip, int half, int* aux) {
    int acc = 0;
    switch (move) {
    case ORC_TRANSLATE_CHAIN: s->TranslateChain(s->delta_cm, ip, acc); break;
    case ORC_STAGING: s->Staging(s->Lstag, ip, acc); break;
    case ORC_MOVE_HEAD: s->MoveHead(s->Lstag, ip, acc); break;
    case ORC_MOVE_TAIL: s->MoveTail(s->Lstag, ip, acc); break;
    case ORC_BISECTION: s->Bisection(s->Nlev, ip, acc); break;
    case ORC_MOVE_HEAD_BISECTION: s->MoveHeadBisection(s->Nlev, ip, acc); break;
    case ORC_MOVE_TAIL_BISECTION: s->MoveTailBisection(s->Nlev, ip, acc); break;
    case ORC_TRANSLATE_HALF: s->TranslateHalfChain(half, s->delta_cm, ip, acc); break;
    case ORC_STAGING_HALF: s->StagingHalfChain(half, s->Lstag, ip, acc); break;
    case ORC_MOVE_HEAD_HALF: s->MoveHeadHalfChain(half, s->Lstag, ip, acc); break;
    case ORC_MOVE_TAIL_HALF: s->MoveTailHalfChain(half, s->Lstag, ip, acc); break;
    case ORC_OPEN: s->iworm = ip; s->OpenChain(s->Lstag, ip, acc); break;
    case ORC_CLOSE: s->CloseChain(s->Lstag, ip, acc); break;
    case ORC_SWAP: {
        int ipar = 0; bool sa = false;
        s->Swap(s->Lstag, ip, acc, ipar, sa);
        s->swap_accepted = sa; if (sa) s->ik = ipar;
        if (aux) *aux = sa ? ipar : 0;
        break; }
    default: break;
    }
    return acc;
}

void orc_potential_energy(orc_sim* s, const double* R, int want_f2, double* Pot, double* F2) {
    if (want_f2) s->PotentialEnergy(R, *Pot, F2); else { s->PotentialEnergy(R, *Pot, nullptr); if (F2) *F2 = 0.0; }
}
void orc_local_energy(orc_sim* s, const double* R, double* E, double* Kin, double* Pot) { s->LocalEnergy(R, *E, *Kin, *Pot); }
void orc_therm_energy(orc_sim* s, double* E, double* Ec, double* Ep) { s->ThermEnergy(s->Path.data(), s->dt, *E, *Ec, *Ep); }
void orc_therm_energy_P(orc_sim* s, const double* Path, double* E, double* Ec, double* Ep) { s->ThermEnergy(Path, s->dt, *E, *Ec, *Ep); }
void orc_pair_correlation(orc_sim* s, const double* R, double* gr) { s->PairCorrelation(R, gr); }
void orc_structure_factor(orc_sim* s, const double* R, double* Sk) { s->StructureFactor(s->Nk, R, Sk); }
void orc_obdm(orc_sim* s, const double* xend, double* nrho) { s->OBDM(xend, nrho); }

/* sample_mod.f90:656-679 */
void orc_normalize_gr(orc_sim* s, int ngr, double* gr) {
    double k_n = std::pow(s->pi, 0.5 * s->dim) / std::tgamma(0.5 * s->dim + 1.0);
    double norm = (double)((float)s->Np * (float)ngr);
    for (int ibin = 1; ibin <= s->Nbin; ++ibin) {
        double r = ((double)(float)ibin - 0.5) * s->rbin;
        double nid = s->density * k_n * (ipow(r + 0.5 * s->rbin, s->dim) - ipow(r - 0.5 * s->rbin, s->dim));
        gr[ibin - 1] = gr[ibin - 1] / (nid * norm);
    }
}
/* sample_mod.f90:683-702 */
void orc_normalize_sk(orc_sim* s, int ngr, double* Sk) {
    double norm = (double)((float)s->Np * (float)ngr);
    for (int i = 0; i < s->dim * s->Nk; ++i) Sk[i] = Sk[i] / norm;
}
/* sample_mod.f90:706-732 */
void orc_normalize_nr(orc_sim* s, double zconf, double* nrho) {
    double k_n = std::pow(s->pi, 0.5 * s->dim) / std::tgamma(0.5 * s->dim + 1.0);
    for (int ibin = 1; ibin <= s->Nbin; ++ibin) {
        double r = ((double)(float)ibin - 0.5) * s->rbin;
        double nid = s->density * k_n * (ipow(r + 0.5 * s->rbin, s->dim) - ipow(r - 0.5 * s->rbin, s->dim));
        for (int m = 0; m <= s->Npw; ++m)
            nrho[m + (s->Npw + 1) * (ibin - 1)] = nrho[m + (s->Npw + 1) * (ibin - 1)] / (s->CWorm * nid * zconf * (double)(float)s->Nobdm);
    }
}
/* sample_mod.f90:921-932 */
double orc_var(int Nitem, double Sum, double Sum2) { return std::sqrt((Sum2 - Sum * Sum) / (double)(float)Nitem); }

void orc_run_block(orc_sim* s, int Nstep, orc_block* out, double* gr, double* Sk, double* nrho) { s->RunBlock(Nstep, *out, gr, Sk, nrho); }
void orc_bead_updates(const orc_sim* s, uint64_t* three) { for (int c = 0; c < 3; ++c) three[c] = s->nupd[c]; }

} /* extern "C" */
