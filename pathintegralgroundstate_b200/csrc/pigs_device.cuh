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
// a warp, partner j <-> lane, and reduced with warp shuffles; beads of one move
// are spread over the warps of the group.
//
// Code-size discipline (ncu, round 1: the first version inlined everything and
// stalled 10 of 22 warp-cycles per issue on instruction fetch, and kept its
// context object in local memory): run parameters live in __constant__ memory
// (they fold into FP64 instruction operands), per-group state lives in shared
// memory, and every shared helper is a real (noinline) function so the 14 moves
// are thin glue around one copy of the hot code.
//
// Data layout in HBM (per chain): path[ib][ip/32][k][ip%32] -- a time slice is NpS/32 blocks of [3][32] doubles
// (x, y, z of 32 consecutive particles side by side, 768 contiguous bytes), so lane<->partner loads are
// unit-stride and one block is one DRAM burst (see pidx below).  The ABI layout Path(dim,Np,0:2Nb) is
// array-of-structures; the library transposes on upload/download.  Unused components (dim<3) are zero.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

// partner-loop unroll factor of the hot loops (register budget permitting)
#ifndef PIGS_UNROLL
#define PIGS_UNROLL 1
#endif
#define PIGS_STR2(x) #x
#define PIGS_STR(x) PIGS_STR2(x)
#define PIGS_PRAGMA_UNROLL _Pragma(PIGS_STR(unroll PIGS_UNROLL))

namespace pigs {

constexpr int MAXS = 132;   // 2*Nb+1 <= MAXS (Nb <= 65)
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
    int chain_offset;         // global index of chain 0 (chains of one run sharded over several handles)
    double L[3], Lh[3], invL[3], qbin[3], a_ho[3];
    double rcut2, dr, inv_dr, rbin, dt, delta_cm, CWorm, density, pi, logCd;
    // GreenFunction (global_mod.f90:19-72) as weights by slice class [even, odd, end]: action (opt 0) and its
    // dt-derivative (opt 1); cF/cFE multiply |F|^2.  Chin: 2dt/3, 4dt/3, dt/3, cF = 4dt^3/18; primitive: dt, cF = 0.
    double wS[3], cF, wE[3], cFE;
    int primitive;
    // proposal widths, precomputed on the host with the reference's arithmetic (no sqrt/div in the kernel):
    //   sig_free[L]  = sqrt(real(L)*dt)                        free end of a length-L move  (vpi_mod.f90:632)
    //   sig_stage[n] = sqrt((real(n)/real(n+1))*dt), float32 ratio (Q15)                    (vpi_mod.f90:531)
    //   sig_bis[l]   = sqrt(0.5*(0.5*real(2**l)*dt))           bisection, delta_ib = 2**l     (vpi_mod.f90:906-907)
    double sig_free[MAXS], sig_stage[MAXS], sig_bis[16];
    double half_inv_dt2;
    double half_inv_dr;       // 0.5/dr (centred difference of the table, cell coordinates)
    double rclamp2;           // ((Nmax+3.5)*dr)^2: where masked pairs are looked up (zero tail of the tables)
    float LhF[3];             // high word of L/2 reinterpreted as float (mimg_hi)
    int tabW_off;             // byte offset of the LogWF copy behind the VTable copy in shared memory
    unsigned long long seed;
    const double* logwf;      // (0:Nmax+1) in global memory
    const double* vtab;
    double* path;             // [chain][ib][NpS/32][k][32]  (see pidx)
    double* xend;             // [chain][2][3]
    int* istate;              // [chain][IS_N]
    int* cyc;                 // [chain][Np]  Particles_in_perm_cycle
    int* hist;                // [chain][Np]  Perm_histogram
    unsigned* mt;             // [chain][624]
    unsigned long long* pctr; // [chain][PCS] Philox counters
    double* pp;               // [chain][Np]  Swap's tower-sampling weights (vpi_mod.f90:2334-2345): rare, so not in smem
    double* acc;              // [chain][nacc]
    long long* cnt;           // [chain][NCNT]
    size_t chain_stride;      // doubles per chain in path
    int nacc;                 // NE + Nbin + dim*Nk + Nbin*(Npw+1)
    int off_gr, off_sk, off_nr;
};

// what the persistent kernel is asked to do
enum SweepOp { OP_BLOCK = 0, OP_MOVE = 1, OP_UNIFORM = 2, OP_GAUSS = 3, OP_SEED = 4 };
struct SweepArgs {
    int op;
    int nstep;           // OP_BLOCK: steps; OP_UNIFORM/OP_GAUSS: draws
    int move, ip0, half; // OP_MOVE
    int chain_only;      // >=0: only that chain (OP_UNIFORM/OP_GAUSS/OP_SEED)
    int seed;            // OP_SEED
    int groups_per_cta, threads_per_chain, tshift;   // T = 1 << tshift
    int team;            // 1: T = 128 and the four warps of a group sweep disjoint slice windows concurrently (Philox only)
    int prefetch;        // 0 off, 1 L1 line prefetch at move start, 2 L2 bulk per phase, 3 L2 bulk rolling (default)
    int pfdist;          // rolling distance in beads (PIGS_PFDIST)
    int* accepted;       // OP_MOVE [n_chains]
    int* aux;            // OP_MOVE [n_chains]
    double* draws;       // OP_UNIFORM/OP_GAUSS [n]
};

// one copy per translation unit; the TU's launcher uploads them (stream-ordered)
// right before its kernel launch
static __constant__ DevParams cP;
static __constant__ SweepArgs cA;

// per-group state in shared memory: header + arrays
// Move descriptor + RNG state parked in shared memory across the partner loop
// (run_move, pigs_sweep.cuh); ctr ping-pongs by phase parity.
struct RngS {
    unsigned long long ctr;
    unsigned w0, w1, w2;
    int nleft;
    unsigned tag;        // Philox stream selector: 0 the chain's stream, 1 + w the stream of window worker w
    unsigned pad;
};
constexpr int PCS = 8;   // Philox counters kept per chain: [0] chain stream, [1 + w] window workers
struct MovePark {
    RngS ctr[2];
    double Sbase, DeltaK;
    int flags, ip0, ii, ie, m0, m1, pad0, pad1;
};
struct GS {
    double* path;        // this chain
    double* xend;        // this chain, [2][3]
    double* acc;         // this chain's accumulators (global)
    int* cyc;
    int* hist;
    unsigned* mt;
    double* pp;          // Swap's Pp(Np) scratch, global memory
    const double* tabV;  // table bases: shared-memory copies when staged, else global
    const double* tabW;
    double bc[8];        // broadcast slots
    double kc[4];        // 1/L[0..2], 1/dr: the partner loop's constants as plain data (PIGS_KC)
    double eacc[NE];
    long long cnt[NCNT];
    int ibc[8];
    // chain state (written by thread 0, read by all after a group sync)
    int mti, isopen, iworm0, iperm, new_pc, end_pc, ik0, swap_acc, idiag_aux, chain;
    // Geometry of the thread group that currently works on this block (set by thread 0 of the owner, read after a
    // sync): normally the chain group of T threads; in team mode one warp (a window worker) during the sweeps.
    int gsize, gshift, gbar;
    // windows of the current pass of the window-ordered sweep (win_sweep): head / tail lengths, middle starts
    int wLh, wLt, wM0, wM1, wnM;
    MovePark pk;
    // followed by: seg_old[3S] seg_new[3S] part[np*8]
};
// ------------------------------------------------------------------ group
struct Grp {
    int tid, lane, warp, nwarps, size;
};
// the size of the group working on block gs: with one warp per chain (the production shape) it is a launch
// constant -- no shared-memory read on the serial path of a move (measured: -8 % at N = 64 when every sync looked it up)
__device__ __forceinline__ int gsize_of(const GS* gs) { return cA.threads_per_chain == 32 ? 32 : gs->gsize; }
__device__ __forceinline__ Grp grp(const GS* gs) {
    Grp g;
    g.size = gsize_of(gs);
    g.tid = threadIdx.x & (g.size - 1);
    g.lane = threadIdx.x & 31;
    g.warp = g.tid >> 5;
    g.nwarps = g.size >> 5;
    return g;
}
__device__ __forceinline__ void gsync(const GS* gs) {
    if (cA.threads_per_chain == 32) { __syncwarp(); return; }
    const int n = gs->gsize;
    if (n == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(gs->gbar), "r"(n) : "memory");
}
// all T threads of the chain group, whatever the current working geometry (team mode: the four window workers)
__device__ __forceinline__ void tsync() {
    if (cA.threads_per_chain == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"((int)(threadIdx.x >> cA.tshift)), "r"(cA.threads_per_chain) : "memory");
}
__device__ __forceinline__ bool gfirst(const GS* gs) { return (threadIdx.x & (gsize_of(gs) - 1)) == 0; }
__host__ __device__ inline int part_slots(int nwarps) { return nwarps < 4 ? 4 : nwarps; }
__host__ __device__ inline size_t grp_smem_bytes(int S, int Np, int nwarps) {
    (void)Np;
    return sizeof(GS) + sizeof(double) * ((size_t)6 * S + (size_t)part_slots(nwarps) * 8);
}
__device__ __forceinline__ double* seg_old(GS* gs) { return reinterpret_cast<double*>(gs + 1); }
__device__ __forceinline__ double* seg_new(GS* gs) { return seg_old(gs) + 3 * cP.S; }
__device__ __forceinline__ double* part_of(GS* gs) { return seg_old(gs) + 6 * cP.S; }
__device__ __forceinline__ double* pp_of(GS* gs) { return gs->pp; }
__device__ __forceinline__ double& so(GS* gs, int k, int ib) { return seg_old(gs)[k * cP.S + ib]; }
__device__ __forceinline__ double& sn(GS* gs, int k, int ib) { return seg_new(gs)[k * cP.S + ib]; }
// A time slice is stored as NpS/32 blocks of [3][32] doubles: x, y, z of 32
// consecutive particles side by side (768 contiguous bytes).  A warp of the
// partner loop reads one block per iteration: one pointer, immediate offsets
// for y and z, and one DRAM burst instead of three pieces 8*NpS bytes apart.
// NpS = Np rounded up to 32; the padding of the last block is zero and never read.
constexpr int PY = 32, PZ = 64, PBLK = 96;
// Device tables carry TAB_PAD zero entries behind the reference's F(0:Nmax+1): a pair that is outside the
// cutoff (or is the moved particle itself) is looked up at r_clamp = (Nmax+3.5)*dr, where value and centred
// difference read only zeros -- its contribution vanishes without a select on every accumulated term.
constexpr int TAB_PAD = 4;
__host__ __device__ __forceinline__ int tab_len(int Nmax) { return Nmax + 2 + TAB_PAD; }
__host__ __device__ __forceinline__ int pidx(int j) { return (j >> 5) * PBLK + (j & 31); }
__device__ __forceinline__ double* slice(GS* gs, int ib) { return gs->path + (size_t)ib * 3 * cP.NpS; }
__device__ __forceinline__ double& pth(GS* gs, int k, int ip0, int ib) { return gs->path[(size_t)ib * 3 * cP.NpS + pidx(ip0) + 32 * k]; }

// ------------------------------------------------------------------ tables
// VAR 0: both tables through the read-only L1/L2 path; 1: VTable in shared
// memory as a PAIR table {F(i),F(i+1)} read with one LDS.128 per interpolant,
// LogWF global; 2: both plain in shared memory; 3: trap (tables global).
template <int VAR> struct VarTraits;
template <> struct VarTraits<0> { static constexpr bool TRAP = false, VSM = false, WSM = false, VPAIR = false; };
template <> struct VarTraits<1> { static constexpr bool TRAP = false, VSM = true,  WSM = false, VPAIR = true;  };
template <> struct VarTraits<2> { static constexpr bool TRAP = false, VSM = true,  WSM = true,  VPAIR = false; };
template <> struct VarTraits<3> { static constexpr bool TRAP = true,  VSM = false, WSM = false, VPAIR = false; };

// Table reads with the address space known at compile time: the staged copies
// sit at the start of dynamic shared memory (VTable first, then LogWF), so the
// loads are LDS with no generic-address detour.  WHICH: 0 VTable, 1 LogWF.
template <bool SM, int WHICH, bool VSM_FIRST>
__device__ __forceinline__ double tab(int i) {
    if (SM) {
        extern __shared__ __align__(16) double pigs_smem_base[];
        const int off = (WHICH == 1 && VSM_FIRST) ? tab_len(cP.Nmax) : 0;
        return pigs_smem_base[off + i];
    }
    return __ldg((WHICH == 0 ? cP.vtab : cP.logwf) + i);
}
template <bool SM>
__device__ __forceinline__ double tld(const double* t, int i) {
    if (SM) return t[i];
    return __ldg(t + i);
}
// 1/sqrt(x) for normal positive x: hardware seed (MUFU.RSQ64H, ~20 bits) and one
// cubically convergent step -- 5 FP64 issue slots, no slow-path call inside the
// partner loop.  Relative error < 2^-52 * 2.
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x * y, y, 1.0);                 // 1 - x y^2
    double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);                        // y (1 + e/2 + 3e^2/8)
}
// sqrt(x) by the same step applied to x*y: one slot and one dependent level less than x * rsqrt_pos(x)
__device__ __forceinline__ double sqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double r0 = x * y;
    double e = fma(-r0, y, 1.0);
    double p = fma(0.375, e, 0.5);
    return fma(r0 * e, p, r0);
}

// Interpolate(opt,...) of interpolate.f90:1-45, fast form for the hot loop:
// x/dx -> x*inv_dx and floor() by the 2^52 trick (no F2I/I2F conversions).
// Every interpolant is continuous in x, so an index flip at a grid point is
// harmless.
struct Lk {
    int i0;          // = ix-1
    double t;        // aux1/dx of interpolate.f90:14: position inside the cell, [0,1)
};
__device__ __forceinline__ Lk lk_prep(double r) {
    const double MAGIC = 6755399441055744.0;                 // 2^52 + 2^51
    double m = __fma_rd(r, cP.inv_dr, MAGIC);                // MAGIC + floor(s): the round-down FMA floors in one slot
    Lk k;
    k.i0 = max(__double2loint(m), 1);                        // r >= dr always in practice; keeps i0-1 in range
    k.t = fma(r, cP.inv_dr, -(m - MAGIC));                   // s - floor(s), one rounding
    return k;
}
// The interpolants in cell coordinates -- (aux1 F(ix) + aux2 F(ix-1))/dx == F(ix-1) + t (F(ix) - F(ix-1)) --
// two FP64 slots per value instead of four and no aux2; the hot loop is two thirds FP64-pipe bound (ncu).
// VTable staged as pairs: entry i holds {F(i), F(i+1)} (16-byte aligned), so a linear
// interpolant is ONE shared-memory request instead of two (ncu: 2 bank-conflict
// cycles per LDS.64 on the random table indices of a warp).
__device__ __forceinline__ double2 tabpair(int i) {
    extern __shared__ __align__(16) double pigs_smem_base[];
    return reinterpret_cast<const double2*>(pigs_smem_base)[i];
}
__device__ __forceinline__ double lk_val_pair(const Lk& k) {
    double2 t = tabpair(k.i0);
    return fma(k.t, t.y - t.x, t.x);
}
// value and centred first difference (opt 0 and 1): d1 = (Fa - Fb)/(2 dx) with Fa, Fb the interpolants one
// cell up and one cell down
__device__ __forceinline__ void lk_d1_from(double t, double fm, double f0, double f1, double f2, double& v, double& d1) {
    v = fma(t, f1 - f0, f0);
    double b = fma(t, (f2 + fm) - (f1 + f0), f1 - fm);       // Fa - Fb = (f1-fm) + t ((f2-f1) - (f0-fm))
    d1 = b * cP.half_inv_dr;
}
__device__ __forceinline__ void lk_val_d1_pair(const Lk& k, double& v, double& d1) {
    double2 lo = tabpair(k.i0 - 1), hi = tabpair(k.i0 + 1);      // {fm,f0}, {f1,f2}
    lk_d1_from(k.t, lo.x, lo.y, hi.x, hi.y, v, d1);
}
template <bool SM, int WHICH, bool VF>
__device__ __forceinline__ double lk_val(const Lk& k) {   // opt 0
    double f0 = tab<SM, WHICH, VF>(k.i0), f1 = tab<SM, WHICH, VF>(k.i0 + 1);
    return fma(k.t, f1 - f0, f0);
}
template <bool SM, int WHICH, bool VF>
__device__ __forceinline__ void lk_val_d1(const Lk& k, double& v, double& d1) {   // opt 0 and 1
    double fm = tab<SM, WHICH, VF>(k.i0 - 1), f0 = tab<SM, WHICH, VF>(k.i0), f1 = tab<SM, WHICH, VF>(k.i0 + 1), f2 = tab<SM, WHICH, VF>(k.i0 + 2);
    lk_d1_from(k.t, fm, f0, f1, f2, v, d1);
}

// Interpolate opt 0, 1 and 2 in the reference's exact operation order, true
// divisions, no FMA contraction.  The second difference divides rounding noise
// by dx^2 (~1e7): only the same arithmetic reproduces the reference's value to
// 1e-10, so the mixed estimator (2 calls per MC step, not the hot loop) pays
// for IEEE divisions here.
template <bool SM>
__device__ __forceinline__ void lk_exact_d1_d2(const double* F, double x, double& d1, double& d2) {
    const double dx = cP.dr;
    int ix = (int)__ddiv_rn(x, dx) + 1;
    ix = max(2, min(ix, cP.Nmax));
    double aux1 = __dadd_rn(x, -__dmul_rn((double)(ix - 1), dx));
    double aux2 = __dadd_rn(dx, -aux1);
    double fm = tld<SM>(F, ix - 2), f0 = tld<SM>(F, ix - 1), f1 = tld<SM>(F, ix), f2 = tld<SM>(F, ix + 1);
    double Fb = __ddiv_rn(__dadd_rn(__dmul_rn(aux1, f0), __dmul_rn(aux2, fm)), dx);
    double Fc = __ddiv_rn(__dadd_rn(__dmul_rn(aux1, f1), __dmul_rn(aux2, f0)), dx);
    double Fa = __ddiv_rn(__dadd_rn(__dmul_rn(aux1, f2), __dmul_rn(aux2, f1)), dx);
    // the centred first difference loses ~log10(1/dx) digits only: a multiplication by 1/dx is enough
    d1 = 0.5 * (Fa - Fb) * cP.inv_dr;
    d2 = __ddiv_rn(__dadd_rn(__dadd_rn(Fa, -__dmul_rn(2.0, Fc)), Fb), __dmul_rn(dx, dx));
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

// Hot-loop form: d - L*rint(d/L) with rint() by the 2^52 trick (3 FP64 issue
// slots instead of 4 + 4 selects).  Identical to mimg() for |d| < 1.5 L except
// exactly at |d| = L/2, where either image gives the same r^2.
// 1.5 * 2^52: adding and subtracting it rounds to the nearest integer.  (Keeping it in a register pair so that 1/L
// could be a constant-bank operand of the FMA instead of an LDC per iteration is not expressible: ptxas folds any
// materialisation of the constant back into an immediate.)
__device__ __forceinline__ double magic52() { return 6755399441055744.0; }
// XR ("reference rounding") instances of the pair loop -- the MT19937 replay kernels and the unit entry, i.e. wherever
// the results must be the reference's own -- take the reference's decisions bit for bit; the Philox production kernel
// keeps the faster forms (measured: the exact forms cost 3.6 % at C3, 6.5 % at C2), which sample the same distribution:
// the two differ on a set of measure zero.
// The reference's decision, d > L/2 / d < -L/2 on the separation itself, as two compares and one FMA: the same three
// FP64 slots as the 2^52 form.  It differs from d - L*rint(d/L) only when |d| sits within an ulp or two of L/2, where
// the rounded product d*(1/L) and the comparison can disagree; both images then have the same |d| to an ulp, but a
// pair that lies EXACTLY on the cutoff sphere (rcut = L/2 along the shortest box edge, zero transverse separation:
// the aligned neighbours of a perfect crystal lattice, BASELINE configs[3] started from config_ini.in) is inside
// the cutoff for one image and outside for the other, and its force component changes sign.
__device__ __forceinline__ double mimg_cmp(double d, double L, double Lh) {
    const int qhi = (d > Lh) ? 0x3ff00000 : ((d < -Lh) ? (int)0xbff00000 : 0);
    return fma(-__hiloint2double(qhi, 0), L, d);
}
// rij2 exactly as MinimumImage accumulates it (pbc_mod.f90:29-52): three rounded products, added in order.  With
// fused multiply-adds the sum can land one ulp away, which matters for one thing only: a partner that lies on the
// cutoff sphere to the last bit -- a whole neighbour shell of a perfect hcp lattice does (BASELINE configs[3] from
// config_ini.in: six neighbours at r = rcut exactly) -- is inside for one rounding and outside for the other.
__device__ __forceinline__ double r2_ref(double d0, double d1, double d2) {
    return __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
}
__device__ __forceinline__ double mimg_fast(double d, double L, double invL) {
#if defined(PIGS_MIMG_FRND)
    // FRND.F64 on the conversion pipe: 2 FP64 slots instead of 3 and no magic constant (-12 FP64 instructions per
    // partner, no LDC of 1/L per iteration) -- measured 1.3 % SLOWER in the sweep kernel (C3 467 -> 461 M), 2.5 %
    // slower in the isolated loop: FRND's latency is longer than the DADD it replaces.  Kept as a build option.
    return fma(-rint(d * invL), L, d);
#endif
    const double MAGIC = magic52();
    double q = fma(d, invL, MAGIC) - MAGIC;
    return fma(-q, L, d);
}

// the minimum image of the pair loops: component k (a literal at every call site)
__device__ __forceinline__ double mimg_hot(double d, int k, double invL) {
    return mimg_fast(d, cP.L[k], invL);
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
// two values: lanes 0..15 end with the sum of a, lanes 16..31 with the sum of b
__device__ __forceinline__ double warp_sum2(double a, double b, int lane) {
    const bool h16 = lane & 16;
    double v = (h16 ? b : a) + shx(h16 ? a : b, 16);
    v += shx(v, 8); v += shx(v, 4); v += shx(v, 2); v += shx(v, 1);
    return v;
}

// ------------------------------------------------------------------ RNG
// MT19937 with the 1998 seeding, bit-faithful to random_mod.f90:5-115.  State
// lives in global memory (624 words per chain); only thread 0 of a group draws.
__device__ inline void mt_seed(unsigned* mt, int& mti, unsigned seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 69069u * mt[i - 1];
    mti = 624;
}
__device__ __forceinline__ unsigned mt_temper(unsigned y);
static __device__ __noinline__ unsigned mt_next(unsigned* mt, int* pmti) {
    int mti = *pmti;
    if (mti >= 624) {
        if (mti == 625) mt_seed(mt, mti, 4357u);
        for (int kk = 0; kk < 624; ++kk) {
            unsigned y = (mt[kk] & 0x80000000u) | (mt[(kk + 1) % 624] & 0x7fffffffu);
            mt[kk] = mt[(kk + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        mti = 0;
    }
    unsigned y = mt[mti++];
    *pmti = mti;
    return mt_temper(y);
}
__device__ __forceinline__ double mt_grnd(GS* gs) {                 // [0,1] inclusive
    return (double)mt_next(gs->mt, &gs->mti) / 4294967295.0;
}
// one attempt of the polar method (random_mod.f90:195-219): two draws in, accepted iff w <= 1, first deviate out
__device__ __forceinline__ bool mt_polar(unsigned r1, unsigned r2, double& g) {
    const double u1 = 2.0 * ((double)r1 / 4294967295.0) - 1.0;
    const double u2 = 2.0 * ((double)r2 / 4294967295.0) - 1.0;
    const double w = u1 * u1 + u2 * u2;
    if (!(w <= 1.0)) return false;
    g = u1 * sqrt((-2.0 * log(w)) / w);
    return true;
}
static __device__ __noinline__ double mt_rangauss(GS* gs) {
    double g;
    unsigned r1, r2;
    do {
        r1 = mt_next(gs->mt, &gs->mti);
        r2 = mt_next(gs->mt, &gs->mti);
    } while (!mt_polar(r1, r2, g));
    return g;
}
__device__ __forceinline__ unsigned mt_temper(unsigned y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// Philox4x32-10 (Salmon et al., SC'11); counter = (ctr_lo, ctr_hi, chain, tag), key = seed
static __device__ __noinline__ uint4 philox_at(unsigned long long c, unsigned chain, unsigned tag) {
    uint4 x = make_uint4((unsigned)c, (unsigned)(c >> 32), chain, 0x50494753u + tag);
    unsigned k0 = (unsigned)cP.seed, k1 = (unsigned)(cP.seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(0xD2511F53u, x.x), lo0 = 0xD2511F53u * x.x;
        unsigned hi1 = __umulhi(0xCD9E8D57u, x.z), lo1 = 0xCD9E8D57u * x.z;
        x = make_uint4(hi1 ^ x.y ^ k0, lo1, hi0 ^ x.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return x;
}
__device__ __forceinline__ double u01_from(unsigned lo, unsigned hi) {        // [0,1)
    unsigned long long x = ((unsigned long long)hi << 32) | lo;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

// Per-thread view of the chain's Philox stream (group-uniform by construction):
// the 64-bit draw counter plus the unused 32-bit words of the last block, so one
// Philox4x32 call serves four uniforms (the reference's grnd() has 32 bits too).
// one uniform, identical in every thread of the group
template <bool MT>
__device__ __forceinline__ double rng_uniform(GS* gs, RngS& st) {
    if (MT) {
        gsync(gs);
        if (gfirst(gs)) gs->bc[0] = mt_grnd(gs);
        gsync(gs);
        return gs->bc[0];
    } else {
        unsigned x;
        if (st.nleft == 0) {
            uint4 r = philox_at(st.ctr, (unsigned)gs->chain, st.tag);
            st.ctr += 1;
            x = r.x; st.w0 = r.y; st.w1 = r.z; st.w2 = r.w; st.nleft = 3;
        } else {
            x = st.w0; st.w0 = st.w1; st.w1 = st.w2; st.nleft -= 1;
        }
        return (double)x * 2.3283064365386963e-10;        // x / 2^32 in [0,1)
    }
}
// unit Gaussians for beads b0 + j*bstride (j<nb), components k<dim, written into
// seg_new[k*S + bead] in the reference's draw order (bead-major, component-minor).
// Ends with a group sync.
template <bool MT>
__device__ __forceinline__ void rng_gauss_fill(GS* gs, RngS* pst, int dim, int b0, int bstride, int nb) {
    const Grp G = grp(gs);
    const int n = nb * dim;
    double* dst = seg_new(gs);
    if (MT) {
#ifdef PIGS_MT_SERIAL_GAUSS
        if (G.tid == 0) {
            for (int i = 0; i < n; ++i) {
                int j = i / dim, k = i - j * dim;
                dst[k * cP.S + b0 + j * bstride] = mt_rangauss(gs);
            }
        }
#else
        // The reference draws the n deviates one after the other; every attempt of the polar method consumes exactly
        // two words of the stream whether it is accepted or not, so attempt a reads words mti + 2a, mti + 2a + 1 of
        // the state block and the i-th deviate is the i-th ACCEPTED attempt: the lanes of the first warp run 32
        // attempts at once (tempering, log, sqrt in parallel), a ballot ranks the accepted ones, and the stream
        // position moves past the attempt that delivered the n-th deviate -- the same words consumed, the same
        // deviates, the same MT19937 state as the serial loop.  The attempt that triggers or straddles the refill
        // of the 624-word block goes through the serial code.
        int produced = 0;
        while (produced < n) {                                  // group-uniform
            const int mti = gs->mti;
            const int avail = (624 - mti) >> 1;                 // whole attempts left in this block (<= 0: refill first)
            if (avail <= 0) {
                if (G.tid == 0) {
                    const unsigned r1 = mt_next(gs->mt, &gs->mti), r2 = mt_next(gs->mt, &gs->mti);
                    double gv;
                    const bool ok = mt_polar(r1, r2, gv);
                    if (ok) { const int j = produced / dim, k = produced - j * dim; dst[k * cP.S + b0 + j * bstride] = gv; }
                    gs->bc[7] = ok ? 1.0 : 0.0;
                }
                gsync(gs);
                produced += (int)gs->bc[7];
                gsync(gs);
                continue;
            }
            if (G.warp == 0) {
                const int A = avail < 32 ? avail : 32;
                bool ok = false;
                double gv = 0.0;
                if (G.lane < A) ok = mt_polar(mt_temper(gs->mt[mti + 2 * G.lane]), mt_temper(gs->mt[mti + 2 * G.lane + 1]), gv);
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                const int need = n - produced, tot = __popc(m);
                const int take = tot < need ? tot : need;
                const int used = tot < need ? A : (int)__fns(m, 0, need) + 1;     // attempts consumed
                const int rank = __popc(m & ((1u << G.lane) - 1u));
                if (ok && rank < take) {
                    const int i = produced + rank, j = i / dim, k = i - j * dim;
                    dst[k * cP.S + b0 + j * bstride] = gv;
                }
                if (G.lane == 0) { gs->mti = mti + 2 * used; gs->bc[7] = (double)take; }
            }
            gsync(gs);
            produced += (int)gs->bc[7];
            gsync(gs);
        }
#endif
    } else {
        unsigned long long ctr = pst->ctr;
        for (int i = G.tid; i < n; i += G.size) {
            int j = dim == 3 ? (i * 43691) >> 17 : (dim == 2 ? i >> 1 : i), k = i - j * dim;
            uint4 r = philox_at(ctr + (unsigned long long)i, (unsigned)gs->chain, pst->tag);
            double u1 = 1.0 - u01_from(r.x, r.y);       // (0,1]
            double u2 = u01_from(r.z, r.w);
            dst[k * cP.S + b0 + j * bstride] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        }
        pst->ctr = ctr + (unsigned long long)n;
    }
    gsync(gs);
}

// ------------------------------------------------------------------ the pair sums of one bead-update
// Per-lane partial sums for one displaced bead (UpdatePot, vpi_mod.f90:2660-2841,
// UpdateWf, :2534-2656), old and new position against partner j.
// KIND: 0 interior even slice   -> pot
//       1 odd slice (Chin term) -> pot, Fnew(3), Fold(3)
//       2 end slice (Jastrow)   -> pot, psi
// (One loop serves the three classes, see pair_loop; the kernel runs at 128 registers per thread.)
//
// Branch-free: a partner outside the cutoff (or the moved particle itself) is
// evaluated at r = rcut with weight 0, so the new and old positions form
// independent dependency chains the scheduler can overlap.
struct PairGeom {
    double d0, d1, d2, ir;
    Lk k;
    bool in_pot, in_wf;
};
template <int KIND, bool TRAP, bool IS_NEW, bool XR = false>
__device__ __forceinline__ PairGeom pair_geom(bool valid, double x0, double x1, double x2, double rx, double ry, double rz) {
    PairGeom g;
    g.d0 = x0 - rx; g.d1 = x1 - ry; g.d2 = x2 - rz;
    if (!TRAP) {
        if (XR) { g.d0 = mimg_cmp(g.d0, cP.L[0], cP.Lh[0]); g.d1 = mimg_cmp(g.d1, cP.L[1], cP.Lh[1]); g.d2 = mimg_cmp(g.d2, cP.L[2], cP.Lh[2]); }
        else { g.d0 = mimg_hot(g.d0, 0, cP.invL[0]); g.d1 = mimg_hot(g.d1, 1, cP.invL[1]); g.d2 = mimg_hot(g.d2, 2, cP.invL[2]); }
    }
    double r2 = XR ? r2_ref(g.d0, g.d1, g.d2) : g.d0 * g.d0 + g.d1 * g.d1 + g.d2 * g.d2;
    // PBC: both positions cut at rcut (Q24).  Trap: UpdatePot cuts only the OLD
    // position (Q12), UpdateWf cuts nothing.
    g.in_pot = valid && (TRAP ? (IS_NEW || r2 <= cP.rcut2) : (r2 <= cP.rcut2));
    g.in_wf = TRAP ? valid : g.in_pot;
    const bool any = (KIND == 2) ? g.in_wf : g.in_pot;
    double r2c = any ? r2 : (TRAP ? cP.rcut2 : cP.rclamp2);
    if (KIND == 1) {           // the force needs 1/r as well
        g.ir = rsqrt_pos(r2c);
        g.k = lk_prep(r2c * g.ir);
    } else {
        g.ir = 0.0;
        g.k = lk_prep(sqrt_pos(r2c));
    }
    if (TRAP) g.k.i0 = min(g.k.i0, cP.Nmax - 1);
    return g;
}

// 0 interior slice without force term, 1 odd slice (Chin force term), 2 end slice (Jastrow)
__device__ __forceinline__ int bead_kind(int ib) { return (ib == 0 || ib == 2 * cP.Nb) ? 2 : (cP.primitive ? 0 : (ib & 1)); }

// coherent global load of a path coordinate (the path is written by this group
// during the kernel, so no .nc; an explicit ld.global avoids the generic-address
// resolution of a plain pointer dereference)
__device__ __forceinline__ double ldpath(const double* p) {
    double v;
    // streamed once per bead-update: keeping it out of L1 (24 KB left beside the tables) is worth +3.6 % on C3;
    // an L2 evict-first policy on top changes nothing
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ldpath_ca(const double* p) {
    double v;
    asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
struct Partner {
    double x, y, z;
};
__device__ __forceinline__ Partner load_partner(const double* Rx, int j) {
    Partner p;
    const double* q = Rx + pidx(j);
    p.x = ldpath(q); p.y = ldpath(q + PY); p.z = ldpath(q + PZ);
    return p;
}

// both positions of the displaced bead against ONE partner (round-1 form; pair_body2 below is the production one)
template <bool TRAP, bool VSM, bool WSM, bool VPAIR, bool XR = false>
__device__ __forceinline__ void pair_body(int kind, bool valid, const Partner& cur, const double (&xo)[3],
                                          const double (&xn)[3], double& pot, double& psi, double (&fn)[3], double (&fo)[3]) {
    constexpr bool MASK = TRAP || VPAIR;        // else: masked pairs read the zero tail of the tables
    if (kind == 1) {
        {
            PairGeom g = pair_geom<1, TRAP, true, XR>(valid, xn[0], xn[1], xn[2], cur.x, cur.y, cur.z);
            double v, dv;
            if (VPAIR) lk_val_d1_pair(g.k, v, dv); else lk_val_d1<VSM, 0, VSM>(g.k, v, dv);
            pot += (!MASK || g.in_pot) ? v : 0.0;
            double s = (!MASK || g.in_pot) ? dv * g.ir : 0.0;
            fn[0] += s * g.d0; fn[1] += s * g.d1; fn[2] += s * g.d2;
        }
        {
            PairGeom g = pair_geom<1, TRAP, false, XR>(valid, xo[0], xo[1], xo[2], cur.x, cur.y, cur.z);
            double v, dv;
            if (VPAIR) lk_val_d1_pair(g.k, v, dv); else lk_val_d1<VSM, 0, VSM>(g.k, v, dv);
            pot -= (!MASK || g.in_pot) ? v : 0.0;
            double s = (!MASK || g.in_pot) ? dv * g.ir : 0.0;
            fo[0] += s * g.d0; fo[1] += s * g.d1; fo[2] += s * g.d2;
        }
    } else {
        // even and end slices share the geometry (in trap mode the end slice has no Jastrow cutoff)
        PairGeom gn = (TRAP && kind == 2) ? pair_geom<2, TRAP, true, XR>(valid, xn[0], xn[1], xn[2], cur.x, cur.y, cur.z)
                                          : pair_geom<0, TRAP, true, XR>(valid, xn[0], xn[1], xn[2], cur.x, cur.y, cur.z);
        PairGeom go = (TRAP && kind == 2) ? pair_geom<2, TRAP, false, XR>(valid, xo[0], xo[1], xo[2], cur.x, cur.y, cur.z)
                                          : pair_geom<0, TRAP, false, XR>(valid, xo[0], xo[1], xo[2], cur.x, cur.y, cur.z);
        double vn = VPAIR ? lk_val_pair(gn.k) : lk_val<VSM, 0, VSM>(gn.k);
        double vo = VPAIR ? lk_val_pair(go.k) : lk_val<VSM, 0, VSM>(go.k);
        pot += ((!MASK || gn.in_pot) ? vn : 0.0) - ((!MASK || go.in_pot) ? vo : 0.0);
        if (kind == 2) {
            double wn = lk_val<WSM, 1, VSM && !VPAIR>(gn.k), wo = lk_val<WSM, 1, VSM && !VPAIR>(go.k);
            psi += ((!MASK || gn.in_wf) ? wn : 0.0) - ((!MASK || go.in_wf) ? wo : 0.0);
        }
    }
}

// ONE partner loop for the three slice classes (kind is warp-uniform, so the
// class-specific parts are skipped by uniform branches): the hot code of all
// warps of a scheduler is the same few KB and stays in the L0 instruction cache
// (three specialised loops were 6.6 KB; ncu showed 1.3 no_instruction stall
// cycles per issue).
// acc: pot, psi, fn[3], fo[3].  `A` = coordinates of partner j0, preloaded by the
// caller (during the previous bead's reduction / the proposal).
// Partner coordinates are software-pipelined one iteration ahead.  (Measured on
// B200: a two-deep pipeline, an A/B ping-pong body and a 2x unrolled body were
// all 5-25% slower -- larger loop bodies lose more than the extra latency
// tolerance gains.)
template <bool TRAP, bool VSM, bool WSM, bool VPAIR, bool XR = false>
__device__ __forceinline__ void pair_loop(int kind, const double* Rx, int ip0, int j0, int jstride, const double (&xo)[3],
                                          const double (&xn)[3], Partner nxt, double& pot, double& psi, double (&fn)[3],
                                          double (&fo)[3]) {
    // running pointer + countdown: no per-iteration address arithmetic, no loop-invariant reloads
    // (j0 = 32*s + lane and jstride = 32*split: the warp walks whole [3][32] blocks)
    const double* p = Rx + pidx(j0);
    const int pstep = 3 * jstride;
    const int self_left = cP.Np - ip0;
PIGS_PRAGMA_UNROLL
    for (int left = cP.Np - j0; left > 0; left -= jstride) {
        const Partner cur = nxt;
        p += pstep;
        if (left > jstride) { nxt.x = ldpath(p); nxt.y = ldpath(p + PY); nxt.z = ldpath(p + PZ); }
        pair_body<TRAP, VSM, WSM, VPAIR, XR>(kind, left != self_left, cur, xo, xn, pot, psi, fn, fo);
    }
}


// ------------------------------------------------------------------ partner loop, second generation
// Compile-time selection of the loop variants (bit mask; measured in isolation with scripts/loopbench.cu on B200,
// N = 256, mixed slice classes, 16 warps/SM, HBM-resident / L2-resident slices, M bead-updates/s):
//   0   round-1 loop (cur/nxt register copy, per-position self mask)                      623 / 655
//   1   partner registers reloaded IN PLACE right after their last use (no cur/nxt copy: -9 MOV, -6 registers);
//       the moved particle excludes itself through a poisoned x coordinate (r^2 ~ 1e268 -> zero tail of the
//       tables) instead of two selects per position                                         640 / 684
//   32  + the cutoff acts on the table INDEX, off the critical path (sqrt of the unclamped r^2)     648 / 695   <- on
//   16  + one Newton step instead of the three-term step (1e-12 in r): +0.7 %, not worth the digits   652 / 703
//   4   wrap count of the minimum image from a compare on the high word of d: -12 FP64 slots, +10 ALU   635 / 678 (slower)
//   8   table reads through explicit ld.shared with a 32-bit base: no change
//   64  L1 prefetch two blocks ahead + L1-allocating loads: 593 / 639 (slower);  128 L1-allocating loads only: 648 / 706
//   256 two partner blocks per iteration (four dependency chains per warp), partner registers carried from bead to bead:
//       isolated loop 515 -> 663 (HBM, no prefetch) / 679 -> 696 (L2) / N=64 1564 -> 1758, but INSIDE the sweep kernel, at
//       its 128-register cap, ptxas serialises the four chains: C3 457 -> 425 M, C2 894 -> 755 M.  Off.
//   512 in-range compaction (pair_loop3: cheap scan of every partner, the 52 % inside the cutoff sphere go through a
//       per-warp shared-memory ring and get the square root / table / force part in dense batches of 32): FP64 slots
//       per bead-update -25 %, yet 716 -> 559 (L2), 528 -> 466 (HBM), N=64 1633 -> 1306: scan -> ballot -> ring ->
//       dense is one serial chain per block, and the loop is latency-bound, not slot-bound.  Off.
//   1024 one block per iteration, partner registers CARRIED from bead to bead: after their last use in a bead they
//       are reloaded with the first block of the next evaluated slice (no preload in the bead loop, no copy): in the
//       sweep kernel C3 455 -> 461 M, C2 898 -> 914 M                                                    <- on
// Default 1057 = 1 + 32 + 1024.  Variants 1, 32 and 1024 give results bit-identical to variant 0.
#ifndef PIGS_LOOPV
#define PIGS_LOOPV 1057
#endif
#ifndef PIGS_KC
#define PIGS_KC 1
#endif
#define PIGS_EXPL_LDS (PIGS_LOOPV & 8)

// d - L*q with q = -1, 0, +1 decided on the HIGH WORD of d: |d| > L/2 is judged with a resolution of 2^-20
// relative.  A component inside the band (L/2, L/2 (1 + 2^-20)] keeps the far image; such a pair lies beyond the
// cutoff unless its transverse distance is below ~1e-3 sigma, and then both images sit at the cutoff radius:
// probability ~1e-12 per pair evaluation, effect |V(rcut)| dt ~ 1e-5 on DeltaS (documented in DESIGN.md section 5).
__device__ __forceinline__ int mimg_one_hi() {
    int v = 0x3ff00000;
    asm("" : "+r"(v));          // opaque: stays in a register instead of becoming a second immediate
    return v;
}
__device__ __forceinline__ double mimg_hi(double d, double L, float LhF) {
    const int hi = __double2hiint(d);
    int one;                                                   // (hi & sign) | 1.0 as ONE LOP3: the second constant in a register
    asm("lop3.b32 %0, %1, 0x80000000, %2, 0xEA;" : "=r"(one) : "r"(hi), "r"(mimg_one_hi()));
    const int qhi = (fabsf(__int_as_float(hi)) > LhF) ? one : 0;
    return fma(-__hiloint2double(qhi, 0), L, d);
}
// The four constants at the head of every partner's dependency chain, as REGISTER values.  Read from the constant
// bank they are re-fetched with an LDC in every loop iteration (ptxas rematerialises them under register
// pressure); read once per evaluation from the group's shared-memory block they cannot be rematerialised: -10 LDC,
// C3 +0.6 %, C2 +0.7 %, C5 +0.9 %.  Going further was measured and LOSES: L[0..2] and rcut^2 as well (-6 LDCU) -2.9 %,
// plus the tables' shared addresses and the zero-tail index (-8 uniform-datapath instructions, loop 242 -> 218
// instructions) -3.6 %: the registers they occupy cost more than the instructions they save
// (profiles/bench_r02_loop_constants_ab.log).
struct LoopK {
    double iL0, iL1, iL2, idr;
};
struct Pos2 {          // geometry of one position against one partner
    double d0, d1, d2, ir;
    Lk k;
};
// One Newton step on the hardware seed (MUFU.RSQ64H, >= 20 bits): relative error <= 1.5 * 2^-40 ~ 1.4e-12 in r and
// 1/r -- three dependent FP64 latencies instead of four, one slot less.  (The three-term step of rsqrt_pos /
// sqrt_pos gives 2^-52; 1e-12 in r moves DeltaS by < 1e-13, inside the 1e-10 parity budget.)
__device__ __forceinline__ double sqrt_q(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double r0 = x * y;
    const double h = 0.5 * r0;
    const double e = fma(-r0, y, 1.0);
    return fma(h, e, r0);
}
__device__ __forceinline__ void rsqrt_sqrt_q(double x, double& ir, double& r) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double r0 = x * y;
    const double e = fma(-r0, y, 1.0);
    ir = fma(0.5 * y, e, y);
    r = fma(0.5 * r0, e, r0);
}
template <bool NEED_IR, bool WRAPPED = false, bool XR = false>
__device__ __forceinline__ Pos2 pos_geom(double d0, double d1, double d2, const LoopK* K = nullptr) {
    Pos2 g;
    const double inv_dr = K ? K->idr : cP.inv_dr;
    if (WRAPPED) {                   // the components are minimum-image components already
        g.d0 = d0; g.d1 = d1; g.d2 = d2;
    } else if (XR) {
        g.d0 = mimg_cmp(d0, cP.L[0], cP.Lh[0]); g.d1 = mimg_cmp(d1, cP.L[1], cP.Lh[1]); g.d2 = mimg_cmp(d2, cP.L[2], cP.Lh[2]);
    } else if (PIGS_LOOPV & 4) {
        g.d0 = mimg_hi(d0, cP.L[0], cP.LhF[0]); g.d1 = mimg_hi(d1, cP.L[1], cP.LhF[1]); g.d2 = mimg_hi(d2, cP.L[2], cP.LhF[2]);
    } else {
        g.d0 = mimg_hot(d0, 0, K ? K->iL0 : cP.invL[0]); g.d1 = mimg_hot(d1, 1, K ? K->iL1 : cP.invL[1]);
        g.d2 = mimg_hot(d2, 2, K ? K->iL2 : cP.invL[2]);
    }
#ifdef PIGS_PHILOX_R2REF      /* build option: the reference's rij2 in the Philox kernel too (measured -1.9 % at C3, -0.9 % at C2) */
    const double r2 = r2_ref(g.d0, g.d1, g.d2);
#else
    const double r2 = XR ? r2_ref(g.d0, g.d1, g.d2) : g.d0 * g.d0 + g.d1 * g.d1 + g.d2 * g.d2;
#endif
    if (PIGS_LOOPV & 32) {
        // the cutoff acts on the table INDEX, off the critical path: sqrt of the unclamped r^2 (finite: the self
        // partner is poisoned with 1e150, not infinity), then i0 = zero tail unless r^2 <= rcut^2 (Q24)
        const bool in = r2 <= cP.rcut2;
        double r;
        if (NEED_IR) { if (PIGS_LOOPV & 16) rsqrt_sqrt_q(r2, g.ir, r); else { g.ir = rsqrt_pos(r2); r = r2 * g.ir; } }
        else { g.ir = 0.0; r = (PIGS_LOOPV & 16) ? sqrt_q(r2) : sqrt_pos(r2); }
        const double MAGIC = magic52();
        const double m = __fma_rd(r, inv_dr, MAGIC);
        g.k.i0 = in ? max(__double2loint(m), 1) : cP.Nmax + 3;
        g.k.t = fma(r, inv_dr, -(m - MAGIC));
        return g;
    }
    const double r2c = (r2 <= cP.rcut2) ? r2 : cP.rclamp2;        // beyond the cutoff, poisoned or NaN: zero tail (Q24)
    if (NEED_IR) {
        if (PIGS_LOOPV & 16) { double r; rsqrt_sqrt_q(r2c, g.ir, r); g.k = lk_prep(r); }
        else { g.ir = rsqrt_pos(r2c); g.k = lk_prep(r2c * g.ir); }
    } else {
        g.ir = 0.0;
        g.k = lk_prep((PIGS_LOOPV & 16) ? sqrt_q(r2c) : sqrt_pos(r2c));
    }
    return g;
}
template <bool SM, int WHICH, bool VF>
__device__ __forceinline__ double lk2_val(const Lk& k, unsigned sb) {
    if (SM && PIGS_EXPL_LDS) {
        const unsigned a = sb + ((unsigned)k.i0 << 3);
        double f0, f1;
        asm("ld.shared.f64 %0, [%1];" : "=d"(f0) : "r"(a));
        asm("ld.shared.f64 %0, [%1+8];" : "=d"(f1) : "r"(a));
        return fma(k.t, f1 - f0, f0);
    }
    return lk_val<SM, WHICH, VF>(k);
}
template <bool SM, int WHICH, bool VF>
__device__ __forceinline__ void lk2_val_d1(const Lk& k, unsigned sb, double& v, double& d1) {
    if (SM && PIGS_EXPL_LDS) {
        const unsigned a = sb + ((unsigned)k.i0 << 3);
        double fm, f0, f1, f2;
        asm("ld.shared.f64 %0, [%1+-8];" : "=d"(fm) : "r"(a));
        asm("ld.shared.f64 %0, [%1];" : "=d"(f0) : "r"(a));
        asm("ld.shared.f64 %0, [%1+8];" : "=d"(f1) : "r"(a));
        asm("ld.shared.f64 %0, [%1+16];" : "=d"(f2) : "r"(a));
        lk_d1_from(k.t, fm, f0, f1, f2, v, d1);
        return;
    }
    lk_val_d1<SM, WHICH, VF>(k, v, d1);
}
// both positions of the displaced bead against ONE partner; dn/dq = x_new - r_j, x_old - r_j before the minimum image
template <bool VSM, bool WSM, bool WRAPPED = false, bool XR = false>
__device__ __forceinline__ void pair_body2(int kind, unsigned sbV, unsigned sbW, double dn0, double dn1, double dn2, double dq0,
                                           double dq1, double dq2, double& pot, double& psi, double (&fn)[3], double (&fo)[3],
                                           const LoopK* K = nullptr) {
    if (kind == 1) {
        {
            const Pos2 g = pos_geom<true, WRAPPED, XR>(dn0, dn1, dn2, K);
            double v, dv;
            lk2_val_d1<VSM, 0, VSM>(g.k, sbV, v, dv);
            pot += v;
            const double s = dv * g.ir;
            fn[0] += s * g.d0; fn[1] += s * g.d1; fn[2] += s * g.d2;
        }
        {
            const Pos2 g = pos_geom<true, WRAPPED, XR>(dq0, dq1, dq2, K);
            double v, dv;
            lk2_val_d1<VSM, 0, VSM>(g.k, sbV, v, dv);
            pot -= v;
            const double s = dv * g.ir;
            fo[0] += s * g.d0; fo[1] += s * g.d1; fo[2] += s * g.d2;
        }
    } else {
        const Pos2 gn = pos_geom<false, WRAPPED, XR>(dn0, dn1, dn2, K);
        const Pos2 go = pos_geom<false, WRAPPED, XR>(dq0, dq1, dq2, K);
        pot += lk2_val<VSM, 0, VSM>(gn.k, sbV) - lk2_val<VSM, 0, VSM>(go.k, sbV);
        if (kind == 2) psi += lk2_val<WSM, 1, VSM>(gn.k, sbW) - lk2_val<WSM, 1, VSM>(go.k, sbW);
    }
}
// Partner loop with in-range compaction (PIGS_LOOPV & 512; whole partner range in one warp).  At rcut = L/2 only
// pi/6 = 52 % of the partners lie inside the cutoff sphere, yet a lane-per-partner loop pays the square root, the
// table gathers and the force for all of them.  Here a cheap SCAN (minimum image + r^2 of both positions, 30 FP64
// slots per partner) pushes the pairs that are in range for either position onto a per-warp ring of 64 entries in
// shared memory (q[6][64]: the six minimum-image components), and the expensive part runs on DENSE batches of 32
// ring entries.  The moved particle excludes itself by never entering the ring.
template <bool VSM, bool WSM>
__device__ __forceinline__ void pair_loop3(int kind, const double* Rx, int ip0, int lane, const double (&xo)[3],
                                           const double (&xn)[3], Partner cur, double* q, double& pot, double& psi,
                                           double (&fn)[3], double (&fo)[3]) {
    const double* p = Rx + pidx(lane);
    const int nblk = (cP.Np + 31) >> 5;
    const unsigned lt = (1u << lane) - 1u;
    int qn = 0, qh = 0;
    for (int blk = 0, j = lane; blk < nblk; ++blk, j += 32) {
        const double a0 = mimg_hot(xn[0] - cur.x, 0, cP.invL[0]), a1 = mimg_hot(xn[1] - cur.y, 1, cP.invL[1]),
                     a2 = mimg_hot(xn[2] - cur.z, 2, cP.invL[2]);
        const double b0 = mimg_hot(xo[0] - cur.x, 0, cP.invL[0]), b1 = mimg_hot(xo[1] - cur.y, 1, cP.invL[1]),
                     b2 = mimg_hot(xo[2] - cur.z, 2, cP.invL[2]);
        p += 96;
        if (blk + 1 < nblk) { cur.x = ldpath(p); cur.y = ldpath(p + PY); cur.z = ldpath(p + PZ); }      // in place, one block ahead
        const double r2n = a0 * a0 + a1 * a1 + a2 * a2, r2o = b0 * b0 + b1 * b1 + b2 * b2;
        const bool in = (j < cP.Np) && (j != ip0) && (fmin(r2n, r2o) <= cP.rcut2);
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (in) {
            double* e = q + ((qh + qn + __popc(m & lt)) & 63);
            e[0] = a0; e[64] = a1; e[128] = a2; e[192] = b0; e[256] = b1; e[320] = b2;
        }
        qn += __popc(m);
        if (qn >= 32) {
            __syncwarp();
            const double* e = q + ((qh + lane) & 63);
            const double c0 = e[0], c1 = e[64], c2 = e[128], d0 = e[192], d1 = e[256], d2 = e[320];
            __syncwarp();
            qh = (qh + 32) & 63; qn -= 32;
            pair_body2<VSM, WSM, true>(kind, 0u, 0u, c0, c1, c2, d0, d1, d2, pot, psi, fn, fo);
        }
    }
    if (qn > 0) {
        __syncwarp();
        const double* e = q + ((qh + lane) & 63);
        const bool act = lane < qn;
        const double c0 = act ? e[0] : 1e150, c1 = act ? e[64] : 0.0, c2 = act ? e[128] : 0.0;
        const double d0 = act ? e[192] : 1e150, d1 = act ? e[256] : 0.0, d2 = act ? e[320] : 0.0;
        __syncwarp();
        pair_body2<VSM, WSM, true>(kind, 0u, 0u, c0, c1, c2, d0, d1, d2, pot, psi, fn, fo);
    }
}
// partner registers carried from bead to bead by the two-block loop (PIGS_LOOPV & 256): blocks j0 and j0 + jstride of
// the slice about to be evaluated, and the slice after it (nullptr: none)
struct Carry {
    Partner a, b;
    const double* next;
    const LoopK* K;
};
template <bool VSM, bool WSM, bool XR = false>
__device__ __forceinline__ void pair_loop2(int kind, const double* Rx, int ip0, int j0, int jstride, const double (&xo)[3],
                                           const double (&xn)[3], Partner cur, double& pot, double& psi, double (&fn)[3],
                                           double (&fo)[3], Carry* cy = nullptr) {
    const double* p = Rx + pidx(j0);
    const int pstep = 3 * jstride;
    const int self_left = cP.Np - ip0;
    unsigned sbV = 0, sbW = 0;
    if (PIGS_EXPL_LDS) {
        extern __shared__ __align__(16) double pigs_smem_base[];
        sbV = (unsigned)__cvta_generic_to_shared(pigs_smem_base);
        sbW = sbV + (unsigned)cP.tabW_off;
    }
    if ((PIGS_LOOPV & 1024) && cy) {
        // one block per iteration, partner registers carried from bead to bead: after its last use in this bead the
        // register triple is reloaded with the first block of the next evaluated slice (no separate preload, no copy)
        Partner a = cy->a;
        const double* pn = cy->next ? cy->next + pidx(j0) : nullptr;
        for (int left = cP.Np - j0; left > 0; left -= jstride) {
            if (left == self_left) a.x = 1e150;
            const double dn0 = xn[0] - a.x, dn1 = xn[1] - a.y, dn2 = xn[2] - a.z;
            const double dq0 = xo[0] - a.x, dq1 = xo[1] - a.y, dq2 = xo[2] - a.z;
            p += pstep;
            if (left > jstride) { a.x = ldpath(p); a.y = ldpath(p + PY); a.z = ldpath(p + PZ); }
            else if (pn) { a.x = ldpath(pn); a.y = ldpath(pn + PY); a.z = ldpath(pn + PZ); }
            pair_body2<VSM, WSM, false, XR>(kind, sbV, sbW, dn0, dn1, dn2, dq0, dq1, dq2, pot, psi, fn, fo, cy->K);
        }
        cy->a = a;
        return;
    }
    if ((PIGS_LOOPV & 256) && cy) {
        // Two partner blocks per iteration: four independent dependency chains per warp instead of two.  The partner
        // registers are a stream that runs across beads: after their last use in a bead they are reloaded with the
        // first two blocks of the NEXT evaluated slice (cy->next), so no load is ever exposed at a loop entry.
        // Same accumulators, same order of additions as the one-block loop: bit-identical results.
        Partner a = cy->a, b = cy->b;
        const double* pn = cy->next ? cy->next + pidx(j0) : nullptr;
        const int left0 = cP.Np - j0;
        for (int left = left0; left > 0; left -= 2 * jstride) {
            if (left <= jstride) b.x = 1e150;                   // no second block: zero tail
            if (left == self_left) a.x = 1e150;
            if (left - jstride == self_left) b.x = 1e150;
            const double an0 = xn[0] - a.x, an1 = xn[1] - a.y, an2 = xn[2] - a.z;
            const double aq0 = xo[0] - a.x, aq1 = xo[1] - a.y, aq2 = xo[2] - a.z;
            const double bn0 = xn[0] - b.x, bn1 = xn[1] - b.y, bn2 = xn[2] - b.z;
            const double bq0 = xo[0] - b.x, bq1 = xo[1] - b.y, bq2 = xo[2] - b.z;
            p += 2 * pstep;
            if (left > 2 * jstride) {
                a.x = ldpath(p); a.y = ldpath(p + PY); a.z = ldpath(p + PZ);
                if (left > 3 * jstride) { b.x = ldpath(p + pstep); b.y = ldpath(p + pstep + PY); b.z = ldpath(p + pstep + PZ); }
            } else if (pn) {
                a.x = ldpath(pn); a.y = ldpath(pn + PY); a.z = ldpath(pn + PZ);
                if (left0 > jstride) { b.x = ldpath(pn + pstep); b.y = ldpath(pn + pstep + PY); b.z = ldpath(pn + pstep + PZ); }
            }
            pair_body2<VSM, WSM, false, XR>(kind, sbV, sbW, an0, an1, an2, aq0, aq1, aq2, pot, psi, fn, fo);
            pair_body2<VSM, WSM, false, XR>(kind, sbV, sbW, bn0, bn1, bn2, bq0, bq1, bq2, pot, psi, fn, fo);
        }
        cy->a = a; cy->b = b;
        return;
    }
PIGS_PRAGMA_UNROLL
    for (int left = cP.Np - j0; left > 0; left -= jstride) {
        if (left == self_left) cur.x = 1e150;                   // the moved particle itself: r^2 ~ 1e268+ -> zero tail
        const double dn0 = xn[0] - cur.x, dn1 = xn[1] - cur.y, dn2 = xn[2] - cur.z;
        const double dq0 = xo[0] - cur.x, dq1 = xo[1] - cur.y, dq2 = xo[2] - cur.z;
        p += pstep;
        if (PIGS_LOOPV & 64) {       // L1 prefetch two blocks ahead (6 lines of 128 B per [3][32] block), L1-allocating loads
            if (left > 2 * jstride && (threadIdx.x & 31) < 6)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const char*>(p + pstep - ((threadIdx.x & 31) )) + ((threadIdx.x & 31) << 7)));
        }
        if (left > jstride) {
            if (PIGS_LOOPV & (64 | 128)) { cur.x = ldpath_ca(p); cur.y = ldpath_ca(p + PY); cur.z = ldpath_ca(p + PZ); }
            else { cur.x = ldpath(p); cur.y = ldpath(p + PY); cur.z = ldpath(p + PZ); }      // in place, one iteration ahead
        }
        pair_body2<VSM, WSM, false, XR>(kind, sbV, sbW, dn0, dn1, dn2, dq0, dq1, dq2, pot, psi, fn, fo);
    }
}

// DeltaS of UpdateAction from the eight reduced values [pot, psi, Fnew(3), Fold(3)]
// (GreenFunction opt 0, global_mod.f90:29-46).
__device__ __forceinline__ double assemble_dS(int ib, const double (&v)[8]) {
    int kind = bead_kind(ib);
    asm volatile("" : "+r"(kind));      // opaque: keeps the class in a register instead of re-deriving it from ib every iteration
    if (kind == 2) return -v[1] + cP.wS[2] * v[0];
    if (kind == 0) return cP.wS[ib & 1] * v[0];
    double f2 = (v[2] * v[2] + v[3] * v[3] + v[4] * v[4]) - (v[5] * v[5] + v[6] * v[6] + v[7] * v[7]);
    return cP.wS[1] * v[0] + cP.cF * f2;
}

// One bead against partners j0, j0+jstride, ... by one warp.
//  part == nullptr: all partners are covered by this warp -> returns DeltaS
//                   (UpdateAction, vpi_mod.f90:2491-2530), identical in every lane;
//  part != nullptr: the warp's partial sums are stored to part[0..7] for a later
//                   combination with other warps (returns 0).
// The lane with add_self adds the one-body (trap) terms once.  `first` holds the
// coordinates of partner j0 (preloaded by the caller; unused lanes pass anything).
template <bool TRAP, bool VSM, bool WSM, bool VPAIR, bool XR = false>
__device__ __forceinline__ double bead_eval(const double* Rx, int ip0, int ib, int j0, int jstride, bool add_self,
                                            const double (&xo)[3], const double (&xn)[3], int lane, double* part,
                                            const Partner& first, double* lin = nullptr, Carry* cy = nullptr,
                                            double* ring = nullptr) {
    // lin != nullptr (whole partner range in this warp): the part of DeltaS that is LINEAR in the per-lane sums -- the
    // potential and Jastrow terms -- is handed back unreduced in *lin (the caller reduces once per evaluation instead
    // of once per bead); only the Chin force term, quadratic in the reduced force, is reduced here and returned.
    int kind = bead_kind(ib);
    asm volatile("" : "+r"(kind));      // opaque: keeps the class in a register instead of re-deriving it from ib every iteration
    double pot = 0.0, psi = 0.0, fn[3] = {0.0, 0.0, 0.0}, fo[3] = {0.0, 0.0, 0.0};
    if (TRAP && add_self) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {          // system_mod.f90:213-252
            if (k < cP.dim) {
                double ak = cP.a_ho[k], a2 = ak * ak, a4 = a2 * a2;
                pot += 0.5 * xn[k] * xn[k] / a4 - 0.5 * xo[k] * xo[k] / a4;
                if (kind == 1) { fn[k] += xn[k] / a4; fo[k] += xo[k] / a4; }
                if (kind == 2) psi += -0.5 * (xn[k] / ak) * (xn[k] / ak) + 0.5 * (xo[k] / ak) * (xo[k] / ak);
            }
        }
    }
    if ((PIGS_LOOPV & 512) && !TRAP && !VPAIR && ring) pair_loop3<VSM, WSM>(kind, Rx, ip0, lane, xo, xn, first, ring, pot, psi, fn, fo);
    else if (PIGS_LOOPV != 0 && !TRAP && !VPAIR) pair_loop2<VSM, WSM, XR>(kind, Rx, ip0, j0, jstride, xo, xn, first, pot, psi, fn, fo, cy);
    else pair_loop<TRAP, VSM, WSM, VPAIR, XR>(kind, Rx, ip0, j0, jstride, xo, xn, first, pot, psi, fn, fo);
    if (lin) {
        if (kind == 0) { *lin = cP.wS[ib & 1] * pot; return 0.0; }
        if (kind == 2) { *lin = fma(cP.wS[2], pot, -psi); return 0.0; }
        const double a6[8] = {0.0, 0.0, fn[0], fn[1], fn[2], fo[0], fo[1], fo[2]};
        const double v6 = warp_sum8(a6, lane);          // quad q holds the complete sum of value q
        const int q6 = lane >> 2;
        const double c6 = cP.cF;
        // +-cF F^2 is quadratic in the reduced force, but the SUM of the six squares over the quads is linear again:
        // one lane per quad adds its square to the deferred part (three shuffle levels less per odd bead)
        const double t6 = (q6 < 2 || (lane & 3) != 0) ? 0.0 : ((q6 < 5) ? c6 * v6 * v6 : -c6 * v6 * v6);
        *lin = fma(cP.wS[1], pot, t6);
        return 0.0;
    }
    if (kind == 0) {
        double v = warp_sum(pot);
        if (!part) return cP.wS[ib & 1] * v;
        if (lane < 8) part[lane] = (lane == 0) ? v : 0.0;
        return 0.0;
    }
    if (kind == 2) {
        double v = warp_sum2(pot, psi, lane);          // lanes 0..15: pot, lanes 16..31: psi
        if (!part) {
            double t = (lane & 16) ? -v : cP.wS[2] * v;
            return t + shx(t, 16);
        }
        if (lane == 0) part[0] = v;
        if (lane == 16) part[1] = v;
        if (lane >= 2 && lane < 8) part[lane] = 0.0;
        return 0.0;
    }
    const double a[8] = {pot, 0.0, fn[0], fn[1], fn[2], fo[0], fo[1], fo[2]};
    double v = warp_sum8(a, lane);                      // quad q holds value q
    const int q = lane >> 2;
    if (!part) {
        const double c = cP.cF;
        double t = (q == 0) ? cP.wS[1] * v : ((q == 1) ? 0.0 : ((q < 5) ? c * v * v : -c * v * v));
        t += shx(t, 4); t += shx(t, 8); t += shx(t, 16);
        return t;
    }
    if ((lane & 3) == 0) part[q] = v;
    return 0.0;
}

// Bulk (TMA-engine) prefetch into L2 of whole time slices: one instruction per
// slice.  With thousands of resident chains the paths live in HBM; ncu showed
// the partner loop stalled on HBM latency that a one-iteration register
// pipeline cannot cover, while L2 latency it can.
__device__ __forceinline__ void prefetch_slice_L2(const double* slice_ptr) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(slice_ptr), "r"(3 * cP.NpS * 8) : "memory");
}
// L1 prefetch of the slices a move is going to read (other particles of the
// displaced beads' time slices): issued at move start so the HBM/L2 latency
// overlaps the proposal generation (ncu: 31% of stall samples sat on the first
// use of these loads).
__device__ __forceinline__ void prefetch_slices(const double* path, int b0, int nslice, int tid, int nthreads) {
    const int lines_per_slice = (3 * cP.NpS * 8 + 127) >> 7;
    const char* base = reinterpret_cast<const char*>(path + (size_t)b0 * 3 * cP.NpS);
    const int nlines = nslice * lines_per_slice;      // slices are contiguous
    for (int i = tid; i < nlines; i += nthreads) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + ((size_t)i << 7)));
}

}  // namespace pigs
