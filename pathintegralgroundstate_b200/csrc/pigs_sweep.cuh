// pigs_sweep.cuh -- the Monte-Carlo moves, the estimators and the driver
// schedule, executed by one chain group (see pigs_device.cuh).
//
// Reference mapping: moves = vpi_mod.f90:313-2487, action = :2491-2841,
// estimators = sample_mod.f90:13-594, schedule = vpi.f90:297-475.
//
// How a move is organised (all 14 share it).  The reference displaces beads one
// at a time, writing each into Path and calling UpdateAction before the next.
// UpdateAction(ip,ib) only reads the OTHER particles of slice ib, and a move
// only displaces beads of one particle, so all bead-updates of a move are
// independent once the proposed positions are known.  The group therefore
//   1. copies the particle's segment into shared memory (seg_old / seg_new),
//   2. draws the Gaussians in the reference's order and builds the proposal
//      (free end, Levy bridge or bisection level) in seg_new,
//   3. evaluates every displaced bead's DeltaS in parallel (beads over warps,
//      partners over lanes),
//   4. answers the Metropolis question once, identically in every thread,
//   5. writes seg_new back into the path only on acceptance (the reference
//      writes eagerly and restores on rejection -- same end state).
//
// All functions are templates on <MT, VAR>: MT = replay of the reference's
// MT19937 stream (else Philox), VAR = table placement / trap (VarTraits).
// `ctr` is the thread-private Philox counter (group-uniform by construction).
#pragma once

#include "pigs_device.cuh"

namespace pigs {

typedef RngS ull;      // the RNG stream state threaded through the moves

// (measured: making the action evaluation a real call costs 35% -- call-site spills;
// it stays inlined into the move engine)

#ifndef PIGS_EVAL_INLINE
#define PIGS_EVAL_INLINE __forceinline__
#endif
#define PIGS_T template <bool MT, int VAR>
#define PIGS_TRAP (VarTraits<VAR>::TRAP)
#define PIGS_VSM (VarTraits<VAR>::VSM)
#define PIGS_WSM (VarTraits<VAR>::WSM)
#define PIGS_VPAIR (VarTraits<VAR>::VPAIR)

template <int VAR> __device__ __forceinline__ double bc_wrap(int k, double x) {      // BoundaryConditions unless trap
    return PIGS_TRAP ? x : mimg(x, cP.L[k], cP.Lh[k]);
}
template <int VAR> __device__ __forceinline__ double wrap_lt(int k, double d) {      // in-line wrap of the bridge code
    return PIGS_TRAP ? d : mimg_lt_first(d, cP.L[k], cP.Lh[k]);
}
__device__ __forceinline__ void bump(GS* gs, int c) {
    if (gfirst(gs)) gs->cnt[c] += 1;
}

// ---------------------------------------------------------------- small-integer helpers (no IDIV in hot code)
__device__ __forceinline__ int div_dim(int i) {          // i / cP.dim for dim in 1..3, 0 <= i < 65536
    return cP.dim == 3 ? (i * 43691) >> 17 : (cP.dim == 2 ? i >> 1 : i);
}

// ---------------------------------------------------------------- action
// Sum over beads ib = b0 + m*bstride (m<nb) of w_m * DeltaS(ip,ib,seg_new,seg_old),
// w_0 = wfirst, w_{nb-1} = wlast, else 1.  Identical in every thread.
// Work split: with at least as many beads as warps each warp takes whole beads
// (split = 1) and reduces to DeltaS with shuffles only; with fewer beads the
// partners of a bead are divided over `split` warps and combined through smem.
template <int VAR, bool XR>
static __device__ PIGS_EVAL_INLINE double eval_action(GS* gs, int ip0, int b0, int bstride, int nb, double wfirst,
                                                      double wlast, bool roll) {
    const Grp G = grp(gs);
    if (G.tid == 0) {
        // bead-update counters by slice class, in closed form (beads b0, b0+bstride, ...)
        const int last = b0 + (nb - 1) * bstride;
        const int nend = (b0 == 0 ? 1 : 0) + ((last == 2 * cP.Nb && last != 0) ? 1 : 0) - ((nb == 1 && b0 == 0 && last == 2 * cP.Nb) ? 1 : 0);
        int nodd;
        if (bstride & 1) nodd = ((last + 1) >> 1) - (b0 >> 1);
        else nodd = (b0 & 1) ? nb : 0;
        gs->cnt[C_UPD_END] += nend;
        gs->cnt[C_UPD_ODD] += nodd;
        gs->cnt[C_UPD_EVEN] += nb - nodd - nend;
    }
    const int nw = G.nwarps;
#ifndef PIGS_NO_WARP_FASTPATH
    if (nw == 1) {
        // one warp per chain (the production shape): beads in order, the whole partner range per bead; no task
        // decomposition, no partial-sum exchange
        const int lane = G.lane;
        const bool have = lane < cP.Np;
        constexpr bool TWO = (PIGS_LOOPV & (256 | 1024)) && !PIGS_TRAP && !PIGS_VPAIR;      // carried partner stream, see pair_loop2
        Partner first;
        Carry cy;
        first.x = first.y = first.z = 0.0;
        cy.a = first; cy.b = first; cy.K = nullptr;
#if PIGS_KC >= 1
        LoopK K;
        { const volatile double* kc = gs->kc; K.iL0 = kc[0]; K.iL1 = kc[1]; K.iL2 = kc[2]; K.idr = kc[3];
        }
        cy.K = &K;
#endif
        const double* Rx = slice(gs, b0);                                  // walks the evaluated slices
        const long long sstride = (long long)bstride * 3 * cP.NpS, pfoff = (long long)cA.pfdist * sstride;
        if (have) first = load_partner(Rx, lane);
        if (TWO) { cy.a = first; if ((PIGS_LOOPV & 256) && lane + 32 < cP.Np) cy.b = load_partner(Rx, lane + 32); }
        double Sw = 0.0, Slin = 0.0;       // Slin: per-lane sum of the terms linear in the pair sums, reduced once below
        for (int m = 0; m < nb; ++m, Rx += sstride) {
            const int ib = b0 + m * bstride;
            double xo[3], xn[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { xo[k] = so(gs, k, ib); xn[k] = sn(gs, k, ib); }
            const Partner cur = first;
            if (TWO) cy.next = (m + 1 < nb) ? Rx + sstride : nullptr;
            else if (m + 1 < nb && have) first = load_partner(Rx + sstride, lane);
            if (roll && lane == 0 && m + cA.pfdist < nb) prefetch_slice_L2(Rx + pfoff);
            double lin;
            const double t = bead_eval<PIGS_TRAP, PIGS_VSM, PIGS_WSM, PIGS_VPAIR, XR>(Rx, ip0, ib, lane, 32, lane == 0, xo, xn,
                                                                                 lane, nullptr, cur, &lin, TWO ? &cy : nullptr);
            const double w = (m == 0) ? wfirst : ((m == nb - 1) ? wlast : 1.0);
            Sw += w * t;
            Slin += w * lin;
        }
        return Sw + warp_sum(Slin);
    }
#endif
    int split = 1;
    if (nb < nw) {
        const int nch = (cP.Np + 31) >> 5;
        split = nw / nb;
        if (split > nch) split = nch;
    }
    double* part = part_of(gs);
    const int ntask = nb * split;
    double Sw = 0.0;
    // the first partner of a task is loaded one task ahead, so its HBM/L2 latency hides behind the previous
    // task's reduction (or, for the first task, behind the coordinate reads below)
    Partner first;
    first.x = first.y = first.z = 0.0;
    if (G.warp < ntask) {
        int m = (split > 1) ? G.warp / split : G.warp, s = (split > 1) ? G.warp - m * split : 0;
        if (s * 32 + G.lane < cP.Np) first = load_partner(slice(gs, b0 + m * bstride), s * 32 + G.lane);
    }
    for (int task = G.warp; task < ntask; task += nw) {
        int m = task, s = 0;
        if (split > 1) { m = task / split; s = task - m * split; }
        const int ib = b0 + m * bstride;
        double xo[3], xn[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { xo[k] = so(gs, k, ib); xn[k] = sn(gs, k, ib); }
        const Partner cur = first;
        const int tn = task + nw;
        if (roll && G.lane == 0 && s == 0) {       // rolling L2 prefetch: the slice this warp reads two tasks from now
            const int m2 = m + cA.pfdist * ((split > 1) ? (nw / split) : nw);
            if (m2 < nb) prefetch_slice_L2(slice(gs, b0 + m2 * bstride));
        }
        if (tn < ntask) {
            int mn = (split > 1) ? tn / split : tn, sn_ = (split > 1) ? tn - mn * split : 0;
            if (sn_ * 32 + G.lane < cP.Np) first = load_partner(slice(gs, b0 + mn * bstride), sn_ * 32 + G.lane);
        }
        double t = bead_eval<PIGS_TRAP, PIGS_VSM, PIGS_WSM, PIGS_VPAIR, XR>(slice(gs, ib), ip0, ib, s * 32 + G.lane, 32 * split,
                                                           (s == 0) && (G.lane == 0), xo, xn, G.lane,
                                                           (split == 1) ? nullptr : part + task * 8, cur);
        if (split == 1) {
            double w = (m == 0) ? wfirst : ((m == nb - 1) ? wlast : 1.0);
            Sw += w * t;
        }
    }
    if (nw == 1) return Sw;
    double S = 0.0;
    if (split == 1) {
        if (G.lane == 0) part[G.warp] = Sw;
        gsync(gs);
        for (int w = 0; w < nw; ++w) S += part[w];
    } else {
        gsync(gs);
        for (int m = 0; m < nb; ++m) {
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                double acc = 0.0;
                for (int s = 0; s < split; ++s) acc += part[(m * split + s) * 8 + q];
                v[q] = acc;
            }
            double w = (m == 0) ? wfirst : ((m == nb - 1) ? wlast : 1.0);
            S += w * assemble_dS(b0 + m * bstride, v);
        }
    }
    gsync(gs);
    return S;
}
PIGS_T __device__ __forceinline__ double uniform(GS* gs, ull& ctr) { return rng_uniform<MT>(gs, ctr); }
PIGS_T __device__ __forceinline__ int draw_int(GS* gs, ull& ctr, int n) {       // int(n*grnd()), clamped for u==1 (Q6)
    int v = (int)((double)n * rng_uniform<MT>(gs, ctr));
    return v >= n ? (n > 0 ? n - 1 : 0) : v;
}
// DeltaK of the broken/mended link (vpi_mod.f90:1859-1873, 2205-2219)
template <int VAR>
__device__ __forceinline__ double link_DeltaK(const double* seg, int ii, int ie, int Ls) {
    double r2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) if (k < cP.dim) {
        double d = seg[k * cP.S + ii] - seg[k * cP.S + ie];
        if (!PIGS_TRAP) d = mimg(d, cP.L[k], cP.Lh[k]);
        r2 += d * d;
    }
    return -0.5 * r2 / ((double)Ls * cP.dt) - 0.5 * (double)cP.dim * log(2.0 * cP.pi * (double)Ls * cP.dt);
}

// ---------------------------------------------------------------- the move engine
// Every move of vpi_mod.f90 is an instance of one pipeline; a descriptor word
// selects the variant.  One copy of each stage => the hot code stays inside the
// instruction cache and no stage is a function call.
enum MoveFlags {
    MV_TRANSLATE = 0, MV_BRIDGE = 1, MV_BISECT = 2, MV_TYPE_MASK = 3,
    MV_HALF1 = 1 << 2, MV_HALF2 = 2 << 2, MV_HALF_MASK = 3 << 2,   // first Path(:,ip,Nb) = xend(:,half)
    MV_FREE_NEXT = 1 << 4,        // free end = bead ii, anchored on ie    (head-like)
    MV_FREE_PREV = 1 << 5,        // free end = bead ie, anchored on ii    (tail-like)
    MV_WFIRST_HALF = 1 << 6,      // first evaluated bead weighs 1/2 (cut bead)
    MV_WLAST_HALF = 1 << 7,       // last evaluated bead weighs 1/2
    MV_GLUE_II_X1 = 1 << 8,       // seg_new(ii) = xend(:,1) before the bridge   (CloseChain, half 2)
    MV_GLUE_IE_X2 = 1 << 9,       // seg_new(ie) = xend(:,2)                     (CloseChain half 1, Swap)
    MV_DK_OLD_ADD = 1 << 10,      // + DeltaK of the old link                    (OpenChain)
    MV_DK_NEW_SUB = 1 << 11,      // - DeltaK of the new link                    (CloseChain)
};
// [m0,m1] = beads evaluated and, on acceptance, committed.  Returns acceptance
// (group-uniform).  seg_old/seg_new stay valid for the caller's epilogue.
//
// Register budget: the partner loop inside eval_action wants every register of
// the 128 a thread has (16 warps/SM), so nothing of the move's own state is
// allowed to stay live across it.  The move descriptor and the RNG state are
// parked in the group's shared-memory block (gs->pk) by the prologue; each phase
// is pre (reads pk, draws, proposes) -> eval_action -> post (reads pk, Metropolis),
// with compiler memory barriers on both sides of the evaluation so the values
// are re-read instead of carried.  All threads of a group hold identical copies,
// so the parked words are written with identical values by all of them; the RNG
// state, the only word that changes, ping-pongs between two slots by phase
// parity (a thread ahead by less than one group sync never overwrites the slot
// a slower thread still has to read).
struct PhaseGeom {
    int type, L, Nl, nphase, lev, delta_ib, b0, bs, nb, iend, ianc;
    bool has_free, gate;
};
__device__ __forceinline__ PhaseGeom phase_geom(int flags, int ii, int ie, int m0, int m1, int ph) {
    PhaseGeom g;
    g.type = flags & MV_TYPE_MASK;
    g.L = ie - ii;
    g.has_free = flags & (MV_FREE_NEXT | MV_FREE_PREV);
    g.iend = (flags & MV_FREE_NEXT) ? ii : ie;
    g.ianc = (flags & MV_FREE_NEXT) ? ie : ii;
    g.Nl = 31 - __clz(g.L);                                          // bisection: L = 2^Nl
    g.nphase = (g.type == MV_BISECT) ? g.Nl + (g.has_free ? 1 : 0) : 1;
    g.gate = (g.type == MV_BISECT) && g.has_free && ph == 0;         // free end of Move{Head,Tail}Bisection
    g.lev = (g.type == MV_BISECT) ? ph + (g.has_free ? 0 : 1) : 0;   // 1..Nl
    g.delta_ib = (g.type == MV_BISECT && !g.gate) ? (1 << (g.Nl - g.lev + 1)) : 2;
    if (g.type == MV_BISECT) {
        if (g.gate) { g.b0 = g.iend; g.bs = 1; g.nb = 1; }
        else { g.b0 = ii + (g.delta_ib >> 1); g.bs = g.delta_ib; g.nb = 1 << (g.lev - 1); }
    } else { g.b0 = m0; g.bs = 1; g.nb = m1 - m0 + 1; }
    return g;
}

PIGS_T __device__ __forceinline__ void move_prologue(GS* gs, ull* pctr, int flags, int ip0, int ii, int ie, int m0, int m1,
                                                     double Sbase) {
    const Grp G = grp(gs);
    ull ctr = *pctr;
    const int type = flags & MV_TYPE_MASK;
    const int half = (flags & MV_HALF_MASK) >> 2;
    const int L = ie - ii;
    const int dim = cP.dim;
    if (cA.prefetch == 1) prefetch_slices(gs->path, (type == MV_TRANSLATE) ? ii : m0, ((type == MV_TRANSLATE) ? ie : m1) - ((type == MV_TRANSLATE) ? ii : m0) + 1, G.tid, G.size);
    if (cA.prefetch == 3) {
        // L2 prefetch schedule: the first slices a move evaluates are requested here, before the segment is even
        // read; later ones one bisection level ahead (phase_pre) or two beads ahead (rolling, inside eval_action),
        // so at most a few slices per chain are in flight and the L2 working set of all resident chains stays small.
        if (type == MV_BISECT) {
            if (G.tid == 0) prefetch_slice_L2(slice(gs, (flags & (MV_FREE_NEXT | MV_FREE_PREV)) ? ((flags & MV_FREE_NEXT) ? ii : ie) : ii + (L >> 1)));
        } else {
            const int nfirst = min(cA.pfdist * G.nwarps, m1 - m0 + 1);
            if (G.tid < nfirst) prefetch_slice_L2(slice(gs, m0 + G.tid));
        }
    }
    if (half) {
        if (G.tid < dim) pth(gs, G.tid, ip0, cP.Nb) = gs->xend[(half - 1) * 3 + G.tid];
        gsync(gs);
    }
    // ---- load the segment (translate: shifted copy)
    {
        double dx[3] = {0.0, 0.0, 0.0};
        if (type == MV_TRANSLATE) {
#pragma unroll
            for (int k = 0; k < 3; ++k) if (k < dim) dx[k] = cP.delta_cm * (2.0 * uniform<MT, VAR>(gs, ctr) - 1.0);
        }
        const int n = L + 1;
        for (int i = G.tid; i < 3 * n; i += G.size) {
            int k = (i >= n) + (i >= 2 * n), ib = ii + (i - k * n);
            double v = pth(gs, k, ip0, ib);
            so(gs, k, ib) = v;
            if (type == MV_TRANSLATE && k < dim) v = bc_wrap<VAR>(k, v + ((k == 0) ? dx[0] : ((k == 1) ? dx[1] : dx[2])));
            sn(gs, k, ib) = v;
        }
        gsync(gs);
    }
    if (flags & (MV_GLUE_II_X1 | MV_GLUE_IE_X2)) {
        if (G.tid < dim) {
            if (flags & MV_GLUE_II_X1) sn(gs, G.tid, ii) = gs->xend[G.tid];
            else sn(gs, G.tid, ie) = gs->xend[3 + G.tid];
        }
        gsync(gs);
    }
    double DeltaK = 0.0;
    if (flags & MV_DK_OLD_ADD) DeltaK = link_DeltaK<VAR>(seg_old(gs), ii, ie, L);
    // beads that need Gaussians: interior + free end
    const int g0 = (flags & MV_FREE_NEXT) ? ii : ii + 1, g1 = (flags & MV_FREE_PREV) ? ie : ie - 1;
    if (!MT && type != MV_TRANSLATE) rng_gauss_fill<MT>(gs, &ctr, dim, g0, 1, g1 - g0 + 1);     // Philox: one pass per move
    MovePark& pk = gs->pk;
    pk.ctr[0] = ctr;
    pk.Sbase = Sbase; pk.DeltaK = DeltaK;
    pk.flags = flags; pk.ip0 = ip0; pk.ii = ii; pk.ie = ie; pk.m0 = m0; pk.m1 = m1;
}

// proposal of phase ph; returns the beads to evaluate and their end weights
PIGS_T __device__ __forceinline__ void phase_pre(GS* gs, int ph, int& b0, int& bs, int& nb, double& wf, double& wl, bool& roll) {
    const Grp G = grp(gs);
    const MovePark& pk = gs->pk;
    const int flags = pk.flags, ii = pk.ii, ie = pk.ie;
    const PhaseGeom g = phase_geom(flags, ii, ie, pk.m0, pk.m1, ph);
    const int type = g.type, L = g.L, dim = cP.dim;
    b0 = g.b0; bs = g.bs; nb = g.nb;
    if (cA.prefetch == 2 && G.tid < nb) prefetch_slice_L2(slice(gs, b0 + G.tid * bs));
    if (cA.prefetch == 3 && type == MV_BISECT && ph + 1 < g.nphase) {        // one level ahead
        const int lev1 = ph + 1 + (g.has_free ? 0 : 1), d1 = 1 << (g.Nl - lev1 + 1);
        if (G.tid < (1 << (lev1 - 1))) prefetch_slice_L2(slice(gs, ii + (d1 >> 1) + G.tid * d1));
    }
    if (MT) {      // the reference's draw order (Appendix A of SURVEY.md); the MT state lives in HBM, not in ctr
        ull dummy;
        if (type == MV_BRIDGE) {
            const int g0 = (flags & MV_FREE_NEXT) ? ii : ii + 1, g1 = (flags & MV_FREE_PREV) ? ie : ie - 1;
            if (flags & MV_FREE_PREV) { rng_gauss_fill<MT>(gs, &dummy, dim, ie, 1, 1); rng_gauss_fill<MT>(gs, &dummy, dim, ii + 1, 1, L - 1); }
            else rng_gauss_fill<MT>(gs, &dummy, dim, g0, 1, g1 - g0 + 1);
        } else if (type == MV_BISECT) {
            rng_gauss_fill<MT>(gs, &dummy, dim, b0, bs, nb);
        }
    }
    if (type == MV_BRIDGE || g.gate) {
        if (G.tid < dim) {
            const int k = G.tid;
            if (g.has_free) {      // xnew = BC(unwrap(anchor) + sigma*g)    (vpi_mod.f90:619-645 / 758-785)
                double xold = so(gs, k, g.iend), gz = sn(gs, k, g.iend), anc = sn(gs, k, g.ianc), base;
                if (flags & MV_FREE_NEXT) base = xold - wrap_lt<VAR>(k, xold - anc);
                else base = xold + wrap_lt<VAR>(k, anc - xold);
                sn(gs, k, g.iend) = bc_wrap<VAR>(k, base + cP.sig_free[L] * gz);
            }
            if (type == MV_BRIDGE) {      // Levy bridge ii -> ie (vpi_mod.f90:509-549)
                double pnext = sn(gs, k, ie), pprev = sn(gs, k, ii);
                for (int j = 1; j <= L - 1; ++j) {
                    int ib = ii + j;
                    double xold = so(gs, k, ib), gz = sn(gs, k, ib);
                    double xprev = xold + wrap_lt<VAR>(k, pprev - xold);
                    double xnext = xold - wrap_lt<VAR>(k, xold - pnext);
                    double sigma = cP.sig_stage[L - j];                                           // float32 ratio (Q15)
                    double xmid = (xnext + xprev * (double)(L - j)) / (double)(L - j + 1);
                    pprev = bc_wrap<VAR>(k, xmid + sigma * gz);
                    sn(gs, k, ib) = pprev;
                }
            }
        }
        gsync(gs);
    } else if (type == MV_BISECT) {      // one bisection level (vpi_mod.f90:905-956)
        const double sigma = cP.sig_bis[g.Nl - g.lev + 1];        // delta_ib = 2^(Nl-lev+1)
        for (int i = G.tid; i < nb * dim; i += G.size) {
            int j = div_dim(i), k = i - j * dim;
            int iprev = ii + j * g.delta_ib, inext = iprev + g.delta_ib, icurr = (iprev + inext) >> 1;
            double xold = so(gs, k, icurr), gz = sn(gs, k, icurr);
            double xprev = xold + wrap_lt<VAR>(k, sn(gs, k, iprev) - xold);
            double xnext = xold - wrap_lt<VAR>(k, xold - sn(gs, k, inext));
            sn(gs, k, icurr) = bc_wrap<VAR>(k, 0.5 * (xprev + xnext) + sigma * gz);
        }
        gsync(gs);
    }
    wf = (type != MV_BISECT && (flags & MV_WFIRST_HALF)) ? 0.5 : 1.0;
    wl = (type != MV_BISECT && (flags & MV_WLAST_HALF)) ? 0.5 : 1.0;
    roll = cA.prefetch == 3 && type != MV_BISECT;
}

// the Metropolis question of phase ph: 0 = go on with the next phase, 1 = rejected, 2 = accepted (last phase)
PIGS_T __device__ __forceinline__ int phase_post(GS* gs, int ph, double S) {
    MovePark& pk = gs->pk;
    const int flags = pk.flags, ii = pk.ii, ie = pk.ie;
    const int type = flags & MV_TYPE_MASK;
    ull ctr = pk.ctr[ph & 1];
    if (type != MV_BISECT) {
        S += pk.Sbase;
        if (flags & MV_DK_OLD_ADD) S += pk.DeltaK;
        if (flags & MV_DK_NEW_SUB) S -= link_DeltaK<VAR>(seg_new(gs), ii, ie, ie - ii);
    }
    // (e.g. vpi_mod.f90:356-364): a uniform is consumed only if exp(-S)<1
    bool accept = true;
    if (!(S <= 0.0)) {
        double e = exp(-S);
        if (!(e >= 1.0)) {                     // NaN falls through: rejected after its draw (Q23)
            double u = uniform<MT, VAR>(gs, ctr);
            if (!(e >= u)) accept = false;
        }
    }
    pk.ctr[(ph + 1) & 1] = ctr;
    if (!accept) return 1;
    gsync(gs);
    const PhaseGeom g = phase_geom(flags, ii, ie, 0, 0, ph);
    return (ph + 1 == g.nphase) ? 2 : 0;
}

// The engine proper.  It is instantiated exactly twice per kernel: inlined into
// the kernel body at the single call site of the diagonal sweep (diag_sweep,
// >95% of all moves), and once as the out-of-line run_move below for the rare
// callers (worm and half-chain moves, unit calls).  The inlined copy matters:
// ptxas compiles a loop inside an ABI callee markedly worse than the same loop
// inside the kernel function -- loop-invariant constants are re-read from the
// constant bank every iteration and the global-load descriptor is re-copied
// into uniform registers before every load (the isolated partner loop,
// scripts/loopbench*.cu, runs 17% slower behind a __noinline__ wrapper).
PIGS_T __device__ __forceinline__ bool run_move_body(GS* gs, ull* pctr, int flags, int ip0, int ii, int ie, int m0, int m1,
                                                     double Sbase) {
    move_prologue<MT, VAR>(gs, pctr, flags, ip0, ii, ie, m0, m1, Sbase);
    int ph = 0, r;
    for (;; ++ph) {
        int b0, bs, nb;
        double wf, wl;
        bool roll;
        asm volatile("" ::: "memory");
        phase_pre<MT, VAR>(gs, ph, b0, bs, nb, wf, wl, roll);
        const int ipp = gs->pk.ip0;
        asm volatile("" ::: "memory");
        const double S = eval_action<VAR, MT>(gs, ipp, b0, bs, nb, wf, wl, roll);      // replay: the reference's own roundings
        asm volatile("" ::: "memory");
        r = phase_post<MT, VAR>(gs, ph, S);
        if (r) break;
    }
    asm volatile("" ::: "memory");
    const Grp G = grp(gs);
    const MovePark& pk = gs->pk;
    const bool accept = (r == 2);
    if (accept) {
        const int type = pk.flags & MV_TYPE_MASK, ip1 = pk.ip0, dim = cP.dim;
        const int c0 = (type == MV_TRANSLATE) ? pk.ii : pk.m0, n = ((type == MV_TRANSLATE) ? pk.ie : pk.m1) - c0 + 1;
        for (int i = G.tid; i < 3 * n; i += G.size) {
            int k = (i >= n) + (i >= 2 * n), ib = c0 + (i - k * n);
            if (k < dim) pth(gs, k, ip1, ib) = sn(gs, k, ib);
        }
    }
    const ull cfin = pk.ctr[(ph + 1) & 1];
    gsync(gs);
    *pctr = cfin;
    return accept;
}
PIGS_T static __device__ __noinline__ bool run_move(GS* gs, ull* pctr, int flags, int ip0, int ii, int ie, int m0, int m1,
                                                    double Sbase) {
    return run_move_body<MT, VAR>(gs, pctr, flags, ip0, ii, ie, m0, m1, Sbase);
}

// ---------------------------------------------------------------- the 14 moves as descriptors
PIGS_T __device__ __forceinline__ void TranslateChain(GS* gs, ull* pctr, int ip0) {              // vpi_mod.f90:313-379
    if (run_move<MT, VAR>(gs, pctr, MV_TRANSLATE, ip0, 0, 2 * cP.Nb, 0, 2 * cP.Nb, 0.0)) bump(gs, C_ACC_CM);
}
PIGS_T __device__ __forceinline__ void TranslateHalfChain(GS* gs, ull* pctr, int half, int ip0) { // vpi_mod.f90:383-476
    const int ibi = (half == 1) ? 0 : cP.Nb, ibf = ibi + cP.Nb;
    // cut bead at full weight (Q21)
    if (run_move<MT, VAR>(gs, pctr, MV_TRANSLATE | (half << 2), ip0, ibi, ibf, ibi, ibf, 0.0)) {
        bump(gs, C_ACC_CM_HALF);
        const int t = threadIdx.x & (gsize_of(gs) - 1);
        if (t < cP.dim) gs->xend[(half - 1) * 3 + t] = sn(gs, t, cP.Nb);
        gsync(gs);
    }
}
PIGS_T __device__ __forceinline__ void Staging(GS* gs, ull* pctr, int L, int ip0) {              // vpi_mod.f90:480-578
    ull ctr = *pctr;
    int ii = draw_int<MT, VAR>(gs, ctr, 2 * cP.Nb - L + 1);
    *pctr = ctr;
    if (run_move<MT, VAR>(gs, pctr, MV_BRIDGE, ip0, ii, ii + L, ii + 1, ii + L - 1, 0.0)) bump(gs, C_ACC_BD);
}
PIGS_T __device__ __forceinline__ void StagingHalfChain(GS* gs, ull* pctr, int half, int L, int ip0) {   // vpi_mod.f90:1376-1491
    ull ctr = *pctr;
    int ii = draw_int<MT, VAR>(gs, ctr, cP.Nb - L + 1) + (half == 1 ? 0 : cP.Nb);
    *pctr = ctr;
    // interior beads never include the cut bead Nb, so xend(:,half) is unchanged on acceptance
    if (run_move<MT, VAR>(gs, pctr, MV_BRIDGE | (half << 2), ip0, ii, ii + L, ii + 1, ii + L - 1, 0.0)) bump(gs, C_ACC_BD_HALF);
}
// MoveHead (vpi_mod.f90:582-720) for half == 0, MoveHeadHalfChain (:1495-1656) otherwise
PIGS_T __device__ __forceinline__ void MoveHead(GS* gs, ull* pctr, int Lmax, int ip0, int half) {
    ull ctr = *pctr;
    int Ls = draw_int<MT, VAR>(gs, ctr, Lmax - 1) + 2;
    *pctr = ctr;
    int ii = (half == 2) ? cP.Nb : 0, ie = ii + Ls;
    int fl = MV_BRIDGE | MV_FREE_NEXT | (half << 2) | (half == 2 ? MV_WFIRST_HALF : 0);     // cut bead weighs 1/2 (:1573-1577)
    if (run_move<MT, VAR>(gs, pctr, fl, ip0, ii, ie, ii, ie - 1, 0.0)) {
        bump(gs, half ? C_ACC_HEAD_HALF : C_ACC_HEAD);
        if (half == 2) {       // the free end IS the cut bead
            const int t = threadIdx.x & (gsize_of(gs) - 1);
            if (t < cP.dim) gs->xend[3 + t] = sn(gs, t, cP.Nb);
            gsync(gs);
        }
    }
}
// MoveTail (vpi_mod.f90:724-860) for half == 0, MoveTailHalfChain (:1660-1817) otherwise
PIGS_T __device__ __forceinline__ void MoveTail(GS* gs, ull* pctr, int Lmax, int ip0, int half) {
    ull ctr = *pctr;
    int Ls = draw_int<MT, VAR>(gs, ctr, Lmax - 1) + 2;
    *pctr = ctr;
    int ie = (half == 1) ? cP.Nb : 2 * cP.Nb, ii = ie - Ls;
    int fl = MV_BRIDGE | MV_FREE_PREV | (half << 2) | (half == 1 ? MV_WLAST_HALF : 0);      // (:1734-1738)
    if (run_move<MT, VAR>(gs, pctr, fl, ip0, ii, ie, ii + 1, ie, 0.0)) {
        bump(gs, half ? C_ACC_TAIL_HALF : C_ACC_TAIL);
        if (half == 1) {
            const int t = threadIdx.x & (gsize_of(gs) - 1);
            if (t < cP.dim) gs->xend[t] = sn(gs, t, cP.Nb);
            gsync(gs);
        }
    }
}
PIGS_T __device__ __forceinline__ void Bisection(GS* gs, ull* pctr, int level, int ip0) {        // vpi_mod.f90:864-998
    ull ctr = *pctr;
    int ii = draw_int<MT, VAR>(gs, ctr, 2 * cP.Nb - (1 << level) + 1), ie = ii + (1 << level);
    *pctr = ctr;
    if (run_move<MT, VAR>(gs, pctr, MV_BISECT, ip0, ii, ie, ii + 1, ie - 1, 0.0)) bump(gs, C_ACC_BD);
}
// MoveHeadBisection (vpi_mod.f90:1002-1184) when head, MoveTailBisection (:1188-1372) otherwise
PIGS_T __device__ __forceinline__ void EndBisection(GS* gs, ull* pctr, int level, int ip0, bool head) {
    ull ctr = *pctr;
    int Nl = draw_int<MT, VAR>(gs, ctr, level - 1) + 2;
    *pctr = ctr;
    int ii = head ? 0 : 2 * cP.Nb - (1 << Nl), ie = ii + (1 << Nl);
    if (run_move<MT, VAR>(gs, pctr, MV_BISECT | (head ? MV_FREE_NEXT : MV_FREE_PREV), ip0, ii, ie, head ? ii : ii + 1,
                          head ? ie - 1 : ie, 0.0))
        bump(gs, head ? C_ACC_HEAD : C_ACC_TAIL);
}
PIGS_T __device__ __forceinline__ int draw_even_Ls(GS* gs, ull& ctr, int Lmax) { return 2 * draw_int<MT, VAR>(gs, ctr, (Lmax - 2) / 2) + 2; }
PIGS_T __device__ __forceinline__ int draw_half(GS* gs, ull& ctr) { int h = (int)(uniform<MT, VAR>(gs, ctr) * 2.0) + 1; return h > 2 ? 2 : h; }

// OpenChain (vpi_mod.f90:1821-2076) when open, CloseChain (:2080-2266) otherwise
PIGS_T static __device__ __noinline__ void OpenClose(GS* gs, ull* pctr, int Lmax, int ip0, bool open) {
    const int t = threadIdx.x & (gsize_of(gs) - 1);
    ull ctr = *pctr;
    int Ls = draw_even_Ls<MT, VAR>(gs, ctr, Lmax), half = draw_half<MT, VAR>(gs, ctr);
    *pctr = ctr;
    const int ii = (half == 1) ? cP.Nb - Ls : cP.Nb, ie = ii + Ls;
    int fl = MV_BRIDGE | (half == 1 ? MV_WLAST_HALF : MV_WFIRST_HALF);
    if (open) fl |= MV_DK_OLD_ADD | (half == 1 ? MV_FREE_PREV : MV_FREE_NEXT);
    else fl |= MV_DK_NEW_SUB | (half == 1 ? MV_GLUE_IE_X2 : MV_GLUE_II_X1);
    const int m0 = (half == 1) ? ii + 1 : ii, m1 = (half == 1) ? ie : ie - 1;
    bool acc = run_move<MT, VAR>(gs, pctr, fl, ip0, ii, ie, m0, m1, open ? -cP.logCd : cP.logCd);
    if (open) {
        if (acc) {
            bump(gs, C_ACC_OPEN);
            if (t < cP.dim) {
                gs->xend[t] = (half == 1) ? sn(gs, t, cP.Nb) : so(gs, t, cP.Nb);
                gs->xend[3 + t] = (half == 1) ? so(gs, t, cP.Nb) : sn(gs, t, cP.Nb);
            }
            if (t == 0) { gs->isopen = 1; gs->new_pc = 1; }
        } else {
            if (t < cP.dim) { gs->xend[t] = so(gs, t, cP.Nb); gs->xend[3 + t] = so(gs, t, cP.Nb); }
            if (t == 0) gs->new_pc = 0;
        }
    } else {
        if (acc) {
            bump(gs, C_ACC_CLOSE);
            if (t < cP.dim) { gs->xend[t] = sn(gs, t, cP.Nb); gs->xend[3 + t] = sn(gs, t, cP.Nb); }
            if (t == 0) { gs->isopen = 0; gs->end_pc = 1; }
        } else {
            if (t == 0) gs->end_pc = 0;
        }
    }
    gsync(gs);
}
PIGS_T static __device__ __noinline__ void Swap(GS* gs, ull* pctr, int Lmax, int iw0) {          // vpi_mod.f90:2270-2487
    const Grp G = grp(gs);
    ull ctr = *pctr;
    if (G.tid == 0) gs->swap_acc = 0;
    int Ls = draw_even_Ls<MT, VAR>(gs, ctr, Lmax), ii = cP.Nb - Ls, ie = cP.Nb;
    const double inv = 1.0 / ((double)Ls * cP.dt);
    double* pp = pp_of(gs);
    double xe2[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 3; ++k) if (k < cP.dim) xe2[k] = gs->xend[3 + k];
    for (int ip = G.tid; ip < cP.Np; ip += G.size) {
        double r2 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) if (k < cP.dim) {
            double d = pth(gs, k, ip, ii) - xe2[k];
            if (!PIGS_TRAP) d = mimg(d, cP.L[k], cP.Lh[k]);
            r2 += d * d;
        }
        pp[ip] = exp(-0.5 * r2 * inv);
    }
    gsync(gs);
    if (G.tid == 0) {
        double Sw = 0.0;
        for (int ip = 0; ip < cP.Np; ++ip) Sw += pp[ip];
        gs->bc[4] = Sw;
    }
    double uran = uniform<MT, VAR>(gs, ctr);
    gsync(gs);
    const double Sw = gs->bc[4];
    if (G.tid == 0) {
        double sum = 0.0;
        int ik = cP.Np - 1;                      // Q18: the reference runs off the array if rounding leaves sum<uran
        for (int ip = 0; ip < cP.Np; ++ip) {
            sum += pp[ip] / Sw;
            if (uran <= sum) { ik = ip; break; }
        }
        gs->ibc[0] = ik;
    }
    gsync(gs);
    const int ik = gs->ibc[0];
    if (ik != iw0) {
        double xk[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < 3; ++k) if (k < cP.dim) xk[k] = pth(gs, k, ik, ie);
        gsync(gs);                                   // everyone has read pp/ibc before pp is reused
        for (int ip = G.tid; ip < cP.Np; ip += G.size) {
            double r2 = 0.0;
#pragma unroll
            for (int k = 0; k < 3; ++k) if (k < cP.dim) {
                double d = pth(gs, k, ip, ii) - xk[k];
                if (!PIGS_TRAP) d = mimg(d, cP.L[k], cP.Lh[k]);
                r2 += d * d;
            }
            pp[ip] = exp(-0.5 * r2 * inv);
        }
        gsync(gs);
        if (G.tid == 0) {
            double Sk = 0.0;
            for (int ip = 0; ip < cP.Np; ++ip) Sk += pp[ip];
            gs->bc[5] = Sk;
        }
        double ug = uniform<MT, VAR>(gs, ctr);      // always consumed (vpi_mod.f90:2373)
        gsync(gs);
        const double Sk = gs->bc[5];
        if (ug <= Sw / Sk) {
            *pctr = ctr;
            bool acc = run_move<MT, VAR>(gs, pctr, MV_BRIDGE | MV_GLUE_IE_X2, ik, ii, ie, ii + 1, ie - 1, 0.0);
            ctr = *pctr;
            if (acc) {
                bump(gs, C_ACC_SWAP);
                // exchange the second halves (vpi_mod.f90:2454-2464)
                const int n = cP.Nb;                     // slices Nb+1..2Nb
                for (int i = G.tid; i < 3 * n; i += G.size) {
                    int k = (i >= n) + (i >= 2 * n), ib = cP.Nb + 1 + (i - k * n);
                    double a = pth(gs, k, iw0, ib), b = pth(gs, k, ik, ib);
                    pth(gs, k, iw0, ib) = b;
                    pth(gs, k, ik, ib) = a;
                }
                if (G.tid < cP.dim) {
                    int k = G.tid;
                    double oldik = so(gs, k, cP.Nb), oldiw = pth(gs, k, iw0, cP.Nb);
                    pth(gs, k, ik, cP.Nb) = oldiw;
                    pth(gs, k, iw0, cP.Nb) = oldik;
                    gs->xend[3 + k] = oldik;
                }
                if (G.tid == 0) { gs->swap_acc = 1; gs->ik0 = ik; }
            }
        }
    }
    gsync(gs);
    *pctr = ctr;
}

// PermutationSampling (sample_mod.f90:530-594).  Thread 0 does the bookkeeping
// (arrays in global memory, scalars in the group's shared state).
static __device__ __noinline__ void PermutationSampling(GS* gs, bool have_swap) {
    if (!cP.swapping) return;        // the reference indexes unallocated arrays here (Q22)
    gsync(gs);
    if (gfirst(gs)) {
        int* cyc = gs->cyc;
        int* hist = gs->hist;
        if (gs->new_pc) {
            for (int i = 0; i < cP.Np; ++i) cyc[i] = 0;
            cyc[0] = gs->iworm0 + 1;
            gs->iperm = 1;
            gs->new_pc = 0;
        }
        if (have_swap && gs->swap_acc) {
            bool found = false;
            for (int i = 0; i < cP.Np; ++i) if (cyc[i] == gs->ik0 + 1) { found = true; break; }
            if (!gs->end_pc && !found && gs->iperm < cP.Np) {
                gs->iperm = gs->iperm < 0 ? 1 : gs->iperm + 1;
                cyc[gs->iperm - 1] = gs->ik0 + 1;
            }
        }
        if (gs->end_pc) {
            // a worm uploaded as open without its permutation record has iperm = 0 (the reference would index
            // Perm_histogram(0)); count it as a cycle of length 1
            int len = gs->iperm < 1 ? 1 : (gs->iperm > cP.Np ? cP.Np : gs->iperm);
            hist[len - 1] += 1;
            if (gs->isopen) {
                for (int i = 0; i < cP.Np; ++i) cyc[i] = 0;
                cyc[0] = gs->iworm0 + 1;
                gs->iperm = 1;
            }
            gs->end_pc = 0;
        }
    }
    gsync(gs);
}

// ---------------------------------------------------------------- estimators
// sum of N doubles over the group, identical in every thread
template <int N>
__device__ __forceinline__ void group_sum(GS* gs, double (&v)[N]) {
    const Grp G = grp(gs);
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    if (G.nwarps == 1) return;
    double* part = part_of(gs);
    if (G.lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) part[G.warp * 8 + i] = v[i];
    }
    gsync(gs);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double s = 0.0;
        for (int w = 0; w < G.nwarps; ++w) s += part[w * 8 + i];
        v[i] = s;
    }
    gsync(gs);
}

// LocalEnergy (sample_mod.f90:154-319) of slice R (SoA).  Thread i owns particle i
// and sums over all j != i; pair quantities are therefore counted twice and halved.
// out[0..2] = E, Kin, Pot (identical in every thread).
template <int VAR>
static __device__ __noinline__ void LocalEnergy(GS* gs, const double* Rx, double* out) {
    const Grp G = grp(gs);
    const double* Ry = Rx + PY;
    const double* Rz = Rx + PZ;
    const double* tV = gs->tabV;  (void)tV;
    const double* tW = gs->tabW;
    double s[4] = {0.0, 0.0, 0.0, 0.0};      // sum |F_i|^2, pair lap (x2), pair pot (x2), one-body pot
    double oneLap = 0.0;
    for (int i = G.tid; i < cP.Np; i += G.size) {
        double xi[3] = {Rx[pidx(i)], Ry[pidx(i)], Rz[pidx(i)]};
        double F[3] = {0.0, 0.0, 0.0}, lap = 0.0, pot = 0.0;
        if (PIGS_TRAP) {
#pragma unroll
            for (int k = 0; k < 3; ++k) if (k < cP.dim) {
                double ak = cP.a_ho[k], a2 = ak * ak;
                F[k] = -(xi[k] / a2);
                s[3] += 0.5 * xi[k] * xi[k] / (a2 * a2);
                oneLap += -1.0 / a2;
            }
        }
        for (int j = 0; j < cP.Np; ++j) {
            if (j == i) continue;
            double d0 = xi[0] - Rx[pidx(j)], d1 = xi[1] - Ry[pidx(j)], d2 = xi[2] - Rz[pidx(j)];
            if (!PIGS_TRAP) {
                d0 = mimg(d0, cP.L[0], cP.Lh[0]); d1 = mimg(d1, cP.L[1], cP.Lh[1]); d2 = mimg(d2, cP.L[2], cP.Lh[2]);
            }
            // reference arithmetic, uncontracted (see lk_exact_d1_d2)
            double r2 = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
            if (PIGS_TRAP || r2 <= cP.rcut2) {
                double r = sqrt(r2);
                double dudr, d2u;
                lk_exact_d1_d2<PIGS_WSM>(tW, r, dudr, d2u);        // only u'' needs the reference's exact arithmetic
                double q = dudr / r;
                lap += fma((double)(cP.dim - 1), q, d2u);
                F[0] += q * d0; F[1] += q * d1; F[2] += q * d2;
                Lk k = lk_prep(r);
                if (PIGS_TRAP) k.i0 = min(k.i0, cP.Nmax - 1);
                pot += PIGS_VPAIR ? lk_val_pair(k) : lk_val<PIGS_VSM, 0, PIGS_VSM>(k);
            }
        }
        s[0] += F[0] * F[0] + F[1] * F[1] + F[2] * F[2];
        s[1] += lap;
        s[2] += pot;
    }
    group_sum<4>(gs, s);
    double t[1] = {oneLap};
    if (PIGS_TRAP) group_sum<1>(gs, t);
    double LapLogPsi = 0.5 * t[0] + 0.5 * s[1];
    double Pot = s[3] + 0.5 * s[2];
    double Kin = -0.5 * (2.0 * LapLogPsi + s[0]);
    out[0] = Kin + Pot; out[1] = Kin; out[2] = Pot;
}

// ThermEnergy (sample_mod.f90:323-388) incl. PotentialEnergy (:13-150) of every slice.
// GreenFunction(opt=1) is linear in (Pot, F2), so every (slice, particle) item adds its
// weighted share directly; no per-slice reduction is needed.  out = E, Ec, Ep.
template <int VAR>
static __device__ __forceinline__ void ThermEnergy(GS* gs, double* out) {      // one call site per kernel: inlined (see run_move_body)
    const Grp G = grp(gs);
    const int nitem = 2 * cP.Nb * cP.Np;
    const double dt = cP.dt;
    const double* tV = gs->tabV;  (void)tV;
    double s[2] = {0.0, 0.0};       // E sum, Ep
    for (int it = G.tid; it < nitem; it += G.size) {
        int ib = it / cP.Np, i = it - ib * cP.Np;
        const double* Rx = slice(gs, ib);
        const double* Ry = Rx + PY;
        const double* Rz = Rx + PZ;
        const bool odd = (ib & 1) && !cP.primitive;        // slices that need the force term
        double xi[3] = {Rx[pidx(i)], Ry[pidx(i)], Rz[pidx(i)]};
        double F[3] = {0.0, 0.0, 0.0}, pot = 0.0, one = 0.0;
        if (PIGS_TRAP) {
#pragma unroll
            for (int k = 0; k < 3; ++k) if (k < cP.dim) {
                double ak = cP.a_ho[k], a4 = ak * ak * ak * ak;
                F[k] = xi[k] / a4;
                one += 0.5 * xi[k] * xi[k] / a4;
            }
        }
        if (odd) {
            // forces are needed for every particle: particle i sums over all j (each pair is seen twice)
            for (int j = 0; j < cP.Np; ++j) {
                if (j == i) continue;
                double d0 = xi[0] - Rx[pidx(j)], d1 = xi[1] - Ry[pidx(j)], d2 = xi[2] - Rz[pidx(j)];
                if (!PIGS_TRAP) {      // the estimators take the reference's decisions bit for bit (see mimg_cmp, r2_ref)
                    d0 = mimg_cmp(d0, cP.L[0], cP.Lh[0]); d1 = mimg_cmp(d1, cP.L[1], cP.Lh[1]); d2 = mimg_cmp(d2, cP.L[2], cP.Lh[2]);
                }
                double r2 = r2_ref(d0, d1, d2);
                if (PIGS_TRAP || r2 <= cP.rcut2) {
                    double ir = rsqrt_pos(r2);
                    double r = r2 * ir;
                    Lk k = lk_prep(r);
                    if (PIGS_TRAP) k.i0 = min(k.i0, cP.Nmax - 1);
                    double v, dv;
                    if (PIGS_VPAIR) lk_val_d1_pair(k, v, dv); else lk_val_d1<PIGS_VSM, 0, PIGS_VSM>(k, v, dv);
                    pot += v;
                    double q = dv * ir;
                    F[0] += q * d0; F[1] += q * d1; F[2] += q * d2;
                }
            }
        } else {
            // potential only: every pair once -- particle i takes the partners (i+m) mod N, m = 1..N/2
            // (for even N the m = N/2 pairs are taken from i < N/2 only); pot is doubled so that the common
            // "half of the double-counted sum" below applies
            const int half = cP.Np >> 1;
            for (int m = 1; m <= half; ++m) {
                if (!(cP.Np & 1) && m == half && i >= half) break;
                int j = i + m; if (j >= cP.Np) j -= cP.Np;
                double d0 = xi[0] - Rx[pidx(j)], d1 = xi[1] - Ry[pidx(j)], d2 = xi[2] - Rz[pidx(j)];
                if (!PIGS_TRAP) {      // the estimators take the reference's decisions bit for bit (see mimg_cmp, r2_ref)
                    d0 = mimg_cmp(d0, cP.L[0], cP.Lh[0]); d1 = mimg_cmp(d1, cP.L[1], cP.Lh[1]); d2 = mimg_cmp(d2, cP.L[2], cP.Lh[2]);
                }
                double r2 = r2_ref(d0, d1, d2);
                if (PIGS_TRAP || r2 <= cP.rcut2) {
                    Lk k = lk_prep(sqrt_pos(r2));
                    if (PIGS_TRAP) k.i0 = min(k.i0, cP.Nmax - 1);
                    pot += 2.0 * (PIGS_VPAIR ? lk_val_pair(k) : lk_val<PIGS_VSM, 0, PIGS_VSM>(k));
                }
            }
        }
        double potsh = one + 0.5 * pot;                 // this particle's share of Pot(slice)
        double w = (ib == 0) ? cP.wE[2] : cP.wE[ib & 1];
        double e = w * potsh;
        if (odd) e += cP.cFE * (F[0] * F[0] + F[1] * F[1] + F[2] * F[2]);
        if (ib == cP.Nb) s[1] += potsh;
        // kinetic link ib -> ib+1 (sample_mod.f90:359-380)
        const double* Nx = slice(gs, ib + 1);
        double l0 = xi[0] - Nx[pidx(i)], l1 = xi[1] - Nx[pidx(i) + PY], l2 = xi[2] - Nx[pidx(i) + PZ];
        if (!PIGS_TRAP) {
            l0 = mimg(l0, cP.L[0], cP.Lh[0]); l1 = mimg(l1, cP.L[1], cP.Lh[1]); l2 = mimg(l2, cP.L[2], cP.Lh[2]);
        }
        double lr2 = l0 * l0 + l1 * l1 + l2 * l2;
        if (PIGS_TRAP || lr2 <= cP.rcut2) e -= lr2 * cP.half_inv_dt2;
        s[0] += e;
    }
    group_sum<2>(gs, s);
    double E = 0.5 * (s[0] / (double)cP.Nb + (double)(cP.dim * cP.Np) / dt);
    out[0] = E; out[1] = E - s[1]; out[2] = s[1];
}

// PairCorrelation (sample_mod.f90:392-431): gr(bin) += 2 per pair.  Counts are
// small integers, so atomic accumulation in any order is exact.
static __device__ __noinline__ void PairCorrelation(const GS* gs, const double* Rx, double* gr) {
    const Grp G = grp(gs);
    const double* Ry = Rx + PY;
    const double* Rz = Rx + PZ;
    const int Np = cP.Np, half = Np / 2;
    // pair (i, (i+m) mod Np), m = 1..half; for even Np the m == half pairs are taken from i < half only
    for (int it = G.tid; it < Np * half; it += G.size) {
        int i = it / half, m = it - i * half + 1;
        if (!(Np & 1) && m == half && i >= half) continue;
        int j = i + m; if (j >= Np) j -= Np;
        double d0 = mimg(Rx[pidx(i)] - Rx[pidx(j)], cP.L[0], cP.Lh[0]);
        double d1 = mimg(Ry[pidx(i)] - Ry[pidx(j)], cP.L[1], cP.Lh[1]);
        double d2 = mimg(Rz[pidx(i)] - Rz[pidx(j)], cP.L[2], cP.Lh[2]);
        // no FMA contraction: keeps the bin index bit-identical to the reference's arithmetic
        double r2 = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
        if (r2 <= cP.rcut2) {
            int ibin = (int)(sqrt(r2) / cP.rbin);
            if (ibin < cP.Nbin) atomicAdd(gr + ibin, 2.0);
        }
    }
}
// StructureFactor (sample_mod.f90:435-476): thread <-> (iq,k)
static __device__ __noinline__ void StructureFactor(const GS* gs, const double* Rx, double* Sk) {
    const Grp G = grp(gs);
    for (int it = G.tid; it < cP.Nk * cP.dim; it += G.size) {
        int iq = it / cP.dim + 1, k = it - (iq - 1) * cP.dim;
        const double* X = Rx + 32 * k;
        double q = (double)iq * cP.qbin[k], sc = 0.0, ss = 0.0;
        for (int ip = 0; ip < cP.Np; ++ip) {
            double s, c;
            sincos(q * X[pidx(ip)], &s, &c);
            sc += c; ss += s;
        }
        Sk[it] += sc * sc + ss * ss;          // Sk(k,iq) column-major == [iq][k]
    }
}
// OBDM (sample_mod.f90:480-526)
static __device__ __noinline__ void OBDM(const GS* gs, const double* xend, double* nrho) {
    if (gfirst(gs)) {
        double d[3] = {0, 0, 0}, r2 = 0.0;
#pragma unroll
        for (int k = 0; k < 3; ++k) if (k < cP.dim) {
            d[k] = mimg(xend[k] - xend[3 + k], cP.L[k], cP.Lh[k]);
            r2 = __dadd_rn(r2, __dmul_rn(d[k], d[k]));
        }
        if (r2 <= cP.rcut2) {
            double r = sqrt(r2);
            int ibin = (int)(r / cP.rbin);
            if (ibin < cP.Nbin) {
                double ct = d[0] / r, st = (cP.dim >= 2 ? d[1] : 0.0) / r;
                double e2r = ct * ct - st * st, e2i = 2.0 * ct * st, mr = 1.0, mi = 0.0;
                for (int m = 0; m <= cP.Npw; ++m) {
                    nrho[ibin * (cP.Npw + 1) + m] += mr;
                    double nr = mr * e2r - mi * e2i, ni = mr * e2i + mi * e2r;
                    mr = nr; mi = ni;
                }
            }
        }
    }
}

// ---------------------------------------------------------------- driver schedule
// The diagonal sweep as ONE loop over its moves with ONE call site of the
// engine (see run_move_body): translate every chain (every CMFreq steps), then
// Nstag passes of {head, tail, middle} per particle.  Same order, same draws and
// same counters as the nested loops of the driver; the wrappers of the section
// above (TranslateChain, MoveHead, ...) restate the same descriptors for the
// unit calls.
PIGS_T __device__ __forceinline__ void diag_sweep(GS* gs, ull* pctr, int istep, int skip0) {       // vpi.f90:329-366 / 412-439
    const int Np = cP.Np, twoNb = 2 * cP.Nb;
    const bool bis = cP.sampling != 0;
    const int nT = (istep % cP.CMFreq == 0) ? Np : 0;
    const int total = nT + cP.Nstag * Np * 3;
    int sub = nT ? 0 : 1, ip = 0;            // sub: 0 translate, 1 head, 2 tail, 3 middle
    for (int q = 0; q < total; ++q) {
        const int cs = sub, cip = ip;
        if (sub == 0) { if (++ip == Np) { ip = 0; sub = 1; } }
        else if (sub == 3) { sub = 1; if (++ip == Np) ip = 0; }
        else ++sub;
        if (cip == skip0) continue;
        int flags, ii, ie, m0, m1, cacc;
        if (cs == 0) {                        // TranslateChain
            bump(gs, C_TRY_CM);
            flags = MV_TRANSLATE; ii = 0; ie = twoNb; m0 = 0; m1 = twoNb; cacc = C_ACC_CM;
        } else {
            if (cs == 1) bump(gs, C_TRY_STAG);
            const int Lm = bis ? (1 << cP.Nlev) : cP.Lstag;
            ull ctr = *pctr;
            const int d = draw_int<MT, VAR>(gs, ctr, (cs == 3) ? twoNb - Lm + 1 : (bis ? cP.Nlev : cP.Lstag) - 1);
            *pctr = ctr;
            const int Ls = bis ? (1 << (d + 2)) : d + 2;      // length of an end move
            const int ty = bis ? MV_BISECT : MV_BRIDGE;
            if (cs == 1) {                    // MoveHead / MoveHeadBisection
                flags = ty | MV_FREE_NEXT; ii = 0; ie = Ls; m0 = 0; m1 = ie - 1; cacc = C_ACC_HEAD;
            } else if (cs == 2) {             // MoveTail / MoveTailBisection
                flags = ty | MV_FREE_PREV; ie = twoNb; ii = ie - Ls; m0 = ii + 1; m1 = ie; cacc = C_ACC_TAIL;
            } else {                          // Staging / Bisection
                flags = ty; ii = d; ie = ii + Lm; m0 = ii + 1; m1 = ie - 1; cacc = C_ACC_BD;
            }
        }
        if (run_move_body<MT, VAR>(gs, pctr, flags, cip, ii, ie, m0, m1, 0.0)) bump(gs, cacc);
    }
}
// The diagonal sweep of the PRODUCTION (Philox) mode: same moves, same counts per Monte-Carlo step as vpi.f90:412-439
// (translate every particle, then Nstag passes in which every particle gets a head, a tail and a middle move), but
// ordered by TIME-SLICE WINDOW instead of by particle: a pass draws its windows once (head length, tail length,
// start of the middle segment -- the reference draws them per particle) and moves all Np particles on one window
// before going to the next.  Every move is the reference's move and satisfies detailed balance on its own, and the
// window is drawn independently of the configuration, so the stationary distribution is unchanged (checked
// statistically against the oracle); MT19937 replay keeps the reference's order (diag_sweep).
// Why: the Np moves of a window re-read the same <= 2^Nlev + 1 slices, which then come from L2 instead of HBM
// (DRAM traffic per step falls to that of the translations), and -- team mode, cA.team -- the four warps of a
// chain group can sweep four DISJOINT windows of the same chain at the same time (head | middle | middle | tail):
// moves of different particles on disjoint slice sets commute exactly.  The only shared data of two concurrent
// moves would be the anchor beads of one particle; the workers therefore walk the particles in orders rotated by
// Np/4 and meet at a barrier every Np/8 moves, so no two of them ever hold the same particle.  (A team pass makes
// two middle sweeps where the reference makes one; the middle windows are drawn inside the range the head and
// tail windows leave free.)  Translations need the action of all slices: the team evaluates them together.
struct WinPass {
    int Lh, Lt;          // head / tail segment lengths of this pass
    int iiM[2];          // start beads of the middle segments
    int nM;              // how many of them are in use
};
template <int VAR>
__device__ __forceinline__ WinPass draw_pass(GS* gs, ull& ctr, bool team) {
    const bool bis = cP.sampling != 0;
    const int twoNb = 2 * cP.Nb, Lm = bis ? (1 << cP.Nlev) : cP.Lstag;
    const int nend = (bis ? cP.Nlev : cP.Lstag) - 1;
    WinPass w;
    const int dh = draw_int<false, VAR>(gs, ctr, nend), dt = draw_int<false, VAR>(gs, ctr, nend);
    w.Lh = bis ? (1 << (dh + 2)) : dh + 2;
    w.Lt = bis ? (1 << (dt + 2)) : dt + 2;
    if (!team) {
        w.iiM[0] = draw_int<false, VAR>(gs, ctr, twoNb - Lm + 1);
        w.iiM[1] = 0;
        w.nM = 1;
    } else {
        // interiors of the middle segments inside [Lh, 2Nb - Lt], Lm - 1 beads each, at a random offset
        const int lo = w.Lh, span = twoNb - w.Lt - lo + 1;
        int n = (span > 0) ? span / (Lm - 1) : 0;
        if (n > 2) n = 2;
        w.nM = n;
        w.iiM[0] = w.iiM[1] = 0;
        if (n > 0) {
            const int off = draw_int<false, VAR>(gs, ctr, span - n * (Lm - 1) + 1);
            w.iiM[0] = lo - 1 + off;
            w.iiM[1] = w.iiM[0] + (Lm - 1);
        }
    }
    return w;
}
template <int VAR>
__device__ __forceinline__ void win_sweep(GS* gs0, GS* gsw, ull* pctr, ull* pwctr, int istep, int skip0) {
    const int Np = cP.Np, twoNb = 2 * cP.Nb;
    const bool team = cA.team != 0;
    const int wk = team ? (int)((threadIdx.x & (cA.threads_per_chain - 1)) >> 5) : 0;
    const bool bis = cP.sampling != 0;
    const int Lm = bis ? (1 << cP.Nlev) : cP.Lstag;
    const int ty = bis ? MV_BISECT : MV_BRIDGE;
    const int nT = (istep % cP.CMFreq == 0) ? Np : 0;
    const int nsub = team ? 1 : 3;                              // windows this thread walks per pass
    const int chunk = team ? (Np >= 8 ? (Np >> 3) : 1) : Np;     // team barrier every `chunk` moves
    const int rot = team ? wk * (Np >> 2) : 0;                   // rotated particle order of worker wk
    const int total = nT + cP.Nstag * nsub * Np;
    int sub = 0, k = 0, left = 0;       // position inside the pass; moves left until the next team barrier
    for (int q = 0; q < total; ++q) {
        GS* g;
        ull* rs;
        int flags, ip, ii, ie, m0, m1, cacc;
        if (q < nT) {                   // TranslateChain, all threads of the chain group together
            ip = q;
            if (ip == skip0) continue;
            g = gs0; rs = pctr;
            bump(g, C_TRY_CM);
            flags = MV_TRANSLATE; ii = 0; ie = twoNb; m0 = 0; m1 = twoNb; cacc = C_ACC_CM;
        } else {
            if (sub == 0 && k == 0) {   // a new pass: its windows, from the chain's stream (identical in every thread)
                if (team) {
                    tsync();
                    if (q == nT) {      // first pass: gs0 becomes the block of window worker 0
                        if ((threadIdx.x & (cA.threads_per_chain - 1)) == 0) { gs0->gsize = 32; gs0->gshift = 5; }
                        tsync();
                    }
                }
                ull c = *pctr;
                const WinPass wp = draw_pass<VAR>(gsw, c, team);
                *pctr = c;
                // parked in the worker's block (every thread writes the same values and reads back its own): nothing of
                // the pass stays in registers across the move engine
                gsw->wLh = wp.Lh; gsw->wLt = wp.Lt; gsw->wM0 = wp.iiM[0]; gsw->wM1 = wp.iiM[1]; gsw->wnM = wp.nM;
                left = 0;
            }
            if (team && left == 0) { tsync(); left = chunk; }
            left -= 1;
            const int kk = k;
            const int cs = team ? wk : sub;             // team: 0 head, 1|2 middle, 3 tail;  else 0 head, 1 tail, 2 middle
            if (++k == Np) { k = 0; if (++sub == nsub) sub = 0; }
            ip = kk + rot; if (ip >= Np) ip -= Np;
            if (ip == skip0) continue;
            g = gsw; rs = team ? pwctr : pctr;
            const bool head = cs == 0, tail = team ? cs == 3 : cs == 1;
            if (head) {                 // MoveHead / MoveHeadBisection
                bump(g, C_TRY_STAG);
                flags = ty | MV_FREE_NEXT; ii = 0; ie = g->wLh; m0 = 0; m1 = ie - 1; cacc = C_ACC_HEAD;
            } else if (tail) {          // MoveTail / MoveTailBisection
                flags = ty | MV_FREE_PREV; ie = twoNb; ii = ie - g->wLt; m0 = ii + 1; m1 = ie; cacc = C_ACC_TAIL;
            } else {                    // Staging / Bisection
                const int j = team ? cs - 1 : 0;
                if (j >= g->wnM) continue;
                flags = ty; ii = j ? g->wM1 : g->wM0; ie = ii + Lm; m0 = ii + 1; m1 = ie - 1; cacc = C_ACC_BD;
            }
        }
        if (run_move_body<false, VAR>(g, rs, flags, ip, ii, ie, m0, m1, 0.0)) bump(g, cacc);
    }
    if (team) {                         // back to the chain group's own geometry
        tsync();
        if ((threadIdx.x & (cA.threads_per_chain - 1)) == 0) { gs0->gsize = cA.threads_per_chain; gs0->gshift = cA.tshift; }
        tsync();
    }
}
PIGS_T __device__ __forceinline__ void mc_step(GS* gs, GS* gsw, ull* pctr, ull* pwctr, int istep) {      // vpi.f90:297-475
    const Grp G = grp(gs);
    ull ctr = *pctr;
    int iupdate = (int)(uniform<MT, VAR>(gs, ctr) * 2.0);
    if (gs->isopen) {
        if (iupdate == 0) {
            *pctr = ctr;
            OpenClose<MT, VAR>(gs, pctr, cP.Lstag, gs->iworm0, false);
            ctr = *pctr;
            bump(gs, C_TRY_CLOSE);
            PermutationSampling(gs, false);
        }
    } else {
        if (iupdate == 1) {
            int iw = draw_int<MT, VAR>(gs, ctr, cP.Np);
            gsync(gs);
            if (G.tid == 0) gs->iworm0 = iw;
            gsync(gs);
            *pctr = ctr;
            OpenClose<MT, VAR>(gs, pctr, cP.Lstag, iw, true);
            ctr = *pctr;
            bump(gs, C_TRY_OPEN);
            PermutationSampling(gs, false);
        }
    }
    *pctr = ctr;
    gsync(gs);
    const bool is_open = gs->isopen;
    if (!is_open) {
        if (G.tid == 0) gs->idiag_aux += 1;
        bump(gs, C_IDIAG);
    }
    // the one inlined instance of the engine: reference order under MT19937 replay, window order in production
    if (MT) diag_sweep<MT, VAR>(gs, pctr, istep, is_open ? gs->iworm0 : -1);
    else win_sweep<VAR>(gs, gsw, pctr, pwctr, istep, is_open ? gs->iworm0 : -1);
    if (is_open) {
        const int iw = gs->iworm0;
        for (int iobdm = 0; iobdm < cP.Nobdm; ++iobdm) {
            for (int j = 1; j <= 2; ++j) {
                bump(gs, C_TRY_CM_HALF);
                TranslateHalfChain<MT, VAR>(gs, pctr, j, iw);
            }
            for (int j = 1; j <= 2; ++j) {
                bump(gs, C_TRY_STAG_HALF);
                MoveHead<MT, VAR>(gs, pctr, cP.Lstag, iw, j);
                MoveTail<MT, VAR>(gs, pctr, cP.Lstag, iw, j);
                StagingHalfChain<MT, VAR>(gs, pctr, j, cP.Lstag, iw);
            }
            if (cP.swapping) {
                bump(gs, C_TRY_SWAP);
                Swap<MT, VAR>(gs, pctr, cP.Lstag, iw);
                PermutationSampling(gs, true);
            }
            if (!PIGS_TRAP) { OBDM(gs, gs->xend, gs->acc + cP.off_nr); gsync(gs); }
        }
    } else {
        double e1[3], e2[3], et[3];
        LocalEnergy<VAR>(gs, slice(gs, 0), e1);
        LocalEnergy<VAR>(gs, slice(gs, 2 * cP.Nb), e2);
        double E = 0.5 * (e1[0] + e2[0]);
        ThermEnergy<VAR>(gs, et);
        double Pot = et[2], Kin = E - Pot, Et = et[0], Kt = et[1];
        if (G.tid == 0) {
            double* eacc = gs->eacc;
            eacc[0] += E; eacc[1] += Kin; eacc[2] += Pot; eacc[3] += Et; eacc[4] += Kt; eacc[5] += Pot;
            eacc[6] += E * E; eacc[7] += Kin * Kin; eacc[8] += Pot * Pot;
            eacc[9] += Et * Et; eacc[10] += Kt * Kt; eacc[11] += Pot * Pot;
        }
        bump(gs, C_NGR);
        if (!PIGS_TRAP) {
            PairCorrelation(gs, slice(gs, cP.Nb), gs->acc + cP.off_gr);
            StructureFactor(gs, slice(gs, cP.Nb), gs->acc + cP.off_sk);
        }
    }
}

PIGS_T __device__ __forceinline__ void do_move(GS* gs, ull* pctr, int move, int ip0, int half) {
    switch (move) {
    case 0: TranslateChain<MT, VAR>(gs, pctr, ip0); break;
    case 1: Staging<MT, VAR>(gs, pctr, cP.Lstag, ip0); break;
    case 2: MoveHead<MT, VAR>(gs, pctr, cP.Lstag, ip0, 0); break;
    case 3: MoveTail<MT, VAR>(gs, pctr, cP.Lstag, ip0, 0); break;
    case 4: Bisection<MT, VAR>(gs, pctr, cP.Nlev, ip0); break;
    case 5: EndBisection<MT, VAR>(gs, pctr, cP.Nlev, ip0, true); break;
    case 6: EndBisection<MT, VAR>(gs, pctr, cP.Nlev, ip0, false); break;
    case 7: TranslateHalfChain<MT, VAR>(gs, pctr, half, ip0); break;
    case 8: StagingHalfChain<MT, VAR>(gs, pctr, half, cP.Lstag, ip0); break;
    case 9: MoveHead<MT, VAR>(gs, pctr, cP.Lstag, ip0, half); break;
    case 10: MoveTail<MT, VAR>(gs, pctr, cP.Lstag, ip0, half); break;
    case 11:
        gsync(gs);
        if (gfirst(gs)) gs->iworm0 = ip0;
        gsync(gs);
        OpenClose<MT, VAR>(gs, pctr, cP.Lstag, ip0, true);
        break;
    case 12: OpenClose<MT, VAR>(gs, pctr, cP.Lstag, ip0, false); break;
    case 13: Swap<MT, VAR>(gs, pctr, cP.Lstag, ip0); break;
    default: break;
    }
}

PIGS_T __device__ __forceinline__ void sweep_body() {
    extern __shared__ __align__(16) double smem[];
    using VT = VarTraits<VAR>;
    const int T = cA.threads_per_chain, Gn = cA.groups_per_cta;
    const int ntab = tab_len(cP.Nmax);
    double* sp = smem;
    const double* tV = cP.vtab;
    const double* tW = cP.logwf;
    if (VT::VPAIR) {
        // pair table: entry i = {F(i), F(i+1)}, i = 0..Nmax  (2*(Nmax+1) doubles, 16-byte aligned)
        for (int i = threadIdx.x; i < ntab - 1; i += blockDim.x) { sp[2 * i] = cP.vtab[i]; sp[2 * i + 1] = cP.vtab[i + 1]; }
        tV = cP.vtab; sp += 2 * (ntab - 1);
    } else if (VT::VSM) {
        tV = sp; sp += ntab;
    }
    if (VT::WSM) { tW = sp; sp += ntab; }
    if (!VT::VPAIR && (VT::VSM || VT::WSM)) {
        // Table staging by the TMA engine: 1-D bulk copies global -> shared (cp.async.bulk, UBLKCP in SASS) issued by
        // one thread and completed on an mbarrier -- 160 KB per CTA in a few microseconds instead of 40 rounds of
        // LDG + STS by every thread (ncu, round 2: the copy loop held 3 % of the samples of a 2-step launch).
        __shared__ __align__(8) unsigned long long stage_bar;
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&stage_bar);
        const unsigned nbytes = (unsigned)ntab * 8u;
        const bool bulk = (nbytes & 15u) == 0;               // bulk copies move multiples of 16 bytes
        unsigned done = 0;
        if (bulk) {
            if (threadIdx.x == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const unsigned total = nbytes * ((VT::VSM ? 1u : 0u) + (VT::WSM ? 1u : 0u));
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(total) : "memory");
                for (int which = 0; which < 2; ++which) {
                    if (which == 0 ? !VT::VSM : !VT::WSM) continue;
                    const char* src = reinterpret_cast<const char*>(which == 0 ? cP.vtab : cP.logwf);
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(which == 0 ? tV : tW);
                    for (unsigned off = 0; off < nbytes; off += 32768u) {
                        const unsigned len = nbytes - off < 32768u ? nbytes - off : 32768u;
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(dst + off), "l"(src + off), "r"(len), "r"(bar) : "memory");
                    }
                }
            }
            for (int spin = 0; !done && spin < (1 << 22); ++spin)      // bounded: a copy that never lands must not hang the GPU
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(bar) : "memory");
        }
        if (!__syncthreads_and((int)done)) {                 // odd table length (or a copy that did not complete): plain loop
            if (VT::VSM) for (int i = threadIdx.x; i < ntab; i += blockDim.x) const_cast<double*>(tV)[i] = cP.vtab[i];
            if (VT::WSM) for (int i = threadIdx.x; i < ntab; i += blockDim.x) const_cast<double*>(tW)[i] = cP.logwf[i];
        }
    }
    __syncthreads();           // last CTA-wide barrier: groups run independently from here on

    const int g = threadIdx.x >> cA.tshift;
    const int tid = threadIdx.x & (T - 1);
    if (g >= Gn) return;
    // team mode: four blocks per chain group, one per window worker (= warp); block 0 is also the group's own
    const bool team = cA.team != 0;
    const int wk = team ? (tid >> 5) : 0;
    const size_t gbytes = grp_smem_bytes(cP.S, cP.Np, T >> 5);
    GS* gs = reinterpret_cast<GS*>(reinterpret_cast<char*>(sp) + (size_t)g * (team ? 4 : 1) * gbytes);
    GS* gsw = reinterpret_cast<GS*>(reinterpret_cast<char*>(gs) + (size_t)wk * gbytes);

    for (int c = blockIdx.x * Gn + g; c < cP.n_chains; c += gridDim.x * Gn) {
        if (cA.chain_only >= 0 && c != cA.chain_only) continue;
        int* ist = cP.istate + (size_t)c * IS_N;
        if ((tid & 31) == 0 && (tid == 0 || team)) {       // thread 0 fills the group's block, lane 0 of a worker warp its own
            GS* b = gsw;
            b->path = cP.path + (size_t)c * cP.chain_stride;
            b->xend = cP.xend + (size_t)c * 6;
            b->acc = cP.acc + (size_t)c * cP.nacc;
            b->cyc = cP.cyc + (size_t)c * cP.Np;
            b->hist = cP.hist + (size_t)c * cP.Np;
            b->mt = cP.mt + (size_t)c * 624;
            b->pp = cP.pp + (size_t)c * cP.Np;
            b->tabV = tV; b->tabW = tW;
            b->chain = c + cP.chain_offset;
            b->kc[0] = cP.invL[0]; b->kc[1] = cP.invL[1]; b->kc[2] = cP.invL[2]; b->kc[3] = cP.inv_dr;
            b->gsize = tid == 0 ? T : 32; b->gshift = tid == 0 ? cA.tshift : 5; b->gbar = g;
        }
        if (tid == 0) {
            gs->mti = ist[IS_MTI]; gs->isopen = ist[IS_OPEN]; gs->iworm0 = ist[IS_IWORM] - 1; gs->iperm = ist[IS_IPERM];
            gs->new_pc = ist[IS_NEWPC]; gs->end_pc = ist[IS_ENDPC]; gs->ik0 = ist[IS_IK] - 1; gs->swap_acc = 0;
            gs->idiag_aux = ist[IS_IDIAG_AUX];
        }
        if (tid < NE) gs->eacc[tid] = 0.0;
        if ((tid & 31) < NCNT && (tid < 32 || team)) gsw->cnt[tid & 31] = 0;
        ull ctr, wctr;
        ctr.ctr = cP.pctr[(size_t)c * PCS]; ctr.w0 = ctr.w1 = ctr.w2 = 0u; ctr.nleft = 0; ctr.tag = 0u; ctr.pad = 0u;
        wctr = ctr;
        if (team) { wctr.ctr = cP.pctr[(size_t)c * PCS + 1 + wk]; wctr.tag = 1u + (unsigned)wk; }
        tsync();

        if (cA.op == OP_BLOCK) {
            for (int istep = 1; istep <= cA.nstep; ++istep) { mc_step<MT, VAR>(gs, gsw, &ctr, &wctr, istep); gsync(gs); }
            if (tid < NE) gs->acc[tid] = gs->eacc[tid];
            if (tid < NCNT) {
                long long v = gs->cnt[tid];
                if (team)
                    for (int w = 1; w < 4; ++w) v += reinterpret_cast<GS*>(reinterpret_cast<char*>(gs) + (size_t)w * gbytes)->cnt[tid];
                if (tid == C_NOPEN) v = gs->isopen;
                cP.cnt[(size_t)c * NCNT + tid] = v;
            }
        } else if (cA.op == OP_MOVE) {
            do_move<MT, VAR>(gs, &ctr, cA.move, cA.ip0, cA.half);
            gsync(gs);
            if (tid == 0) {
                long long nacc = 0;
                for (int i = C_ACC_CM; i <= C_ACC_SWAP; ++i)
                    if (i != C_TRY_OPEN && i != C_TRY_CLOSE && i != C_TRY_SWAP) nacc += gs->cnt[i];
                if (cA.accepted) cA.accepted[c] = (int)nacc;
                if (cA.aux) cA.aux[c] = (cA.move == 13 && gs->swap_acc) ? gs->ik0 + 1 : 0;
                for (int i = 0; i < 3; ++i) cP.cnt[(size_t)c * NCNT + C_UPD_EVEN + i] = gs->cnt[C_UPD_EVEN + i];
            }
        } else if (cA.op == OP_UNIFORM) {
            for (int i = 0; i < cA.nstep; ++i) {
                double u = rng_uniform<MT>(gs, ctr);
                if (tid == 0) cA.draws[i] = u;
            }
        } else if (cA.op == OP_GAUSS) {
            // one Gaussian per call, through the same fill routine the moves use
            for (int i = 0; i < cA.nstep; ++i) {
                rng_gauss_fill<MT>(gs, &ctr, 1, 0, 1, 1);
                if (tid == 0) cA.draws[i] = seg_new(gs)[0];
                gsync(gs);
            }
        } else if (cA.op == OP_SEED) {
            if (tid == 0) mt_seed(gs->mt, gs->mti, (unsigned)(cA.seed + (cA.chain_only >= 0 ? 0 : c + cP.chain_offset)));
        }
        gsync(gs);
        if (tid == 0) {
            ist[IS_OPEN] = gs->isopen; ist[IS_IWORM] = gs->iworm0 + 1; ist[IS_IPERM] = gs->iperm;
            ist[IS_NEWPC] = gs->new_pc; ist[IS_ENDPC] = gs->end_pc; ist[IS_IK] = gs->ik0 + 1; ist[IS_IDIAG_AUX] = gs->idiag_aux;
            ist[IS_MTI] = gs->mti;
            cP.pctr[(size_t)c * PCS] = ctr.ctr;
        }
        if (team && (tid & 31) == 0) cP.pctr[(size_t)c * PCS + 1 + wk] = wctr.ctr;
        gsync(gs);
    }
}

}  // namespace pigs
