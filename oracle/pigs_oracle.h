/*
 * pigs_oracle.h -- C interface of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a serial, line-faithful C++ restatement
 * of the Fortran 90 reference (amaciarey/PathIntegralGroundState).  It exists
 * to check the CUDA product path (libpigs_cuda) and to serve as the reported
 * CPU baseline in bench.py.  Nothing in pathintegralgroundstate_b200/ may
 * import, link or call it: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do.
 *
 * PARITY PINNING: the reference ships no tests, golden vectors or fixtures and no Fortran compiler exists in
 * this image, so the reference BINARY cannot be run.  The oracle is pinned instead by oracle/_ref: the
 * reference's own Fortran sources, machine-translated statement by statement into C++ by the committed recipe
 * oracle/f90toc/f90toc.py (read from /root/reference where they lie; the generated file is git-ignored) and
 * compiled with the oracle's flags.  tests/test_ref_pin.py demands BIT EQUALITY between that translation and this
 * oracle: leaf functions, tables, the MT19937 stream (the translated generator reproduces the published 1998
 * head 3510405877, 4290933890), UpdateAction, the estimators, each of the 14 moves from identical state and
 * stream, and the complete `./vpi < vpi.in` program block by block (worm open/close/swap all accepted);
 * tests/golden/ref_golden.json holds vectors taken from the translation for boxes without the reference.
 * (The pin found one discrepancy in round 2 -- (Rm/r)**5 is (x*x^2)*x^2 in gfortran's expansion, not a left-to-right
 * product: 1 ulp in a third of the Jastrow table entries -- now fixed here and in the product's table generators.)
 * What the pin cannot see: gfortran's own code generation beyond IEEE semantics and the documented expansion of
 * integer powers (GCC's powi table), and list-directed output formatting.
 *
 * All indices crossing this interface are Fortran-style: particles ip=1..Np,
 * beads ib=0..2*Nb, arrays column-major Path(dim,Np,0:2*Nb), xend(dim,2),
 * tables F(0:Nmax+1).
 */
#ifndef PIGS_ORACLE_H
#define PIGS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Inputs of vpi.in (namelists &system &samp &obdm &wavefun &jastrow &extpot;
 * vpi_mod.f90:28-32, system_mod.f90:21-22) as the driver reads them. */
typedef struct orc_params {
    int32_t dim, Np;
    double  density;
    int32_t crystal, trap;               /* logicals */
    double  dt;
    int32_t Nb, seed;
    double  delta_cm;                    /* UNscaled, as in vpi.in */
    int32_t CMFreq;
    int32_t sampling;                    /* 0 = 'sta', 1 = 'bis' */
    int32_t Lstag, Nlev, Nstag, Nbin, Nk;
    int32_t swapping;
    double  CWorm;
    int32_t Nobdm, Npw;
    int32_t Nmax, wf_table, v_table;
    double  Rm;
    double  a_ho[3];
    double  Lbox_crystal[3];             /* used only when crystal != 0 (config_ini.in line 2) */
    int32_t action;                      /* 0 = Chin (live code), 1 = primitive (the commented-out lines
                                            global_mod.f90:48,67: GreenFunction = dt*Pot / Pot) */
    int32_t pad_;
} orc_params;

/* Raw block sums exactly as the driver holds them at the end of the step
 * loop (vpi.f90:297-475), before NormalizeAv. */
typedef struct orc_block {
    double  sumE, sumK, sumV, sumEt, sumKt, sumVt;
    double  sumE2, sumK2, sumV2, sumEt2, sumKt2, sumVt2;
    int32_t idiag_block, ngr;
    /* try/acc counters, vpi.f90:250-273 */
    double  try_cm, try_stag, try_cm_half, try_stag_half;
    int32_t acc_cm, acc_bd, acc_head, acc_tail;
    int32_t acc_cm_half, acc_bd_half, acc_head_half, acc_tail_half;
    int32_t try_open, acc_open, try_close, acc_close, try_swap, acc_swap;
    int32_t idiag_aux;                   /* running (not reset per block; vpi.f90:236,409,536) */
    int32_t pad_;
    uint64_t bead_updates[3];            /* UpdateAction calls: [0] interior even, [1] odd, [2] end slices */
} orc_block;

typedef struct orc_sim orc_sim;

/* geometry (vpi.f90:80-128), allocation; tables are NOT filled, state NOT initialised */
orc_sim* orc_create(const orc_params* p);
void     orc_destroy(orc_sim* s);

/* derived globals */
void   orc_get_geometry(const orc_sim* s, double* Lbox3, double* rcut, double* dr, double* rbin,
                        double* density, double* delta_cm_scaled);

/* JastrowTable / PotentialTable (vpi_mod.f90:84-145); writes no files */
void   orc_fill_tables(orc_sim* s);
void   orc_set_tables(orc_sim* s, const double* LogWF, const double* VTable);   /* (0:Nmax+1) each */
void   orc_get_tables(const orc_sim* s, double* LogWF, double* VTable);

/* init (vpi_mod.f90:149-259), non-resume branch; crystal positions come from R0(dim,Np) */
void   orc_init(orc_sim* s, const double* R0_or_null);
void   orc_set_state(orc_sim* s, const double* Path, const double* xend, int isopen, int iworm);
void   orc_get_state(const orc_sim* s, double* Path, double* xend, int* isopen, int* iworm);
void   orc_set_perm(orc_sim* s, int iperm, const int32_t* cyc, const int32_t* hist, int new_pc, int end_pc);
void   orc_get_perm(const orc_sim* s, int* iperm, int32_t* cyc, int32_t* hist, int* new_pc, int* end_pc);

/* random_mod.f90 */
void   orc_sgrnd(orc_sim* s, int32_t seed);
double orc_grnd(orc_sim* s);
double orc_rangauss(orc_sim* s);                      /* first deviate of rangauss(1,0,..) */
void   orc_get_mt(const orc_sim* s, uint32_t* mt624, int32_t* mti);
void   orc_set_mt(orc_sim* s, const uint32_t* mt624, int32_t mti);
uint32_t orc_mt_raw(orc_sim* s);                      /* tempered 32-bit output */

/* primitives (stateless where possible) */
double orc_interpolate(int opt, int N, double dx, const double* F, double x);   /* interpolate.f90 */
double orc_potential(double rij);                                               /* system_mod.f90:136-182 */
double orc_logpsi(int opt, double Rm, double rij);                              /* system_mod.f90:38-66 */
double orc_green_function(const orc_sim* s, int opt, int ib, double dt, double Pot, double F2);
void   orc_minimum_image(const orc_sim* s, double* xij, double* rij2);
double orc_boundary_conditions(const orc_sim* s, int k, double x);

/* UpdateAction on the simulation's own Path (vpi_mod.f90:2491-2530); the
 * slice must already hold xnew for ip or not -- irrelevant, jp==ip is skipped */
double orc_update_action(orc_sim* s, int ip, int ib, const double* xnew, const double* xold);
/* the same on an explicit slice R(dim,Np) */
double orc_update_action_R(orc_sim* s, const double* R, int ip, int ib, const double* xnew, const double* xold);

/* moves; return 1 if accepted */
enum {
    ORC_TRANSLATE_CHAIN = 0, ORC_STAGING = 1, ORC_MOVE_HEAD = 2, ORC_MOVE_TAIL = 3,
    ORC_BISECTION = 4, ORC_MOVE_HEAD_BISECTION = 5, ORC_MOVE_TAIL_BISECTION = 6,
    ORC_TRANSLATE_HALF = 7, ORC_STAGING_HALF = 8, ORC_MOVE_HEAD_HALF = 9, ORC_MOVE_TAIL_HALF = 10,
    ORC_OPEN = 11, ORC_CLOSE = 12, ORC_SWAP = 13
};
/* ip: particle (for ORC_SWAP: the worm iw); half: 1|2 for the *_HALF moves.
 * Uses the sim's dt, delta_cm (scaled), Lstag, Nlev.  For ORC_SWAP, *aux
 * receives ipar (partner) when the swap is accepted. */
int    orc_move(orc_sim* s, int move, int ip, int half, int* aux);

/* estimators (sample_mod.f90) */
void   orc_potential_energy(orc_sim* s, const double* R, int want_f2, double* Pot, double* F2);
void   orc_local_energy(orc_sim* s, const double* R, double* E, double* Kin, double* Pot);
void   orc_therm_energy(orc_sim* s, double* E, double* Ec, double* Ep);          /* on the sim's Path */
void   orc_therm_energy_P(orc_sim* s, const double* Path, double* E, double* Ec, double* Ep);
void   orc_pair_correlation(orc_sim* s, const double* R, double* gr);            /* gr(Nbin) += */
void   orc_structure_factor(orc_sim* s, const double* R, double* Sk);            /* Sk(dim,Nk) += */
void   orc_obdm(orc_sim* s, const double* xend, double* nrho);                   /* nrho(0:Npw,Nbin) += */
void   orc_normalize_gr(orc_sim* s, int ngr, double* gr);
void   orc_normalize_sk(orc_sim* s, int ngr, double* Sk);
void   orc_normalize_nr(orc_sim* s, double zconf, double* nrho);
double orc_var(int Nitem, double Sum, double Sum2);

/* one block of the driver's step loop (vpi.f90:250-475).  gr(Nbin), Sk(dim,Nk)
 * are zeroed first (as the driver does per block); nrho(0:Npw,Nbin) is
 * accumulated into (the driver resets it only when an OBDM block closes). */
void   orc_run_block(orc_sim* s, int Nstep, orc_block* out, double* gr, double* Sk, double* nrho);
/* total UpdateAction calls so far, by class */
void   orc_bead_updates(const orc_sim* s, uint64_t* three);

#ifdef __cplusplus
}
#endif
#endif
