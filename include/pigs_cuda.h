/*
 * pigs_cuda.h -- C ABI of libpigs_cuda, the B200 (sm_100a) implementation of
 * the PIGS worldline-update + estimator hot path of
 * amaciarey/PathIntegralGroundState.
 *
 * The reference has no FFI: its boundary is the set of Fortran module
 * procedures the driver vpi.f90 calls.  Every entry point below names the
 * reference procedure (file:line) it replaces.  Calling convention: plain C,
 * by-value scalars, caller-owned host buffers, Fortran array layouts kept
 * as they are in the reference:
 *     Path(dim,Np,0:2*Nb)  column-major  == C array [2*Nb+1][Np][dim]
 *     xend(dim,2)                        == C array [2][dim]
 *     LogWF(0:Nmax+1), VTable(0:Nmax+1)  == C array [Nmax+2]
 *     R(dim,Np)                          == C array [Np][dim]
 *     gr(Nbin), Sk(dim,Nk) == [Nk][dim], nrho(0:Npw,Nbin) == [Nbin][Npw+1]
 * Particle indices ip are 1-based, bead indices ib run 0..2*Nb, as in Fortran.
 * Logicals are int (gfortran default-kind logical is 4 bytes).
 *
 * Every function returns 0 on success and a negative code on failure
 * (PIGS_E_*); pigs_last_error() gives the message.  The library never aborts
 * the host process and has NO CPU fallback: without a CUDA device every
 * compute call fails with PIGS_E_CUDA.
 *
 * One handle = one GPU = n_chains independent Markov chains ("replicas" of the
 * reference program's global state).  A handle is used by one host thread.
 */
#ifndef PIGS_CUDA_H
#define PIGS_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PIGS_OK          0
#define PIGS_E_ARG      -1   /* bad argument / unsupported configuration */
#define PIGS_E_CUDA     -2   /* CUDA runtime error (including: no device) */
#define PIGS_E_STATE    -3   /* call order (tables/state not set) */

/* rng_mode */
#define PIGS_RNG_PHILOX     0   /* production: counter-based Philox4x32-10, Box-Muller */
#define PIGS_RNG_MT_REPLAY  1   /* replay: the reference's MT19937 (random_mod.f90) draw for draw */

/* moves, for pigs_move() -- one per reference procedure */
#define PIGS_TRANSLATE_CHAIN        0   /* TranslateChain      vpi_mod.f90:313  */
#define PIGS_STAGING                1   /* Staging             vpi_mod.f90:480  */
#define PIGS_MOVE_HEAD              2   /* MoveHead            vpi_mod.f90:582  */
#define PIGS_MOVE_TAIL              3   /* MoveTail            vpi_mod.f90:724  */
#define PIGS_BISECTION              4   /* Bisection           vpi_mod.f90:864  */
#define PIGS_MOVE_HEAD_BISECTION    5   /* MoveHeadBisection   vpi_mod.f90:1002 */
#define PIGS_MOVE_TAIL_BISECTION    6   /* MoveTailBisection   vpi_mod.f90:1188 */
#define PIGS_TRANSLATE_HALF         7   /* TranslateHalfChain  vpi_mod.f90:383  */
#define PIGS_STAGING_HALF           8   /* StagingHalfChain    vpi_mod.f90:1376 */
#define PIGS_MOVE_HEAD_HALF         9   /* MoveHeadHalfChain   vpi_mod.f90:1495 */
#define PIGS_MOVE_TAIL_HALF        10   /* MoveTailHalfChain   vpi_mod.f90:1660 */
#define PIGS_OPEN                  11   /* OpenChain           vpi_mod.f90:1821 */
#define PIGS_CLOSE                 12   /* CloseChain          vpi_mod.f90:2080 */
#define PIGS_SWAP                  13   /* Swap                vpi_mod.f90:2270 */

/* All the "hidden inputs" the reference procedures read from module globals
 * (global_mod.f90:5-12, system_mod.f90:8-9) and driver locals (vpi.f90:11-56),
 * made explicit.  Geometry is passed already derived, exactly as the driver
 * derives it (vpi.f90:80-128): the caller (Fortran driver or host.py) keeps
 * that arithmetic so its float32 casts stay the reference's. */
typedef struct pigs_params {
    int32_t dim, Np, Nb;               /* Path(dim,Np,0:2*Nb) */
    int32_t Nmax;                      /* tables are (0:Nmax+1) */
    int32_t Nbin, Nk, Npw;             /* gr(Nbin), Sk(dim,Nk), nrho(0:Npw,Nbin) */
    int32_t trap;                      /* logical */
    double  Lbox[3];                   /* ignored when trap */
    double  a_ho[3];                   /* used when trap */
    double  rcut;                      /* vpi.f90:92,122 */
    double  dr;                        /* table step, rcut/real(Nmax-1), vpi_mod.f90:94 */
    double  density;                   /* vpi.f90:90,105 or input */
    double  dt;
    double  delta_cm;                  /* ALREADY scaled (vpi.f90:93,123) */
    double  CWorm;
    int32_t CMFreq;
    int32_t sampling;                  /* 0 = 'sta', 1 = 'bis' */
    int32_t Lstag, Nlev, Nstag, Nobdm;
    int32_t swapping;                  /* logical */
    /* --- what the reference does not have --- */
    int32_t n_chains;                  /* independent Markov chains on this GPU */
    int32_t rng_mode;                  /* PIGS_RNG_* */
    uint64_t seed;                     /* Philox key (chain index is mixed in); MT: sgrnd(seed+chain) */
    int32_t device;                    /* CUDA device ordinal */
    int32_t threads_per_chain;         /* 0 = auto; 32,64,128,256,512 */
    int32_t table_mode;                /* -1 = auto; 0 tables via L1/L2; 1 VTable in smem; 2 both in smem */
    int32_t action;                    /* 0 = Chin (the live propagator, global_mod.f90:33-46), 1 = primitive
                                          (the commented-out alternative, global_mod.f90:48,67) */
    int32_t schedule;                  /* Philox mode only: -1 = auto, 0 = one thread group per chain walks the slice
                                          windows of a pass one after the other, 1 = team: four warps per chain sweep
                                          four disjoint windows concurrently (few chains per GPU).  MT19937 replay
                                          always runs the reference's order (vpi.f90:412-439). */
    int32_t chain_offset;              /* global index of this handle's chain 0 when the chains of one run are
                                          sharded over several handles/GPUs: chain c draws the Philox stream of
                                          global chain chain_offset + c (MT: sgrnd(seed + chain_offset + c)), so a
                                          run is reproducible whatever the number of shards */
    int32_t gpus;                      /* 0 or 1: one GPU (`device`).  G > 1: the handle shards the chains in
                                          contiguous blocks over the GPUs device .. device+G-1 of this process and
                                          sums the block accumulators itself at every block boundary (the
                                          reference's reduction point, vpi.f90:477-520); no MPI, NCCL or torch needed */
} pigs_params;

/* Raw block sums, summed over chains, exactly the quantities the driver holds
 * at the end of its step loop (vpi.f90:297-475) before NormalizeAv. */
typedef struct pigs_block_result {
    double  sumE, sumK, sumV, sumEt, sumKt, sumVt;        /* vpi.f90:456-457 */
    double  sumE2, sumK2, sumV2, sumEt2, sumKt2, sumVt2;  /* vpi.f90:459-460 */
    int64_t idiag_block, ngr;                             /* vpi.f90:410,464 */
    int64_t try_cm, try_stag, try_cm_half, try_stag_half; /* vpi.f90:335,350,375,381 */
    int64_t acc_cm, acc_bd, acc_head, acc_tail;
    int64_t acc_cm_half, acc_bd_half, acc_head_half, acc_tail_half;
    int64_t try_open, acc_open, try_close, acc_close, try_swap, acc_swap;
    int64_t bead_updates[3];          /* UpdateAction evaluations: [0] interior even, [1] odd, [2] end slices */
    int64_t n_open_chains;            /* chains whose worm is open at block end */
} pigs_block_result;

typedef struct pigs_ctx* pigs_handle;

const char* pigs_last_error(void);
int  pigs_version(void);

/* allocation of Path/xend/tables (vpi.f90:134-153) for n_chains replicas */
int  pigs_create(const pigs_params* p, pigs_handle* out);
int  pigs_destroy(pigs_handle h);

/* JastrowTable/PotentialTable results (vpi_mod.f90:84-145): host-filled, so the
 * reference's one-step index shift and NaN/-inf head entries are inherited;
 * zero tables give the non-interacting trap. */
int  pigs_set_tables(pigs_handle h, const double* LogWF, const double* VTable);

/* init / CheckPoint state of one chain (vpi_mod.f90:149-309) */
int  pigs_set_state(pigs_handle h, int chain, const double* Path, const double* xend, int isopen, int iworm);
int  pigs_get_state(pigs_handle h, int chain, double* Path, double* xend, int* isopen, int* iworm);
/* all chains at once: Path[n_chains][2Nb+1][Np][dim], xend[n_chains][2][dim], isopen/iworm[n_chains]
 * (host buffers; one H2D / D2H copy each) */
int  pigs_set_state_all(pigs_handle h, const double* Path, const double* xend, const int32_t* isopen, const int32_t* iworm);
int  pigs_get_state_all(pigs_handle h, double* Path, double* xend, int32_t* isopen, int32_t* iworm);
/* permutation bookkeeping of one chain (vpi.f90:67-72): iperm, new/end_perm_cycle,
 * Particles_in_perm_cycle(Np), Perm_histogram(Np) */
int  pigs_get_perm(pigs_handle h, int chain, int* iperm, int32_t* cycle, int32_t* hist, int* new_pc, int* end_pc);
int  pigs_set_perm(pigs_handle h, int chain, int iperm, const int32_t* cycle, const int32_t* hist, int new_pc, int end_pc);

/* random_mod.f90: sgrnd(seed) for one chain (chain<0: every chain c gets seed+c);
 * raw state access = mtsavef/mtgetf without the file (random_mod.f90:125-191) */
int  pigs_sgrnd(pigs_handle h, int chain, int32_t seed);
int  pigs_get_mt(pigs_handle h, int chain, uint32_t* mt624, int32_t* mti);
int  pigs_set_mt(pigs_handle h, int chain, const uint32_t* mt624, int32_t mti);
/* n draws of grnd() (random_mod.f90:35) / rangauss first deviate (random_mod.f90:195)
 * from one chain's device stream, in order */
int  pigs_grnd(pigs_handle h, int chain, int n, double* out);
int  pigs_rangauss(pigs_handle h, int chain, int n, double* out);

/* ---- production: the driver's step loop (vpi.f90:297-475) for Nstep steps on
 * every chain, on the device.  Block accumulators are zeroed first. */
int  pigs_run_block(pigs_handle h, int Nstep);
/* asynchronous variant: returns after the launch; pigs_sync() waits */
int  pigs_run_block_async(pigs_handle h, int Nstep);
int  pigs_sync(pigs_handle h);
/* sums over chains of the last block; gr[Nbin], Sk[Nk][dim], nrho[Nbin][Npw+1]
 * may be NULL.  gr/Sk/nrho are the raw histograms of THIS block (the driver
 * keeps summing nrho until an OBDM block closes, vpi.f90:522-539). */
int  pigs_get_block(pigs_handle h, pigs_block_result* out, double* gr, double* Sk, double* nrho);
/* same for one chain */
int  pigs_get_block_chain(pigs_handle h, int chain, pigs_block_result* out, double* gr, double* Sk, double* nrho);
/* The same for chains [chain0, chain0 + n) in one transfer: out[n], gr[n][Nbin], Sk[n][Nk][dim],
 * nrho[n][Nbin][Npw+1] (any of them may be NULL).  Cross-chain statistics replace the reference's single-chain
 * Var (sample_mod.f90:921-932). */
int  pigs_get_block_chains(pigs_handle h, int chain0, int n, pigs_block_result* out, double* gr, double* Sk, double* nrho);
/* device pointer + length (doubles) of the chain-summed accumulator vector of
 * the last block, laid out [12 energy sums | 24 counters as doubles | gr | Sk |
 * nrho]; for an in-place NCCL all-reduce by the caller (multi-GPU). */
int  pigs_block_vector(pigs_handle h, double** dev_ptr, int* n);
/* unpack such a vector (host copy, e.g. after the all-reduce) */
int  pigs_unpack_block_vector(pigs_handle h, const double* vec, pigs_block_result* out, double* gr, double* Sk, double* nrho);
/* elapsed device time of the last pigs_run_block, ms (CUDA events on the library's stream) */
int  pigs_last_block_ms(pigs_handle h, float* ms);
/* the CUDA stream (cudaStream_t) the library launches on, and launch counters */
int  pigs_stream(pigs_handle h, void** stream);
int  pigs_launch_count(pigs_handle h, int64_t* n);
/* the launch policy chosen for this handle (diagnostics): threads per chain group, chain groups per CTA, CTAs,
 * team schedule (1: four window workers per chain), table placement (0 L1/L2, 1 VTable pairs in smem, 2 both in
 * smem, 3 trap) */
int  pigs_launch_plan(pigs_handle h, int* threads_per_chain, int* groups_per_cta, int* grid, int* team, int* table_mode);

/* ---- multi-chain checkpoint (what the reference lacks) -------------------------------------------------------
 * The reference saves ONE chain as text (CheckPoint, vpi_mod.f90:263-309) plus its MT19937 state (mtsavef,
 * random_mod.f90:125-158), and loses every accumulator on restart; the drivers keep writing that pair for chain 0.
 * These two calls save / restore the complete device state of EVERY chain (paths, xend, worm and permutation
 * bookkeeping, Philox counters, MT19937 states) in one binary file in global chain order, so a run resumes bit for
 * bit -- also on a different number of GPUs.  The drivers add their own accumulators in checkpoint_driver.bin. */
int  pigs_save_checkpoint(pigs_handle h, const char* path);
int  pigs_load_checkpoint(pigs_handle h, const char* path);

/* ---- unit API: one reference procedure per call, applied to EVERY chain's
 * device state with that chain's own random stream ---- */
/* any of the 14 moves (PIGS_* above) for particle ip (1-based; for PIGS_SWAP the
 * worm iw) and half (1|2, *_HALF moves only).  accepted[n_chains] (may be NULL)
 * receives 0/1; aux[n_chains] (may be NULL) receives ipar for accepted swaps. */
int  pigs_move(pigs_handle h, int move, int ip, int half, int32_t* accepted, int32_t* aux);

/* UpdateAction (vpi_mod.f90:2491): n independent evaluations on host data.
 * R[n][Np][dim] slices, ip[n], ib[n], xnew[n][dim], xold[n][dim] -> DeltaS[n].
 * Evaluated with the reference's own roundings of the minimum image and of rij2 (pbc_mod.f90:29-52), as the MT19937
 * replay kernels do, so that partners lying exactly on the cutoff sphere (perfect crystal lattices) are decided as the
 * reference decides them.  A handle created with table_mode = 2 explicitly runs the production (Philox) instance of
 * the pair loop instead (shared-memory tables, fused r^2): identical except on that set of measure zero. */
int  pigs_update_action(pigs_handle h, int n, const double* R, const int32_t* ip, const int32_t* ib,
                        const double* xnew, const double* xold, double* DeltaS);
/* LocalEnergy (sample_mod.f90:154): R[n][Np][dim] -> E[n],Kin[n],Pot[n] */
int  pigs_local_energy(pigs_handle h, int n, const double* R, double* E, double* Kin, double* Pot);
/* ThermEnergy (sample_mod.f90:323): Path[n][2Nb+1][Np][dim] -> E[n],Ec[n],Ep[n] */
int  pigs_therm_energy(pigs_handle h, int n, const double* Path, double* E, double* Ec, double* Ep);
/* PairCorrelation (sample_mod.f90:392): gr[n][Nbin] += */
int  pigs_pair_correlation(pigs_handle h, int n, const double* R, double* gr);
/* StructureFactor (sample_mod.f90:435): Sk[n][Nk][dim] += */
int  pigs_structure_factor(pigs_handle h, int n, const double* R, double* Sk);
/* OBDM (sample_mod.f90:480): xend[n][2][dim]; nrho[n][Nbin][Npw+1] += */
int  pigs_obdm(pigs_handle h, int n, const double* xend, double* nrho);

/* FP64 FMA micro-benchmark: measured DFMA peak of the device in TFLOP/s (the
 * roofline denominator missing from MEASURED_PEAKS.json) */
int  pigs_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif
