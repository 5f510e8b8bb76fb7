// pigs_capi.cu -- the C ABI of libpigs_cuda (include/pigs_cuda.h): context,
// state transfer, launch policy.  No torch types, no CPU fallback.
#include "../../include/pigs_cuda.h"
#include "pigs_launch.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace pigs;

static thread_local std::string g_err;
// one parameter block per translation unit lives in __constant__ memory: the
// (upload, launch) pairs of different handles must not interleave
static std::mutex g_launch_mutex;
static int fail(int code, const std::string& m) { g_err = m; return code; }
#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(PIGS_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

struct pigs_ctx {
    pigs_params hp;
    DevParams P;
    int var = 0, mt = 0, T = 32, G = 1, grid = 1, block = 32, prefetch = 3, pfdist = 2, team = 0;
    size_t smem = 0;
    int nvec = 0;
    bool tables_set = false;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    long long launches = 0;
    // device buffers
    double *d_logwf = nullptr, *d_vtab = nullptr, *d_path = nullptr, *d_xend = nullptr, *d_acc = nullptr, *d_vec = nullptr, *d_pp = nullptr;
    double* d_stage = nullptr;      // AoS staging for state transfer
    int *d_istate = nullptr, *d_cyc = nullptr, *d_hist = nullptr, *d_iout = nullptr;
    unsigned* d_mt = nullptr;
    unsigned long long* d_pctr = nullptr;
    long long* d_cnt = nullptr;
    size_t stage_doubles = 0;
    int* d_flag = nullptr;          // device flag: coordinates outside the box seen by an upload
    // gpus > 1: this context owns no device memory; it shards the chains over one sub-context per GPU
    std::vector<pigs_ctx*> sub;
    std::vector<int> first;         // first[k] = global index of sub-context k's chain 0; first[gpus] = n_chains
};
// which sub-context holds global chain c (multi-GPU handles)
static int shard_of(const pigs_ctx* h, int c) {
    int k = 0;
    while (k + 1 < (int)h->sub.size() && c >= h->first[k + 1]) ++k;
    return k;
}
#define MULTI(h, expr) do { if ((h) && !(h)->sub.empty()) return (expr); } while (0)
// route a per-chain call of a multi-GPU handle to the shard that owns the chain
#define MULTI_CHAIN(h, c, call)                                                                                \
    do {                                                                                                       \
        if ((h) && !(h)->sub.empty()) {                                                                        \
            if ((c) < 0 || (c) >= (h)->hp.n_chains) return fail(PIGS_E_ARG, "chain index out of range");      \
            const int k_ = shard_of((h), (c));                                                                 \
            pigs_ctx* s_ = (h)->sub[k_];                                                                       \
            const int lc_ = (c) - (h)->first[k_];                                                              \
            (void)lc_;                                                                                         \
            return (call);                                                                                     \
        }                                                                                                      \
    } while (0)

extern "C" const char* pigs_last_error(void) { return g_err.c_str(); }
extern "C" int pigs_version(void) { return 100; }

static SweepArgs base_args(pigs_ctx* h) {
    SweepArgs A;
    std::memset(&A, 0, sizeof A);
    A.chain_only = -1;
    A.groups_per_cta = h->G;
    A.threads_per_chain = h->T;
    A.tshift = 0;
    while ((1 << A.tshift) < h->T) ++A.tshift;
    A.prefetch = h->prefetch;
    A.pfdist = h->pfdist;
    A.team = h->team;
    return A;
}
// The kernels read their parameters from one __constant__ block per translation unit and device, uploaded on the
// launching stream right before each launch.  A launch of ANOTHER handle on the same device must not overwrite the
// block while an earlier kernel (e.g. an asynchronous block of the persistent sweep kernel, which re-reads it in
// every phase) is still running: every (upload, launch) pair first waits for the previous handle's last launch.
struct ConstGuard {
    cudaEvent_t ev = nullptr;
    const pigs_ctx* last = nullptr;
};
static ConstGuard g_guard[64];
static int guard_enter(pigs_ctx* h) {          // g_launch_mutex held
    ConstGuard& G = g_guard[h->hp.device & 63];
    if (G.last && G.last != h && G.ev) CK(cudaStreamWaitEvent(h->st, G.ev, 0));
    return PIGS_OK;
}
static int guard_leave(pigs_ctx* h) {
    ConstGuard& G = g_guard[h->hp.device & 63];
    if (!G.ev) CK(cudaEventCreateWithFlags(&G.ev, cudaEventDisableTiming));
    CK(cudaEventRecord(G.ev, h->st));
    G.last = h;
    return PIGS_OK;
}
static int launch(pigs_ctx* h, const SweepArgs& A) {
    std::lock_guard<std::mutex> lk(g_launch_mutex);
    int rc = guard_enter(h);
    if (rc) return rc;
    CK(launch_sweep(h->mt, h->var, h->P, A, h->grid, h->block, h->smem, h->st));
    h->launches += 1;
    return guard_leave(h);
}

// launch policy: threads per chain T, chain groups per CTA G, table placement.
static int plan(pigs_ctx* h) {
    const pigs_params& p = h->hp;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, p.device));
    const int nsm = prop.multiProcessorCount;
    const size_t smem_max = prop.sharedMemPerBlockOptin - 64;      // static shared memory of the kernel (the staging mbarrier)
    int T = p.threads_per_chain;
    // Team mode (Philox only): T = 128, the four warps of a chain group sweep four disjoint slice windows of the
    // chain at once (win_sweep, pigs_sweep.cuh).  Chosen when one warp per chain would leave most of the GPU idle
    // (at most 8 chains per SM) and the path is long enough for a head, a tail and a middle window side by side.
    int team = 0;
    if (!h->mt && (T == 0 || T == 128)) {
        const int Lend = p.sampling ? (1 << (p.Nlev < 2 ? 2 : p.Nlev)) : p.Lstag;
        const int Lm = p.sampling ? (1 << p.Nlev) : p.Lstag;
        const bool fits = p.Np >= 4 && Lm >= 2 && (2 * p.Nb - 2 * Lend + 1) >= (Lm - 1);
        int want_team = p.schedule;
        const char* e = getenv("PIGS_SCHEDULE");      // tuning knob
        if (e) want_team = atoi(e);
        // measured on B200, N=64: 512 chains 294 -> 659 M bead-updates/s, 1024 chains 530 -> 615 M (two passes of four
        // teams per SM); from 2048 chains on one warp per chain wins (755 M)
        if (want_team < 0) want_team = (p.n_chains <= 8 * nsm) ? 1 : 0;
        if (want_team == 1 && !fits && p.schedule == 1) return fail(PIGS_E_ARG, "schedule = 1 (team): the path is too short for concurrent windows");
        team = (want_team == 1 && fits) ? 1 : 0;
        if (team) T = 128;
    } else if (p.schedule == 1) return fail(PIGS_E_ARG, "schedule = 1 (team) needs the Philox stream and threads_per_chain 0 or 128");
    h->team = team;
    if (T == 0) {
        // fill ~32 warps per SM: few chains -> wide groups, many chains -> one warp each
        int maxt0 = 1024;
        CK(sweep_max_threads(h->mt, p.trap ? 3 : 0, &maxt0));
        long long want = (long long)nsm * maxt0 / (p.n_chains > 0 ? p.n_chains : 1);
        // ... but a group only has nb x Np/32 independent warp-tasks per evaluation: wider than Np/64 warps
        // never paid (measured, 512 chains: N=64 T=32/64/128 -> 269/256/226 M bead-updates/s, N=256 -> 138/175/182)
        int tcap = 32;
        while (tcap * 2 <= 32 * (p.Np / 64) && tcap < 256) tcap *= 2;
        T = 32;
        while (T * 2 <= want && T < tcap) T *= 2;
    }
    if (T != 32 && T != 64 && T != 128 && T != 256 && T != 512) return fail(PIGS_E_ARG, "threads_per_chain must be 0,32,64,128,256,512");
    const size_t gbytes = ((grp_smem_bytes(h->P.S, h->P.Np, T / 32) + 15) & ~(size_t)15) * (team ? 4 : 1);      // team: one block per window worker
    const size_t tabbytes = (size_t)tab_len(p.Nmax) * sizeof(double);
    const size_t pairbytes = (size_t)(tab_len(p.Nmax) - 1) * 2 * sizeof(double);      // table_mode 1: {F(i),F(i+1)} pairs
    int maxt = 1024;
    CK(sweep_max_threads(h->mt, p.trap ? 3 : 0, &maxt));
    if (T > maxt) return fail(PIGS_E_ARG, "threads_per_chain exceeds the kernel's launch bound");
    int Gmax = maxt / T;
    if (T > 32 && Gmax > 16) Gmax = 16;        // named barriers 0..15
    int Gneed = (p.n_chains + nsm - 1) / nsm;  // chains per SM if spread evenly
    if (Gneed < 1) Gneed = 1;
    int var;
    if (p.trap) var = 3;
    else {
        int tm = p.table_mode;
        if (tm < 0) {
            // both tables in shared memory when at least min(Gneed,4) groups still fit beside them
            int gmin = Gneed < 4 ? Gneed : 4;
            int gfull = Gneed < Gmax ? Gneed : Gmax;
            // measured on B200 (C2/C3): both plain tables in smem > V pair table in smem > tables via L1/L2
            if (2 * tabbytes + (size_t)gfull * gbytes <= smem_max) tm = 2;
            else if (pairbytes + (size_t)gfull * gbytes <= smem_max) tm = 1;
            else if (2 * tabbytes + (size_t)gmin * gbytes <= smem_max) tm = 2;
            else tm = 0;
        }
        if (tm < 0 || tm > 2) return fail(PIGS_E_ARG, "table_mode must be -1,0,1,2");
        var = tm;
    }
    const size_t fixed = (var == 1 ? pairbytes : (var == 2 ? 2 * tabbytes : 0));
    if (fixed + gbytes > smem_max) return fail(PIGS_E_ARG, "configuration does not fit in shared memory; lower table_mode");
    int G = (int)((smem_max - fixed) / gbytes);
    if (G > Gmax) G = Gmax;
    if (G > Gneed) G = Gneed;
    if (G < 1) G = 1;
    if (Gneed > G) {
        // more chains than resident groups: the persistent CTAs make several passes.  Balance them when the last
        // pass would leave many SMs idle -- 3000 C3 chains on 148 SMs run as 2 passes of 11 groups per CTA (+13 %
        // over 16 + 5 with three quarters of the SMs idle in the second pass; a warp runs faster with fewer
        // neighbours).  A nearly full last pass is left alone (4096 C2 chains: 16 + 12 beat 14 + 14 by 4 %).
        const int passes = (Gneed + G - 1) / G;
        const int Gbal = (int)((p.n_chains + (long long)nsm * passes - 1) / ((long long)nsm * passes));
        if (Gbal >= 1 && Gbal <= G - 3) G = Gbal;
    }
    h->T = T; h->G = G; h->var = var; h->block = T * G;
    h->smem = fixed + (size_t)G * gbytes;
    CK(sweep_set_smem(h->mt, var, h->smem));
    int per_sm = 1;
    CK(sweep_occupancy(h->mt, var, h->block, h->smem, &per_sm));
    if (per_sm < 1) return fail(PIGS_E_CUDA, "sweep kernel cannot be resident with this configuration");
    int ctas = (p.n_chains + G - 1) / G;
    int cap = nsm * per_sm;
    h->grid = ctas < cap ? ctas : cap;
    return PIGS_OK;
}

// everything a single-GPU context owns; on any failure the caller destroys the half-built context (no leak)
static int init_single(pigs_ctx* h, const pigs_params* p) {
    CK(cudaSetDevice(p->device));
    h->hp = *p;
    h->mt = p->rng_mode == PIGS_RNG_MT_REPLAY ? 1 : 0;
    {
        const char* e = getenv("PIGS_PREFETCH");      // tuning knob (default on)
        if (e) h->prefetch = atoi(e);
        e = getenv("PIGS_PFDIST");
        if (e && atoi(e) >= 1) h->pfdist = atoi(e);
    }
    DevParams& P = h->P;
    std::memset(&P, 0, sizeof P);
    P.dim = p->dim; P.Np = p->Np; P.Nb = p->Nb; P.S = 2 * p->Nb + 1;
    P.NpS = (p->Np + 31) & ~31;               // whole [3][32] blocks (pidx, pigs_device.cuh)
    P.Nmax = p->Nmax; P.Nbin = p->Nbin; P.Nk = p->Nk; P.Npw = p->Npw;
    P.trap = p->trap != 0; P.sampling = p->sampling; P.Lstag = p->Lstag; P.Nlev = p->Nlev; P.Nstag = p->Nstag;
    P.Nobdm = p->Nobdm; P.swapping = p->swapping != 0; P.CMFreq = p->CMFreq; P.n_chains = p->n_chains;
    P.chain_offset = p->chain_offset;
    for (int k = 0; k < 3; ++k) {
        bool used = k < p->dim && !p->trap;
        P.L[k] = used ? p->Lbox[k] : 1e300;
        P.Lh[k] = used ? 0.5 * p->Lbox[k] : 0.5e300;          // LboxHalf, vpi.f90:118
        P.invL[k] = used ? 1.0 / p->Lbox[k] : 0.0;
        P.qbin[k] = used ? 2.0 * std::acos(-1.0) / p->Lbox[k] : 0.0;   // vpi.f90:119
        P.a_ho[k] = k < p->dim ? p->a_ho[k] : 1.0;
        unsigned long long bits;
        std::memcpy(&bits, &P.Lh[k], sizeof bits);
        const unsigned hi = (unsigned)(bits >> 32);
        std::memcpy(&P.LhF[k], &hi, sizeof hi);             // high word of L/2 as a float (mimg_hi)
    }
    P.tabW_off = tab_len(p->Nmax) * (int)sizeof(double);
    P.rcut2 = p->rcut * p->rcut;              // vpi.f90:127
    P.dr = p->dr; P.inv_dr = 1.0 / p->dr; P.half_inv_dr = 0.5 * P.inv_dr;
    P.rclamp2 = ((double)p->Nmax + 3.5) * p->dr * ((double)p->Nmax + 3.5) * p->dr;
    P.rbin = p->rcut / (double)(float)p->Nbin;   // vpi.f90:128
    P.dt = p->dt; P.delta_cm = p->delta_cm; P.CWorm = p->CWorm; P.density = p->density;
    P.pi = std::acos(-1.0);
    P.primitive = p->action == 1;
    if (P.primitive) {
        P.wS[0] = P.wS[1] = P.wS[2] = p->dt; P.cF = 0.0;
        P.wE[0] = P.wE[1] = P.wE[2] = 1.0; P.cFE = 0.0;
    } else {       // Chin (global_mod.f90:33-46, 52-65)
        P.wS[0] = 2.0 * p->dt / 3.0; P.wS[1] = 4.0 * p->dt / 3.0; P.wS[2] = p->dt / 3.0;
        P.cF = 4.0 * p->dt * p->dt * p->dt / 18.0;
        P.wE[0] = 2.0 / 3.0; P.wE[1] = 4.0 / 3.0; P.wE[2] = 1.0 / 3.0;
        P.cFE = (4.0 / 3.0) * (p->dt * p->dt * 0.5);
    }
    for (int n = 0; n < MAXS; ++n) {
        P.sig_free[n] = std::sqrt((double)(float)n * p->dt);
        P.sig_stage[n] = std::sqrt((double)((float)n / (float)(n + 1)) * p->dt);
    }
    for (int l = 0; l < 16; ++l) P.sig_bis[l] = std::sqrt(0.5 * (0.5 * (double)(float)(1 << l) * p->dt));
    P.half_inv_dt2 = 0.5 / (p->dt * p->dt);
    P.logCd = std::log(p->CWorm * p->density);   // -inf when CWorm = 0: every open attempt is rejected (F8)
    P.seed = p->seed;
    P.chain_stride = (size_t)P.S * 3 * P.NpS;
    P.off_gr = NE; P.off_sk = P.off_gr + P.Nbin; P.off_nr = P.off_sk + P.dim * P.Nk;
    P.nacc = P.off_nr + P.Nbin * (P.Npw + 1);
    h->nvec = NE + NCNT + (P.nacc - NE);

    const size_t nc = (size_t)p->n_chains;
#define ALLOC(ptr, count) CK(cudaMalloc((void**)&(ptr), sizeof(*(ptr)) * (count)))
    ALLOC(h->d_logwf, tab_len(p->Nmax)); ALLOC(h->d_vtab, tab_len(p->Nmax));      // zero tail: TAB_PAD
    ALLOC(h->d_path, nc * P.chain_stride); ALLOC(h->d_xend, nc * 6);
    ALLOC(h->d_istate, nc * IS_N); ALLOC(h->d_cyc, nc * P.Np); ALLOC(h->d_hist, nc * P.Np);
    ALLOC(h->d_mt, nc * 624); ALLOC(h->d_pctr, nc * PCS); ALLOC(h->d_acc, nc * P.nacc); ALLOC(h->d_cnt, nc * NCNT);
    ALLOC(h->d_vec, h->nvec); ALLOC(h->d_iout, nc * 2); ALLOC(h->d_pp, nc * P.Np);
#undef ALLOC
    CK(cudaMemset(h->d_path, 0, sizeof(double) * nc * P.chain_stride));
    CK(cudaMemset(h->d_xend, 0, sizeof(double) * nc * 6));
    CK(cudaMemset(h->d_istate, 0, sizeof(int) * nc * IS_N));
    CK(cudaMemset(h->d_cyc, 0, sizeof(int) * nc * P.Np));
    CK(cudaMemset(h->d_hist, 0, sizeof(int) * nc * P.Np));
    CK(cudaMemset(h->d_pctr, 0, sizeof(unsigned long long) * nc * PCS));
    CK(cudaMemset(h->d_acc, 0, sizeof(double) * nc * P.nacc));
    CK(cudaMemset(h->d_cnt, 0, sizeof(long long) * nc * NCNT));
    CK(cudaMemset(h->d_logwf, 0, sizeof(double) * tab_len(p->Nmax)));
    CK(cudaMemset(h->d_vtab, 0, sizeof(double) * tab_len(p->Nmax)));
    P.logwf = h->d_logwf; P.vtab = h->d_vtab; P.path = h->d_path; P.xend = h->d_xend; P.istate = h->d_istate;
    P.cyc = h->d_cyc; P.hist = h->d_hist; P.mt = h->d_mt; P.pctr = h->d_pctr; P.acc = h->d_acc; P.cnt = h->d_cnt; P.pp = h->d_pp;
    CK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->ev0)); CK(cudaEventCreate(&h->ev1));
    CK(cudaMalloc((void**)&h->d_flag, sizeof(int)));
    CK(cudaMemset(h->d_flag, 0, sizeof(int)));
    int rc = plan(h);
    if (rc != PIGS_OK) return rc;
    // sgrnd(seed + chain) for every chain (random_mod.f90:5-31)
    return pigs_sgrnd(h, -1, (int32_t)p->seed);
}

// gpus > 1: one sub-context per device device, device+1, ...; chains sharded in contiguous blocks, every shard
// knows the global index of its first chain (chain_offset), so each chain draws the same stream as in a 1-GPU run
static int create_multi(const pigs_params* p, int ndev, pigs_handle* out) {
    const int G = p->gpus;
    const bool same = getenv("PIGS_MULTI_SAME_DEVICE") != nullptr;      // testing aid: all shards on one device
    if (!same && p->device + G > ndev) return fail(PIGS_E_ARG, "gpus exceeds the devices available from `device` on");
    if (G > p->n_chains) return fail(PIGS_E_ARG, "more GPUs than chains: a single chain does not shard (replicas only)");
    pigs_ctx* h = new pigs_ctx();
    h->hp = *p;
    h->first.resize(G + 1);
    for (int k = 0; k <= G; ++k) h->first[k] = (int)((long long)p->n_chains * k / G);
    for (int k = 0; k < G; ++k) {
        pigs_params q = *p;
        q.gpus = 1;
        q.device = same ? p->device : p->device + k;
        q.n_chains = h->first[k + 1] - h->first[k];
        q.chain_offset = p->chain_offset + h->first[k];
        pigs_handle sk = nullptr;
        int rc = pigs_create(&q, &sk);
        if (rc != PIGS_OK) { pigs_destroy(h); return rc; }
        h->sub.push_back(sk);
    }
    h->P = h->sub[0]->P;
    h->nvec = h->sub[0]->nvec;
    *out = h;
    return PIGS_OK;
}

extern "C" int pigs_create(const pigs_params* p, pigs_handle* out) {
    if (!p || !out) return fail(PIGS_E_ARG, "null argument");
    *out = nullptr;
    if (p->dim < 1 || p->dim > 3) return fail(PIGS_E_ARG, "dim must be 1..3");
    if (2 * p->Nb + 1 > MAXS) return fail(PIGS_E_ARG, "Nb too large (2*Nb+1 must not exceed 132)");
    if (p->Np < 2 || p->Nb < 1 || p->Nmax < 4 || p->n_chains < 1) return fail(PIGS_E_ARG, "Np>=2, Nb>=1, Nmax>=4, n_chains>=1 required");
    if (p->Nbin < 1 || p->Nk < 0 || p->Npw < 0) return fail(PIGS_E_ARG, "bad Nbin/Nk/Npw");
    if (p->sampling != 0 && p->sampling != 1) return fail(PIGS_E_ARG, "sampling must be 0 ('sta') or 1 ('bis')");
    if (p->CMFreq < 1 || p->Nstag < 0 || p->Nobdm < 0) return fail(PIGS_E_ARG, "bad CMFreq/Nstag/Nobdm");
    // constraints implied by the reference's index arithmetic
    if (p->sampling == 1) {
        int lv = p->Nlev < 2 ? 2 : p->Nlev;     // head/tail bisection draw Nlev' in [2, max(2,Nlev)]
        if ((1 << lv) > 2 * p->Nb) return fail(PIGS_E_ARG, "2**Nlev must not exceed 2*Nb");
    } else if (p->Lstag < 2 || p->Lstag > 2 * p->Nb) return fail(PIGS_E_ARG, "2 <= Lstag <= 2*Nb required");
    if ((p->Nobdm > 0 || p->CWorm > 0 || p->swapping) && (p->Lstag < 2 || p->Lstag > p->Nb))
        return fail(PIGS_E_ARG, "worm moves need 2 <= Lstag <= Nb");
    if (p->Lstag < 2) return fail(PIGS_E_ARG, "Lstag >= 2 required (OpenChain is attempted even when CWorm = 0)");
    if (p->Lstag > p->Nb) return fail(PIGS_E_ARG, "Lstag <= Nb required (OpenChain is attempted even when CWorm = 0)");
    if (p->rng_mode != PIGS_RNG_PHILOX && p->rng_mode != PIGS_RNG_MT_REPLAY) return fail(PIGS_E_ARG, "bad rng_mode");
    if (p->action != 0 && p->action != 1) return fail(PIGS_E_ARG, "action must be 0 (Chin) or 1 (primitive)");
    if (p->schedule < -1 || p->schedule > 1) return fail(PIGS_E_ARG, "schedule must be -1 (auto), 0 or 1 (team)");
    if (p->chain_offset < 0) return fail(PIGS_E_ARG, "chain_offset < 0");
    if (!(p->dt > 0) || !(p->dr > 0) || !(p->rcut > 0)) return fail(PIGS_E_ARG, "dt, dr, rcut must be positive");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(PIGS_E_CUDA, std::string("no CUDA device: libpigs_cuda has no CPU fallback (") + cudaGetErrorString(e) + ")");
    if (p->device < 0 || p->device >= ndev) return fail(PIGS_E_ARG, "bad device ordinal");
    if (p->gpus < 0) return fail(PIGS_E_ARG, "gpus < 0");
    if (p->gpus > 1) return create_multi(p, ndev, out);
    pigs_ctx* h = new pigs_ctx();
    const int rc = init_single(h, p);
    if (rc != PIGS_OK) { pigs_destroy(h); return rc; }
    *out = h;
    return PIGS_OK;
}

extern "C" int pigs_destroy(pigs_handle h) {
    if (!h) return PIGS_OK;
    if (!h->sub.empty() || !h->first.empty()) {
        for (pigs_ctx* s : h->sub) pigs_destroy(s);
        delete h;
        return PIGS_OK;
    }
    cudaSetDevice(h->hp.device);
    {   // nothing of this handle may still be reading the constant block when the next handle uploads
        std::lock_guard<std::mutex> lk(g_launch_mutex);
        ConstGuard& G = g_guard[h->hp.device & 63];
        if (G.last == h) { if (h->st) cudaStreamSynchronize(h->st); G.last = nullptr; }
    }
    cudaFree(h->d_flag);
    cudaFree(h->d_logwf); cudaFree(h->d_vtab); cudaFree(h->d_path); cudaFree(h->d_xend); cudaFree(h->d_istate);
    cudaFree(h->d_cyc); cudaFree(h->d_hist); cudaFree(h->d_mt); cudaFree(h->d_pctr); cudaFree(h->d_acc);
    cudaFree(h->d_cnt); cudaFree(h->d_vec); cudaFree(h->d_iout); cudaFree(h->d_stage); cudaFree(h->d_pp);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return PIGS_OK;
}

#define NEED(h) do { if (!(h)) return fail(PIGS_E_ARG, "null handle"); CK(cudaSetDevice((h)->hp.device)); } while (0)
#define NEED_CHAIN(h, c) do { if ((c) < 0 || (c) >= (h)->hp.n_chains) return fail(PIGS_E_ARG, "chain index out of range"); } while (0)

static int ensure_stage(pigs_ctx* h, size_t doubles) {
    if (h->stage_doubles >= doubles) return PIGS_OK;
    if (h->d_stage) CK(cudaFree(h->d_stage));
    h->d_stage = nullptr; h->stage_doubles = 0;
    CK(cudaMalloc((void**)&h->d_stage, doubles * sizeof(double)));
    h->stage_doubles = doubles;
    return PIGS_OK;
}

extern "C" int pigs_set_tables(pigs_handle h, const double* LogWF, const double* VTable) {
    if (h && !h->sub.empty()) {
        for (pigs_ctx* s : h->sub) { int rc = pigs_set_tables(s, LogWF, VTable); if (rc) return rc; }
        h->tables_set = true;
        return PIGS_OK;
    }
    NEED(h);
    if (!LogWF || !VTable) return fail(PIGS_E_ARG, "null table");
    size_t n = sizeof(double) * (h->hp.Nmax + 2);
    CK(cudaMemcpyAsync(h->d_logwf, LogWF, n, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->d_vtab, VTable, n, cudaMemcpyHostToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));
    h->tables_set = true;
    return PIGS_OK;
}

// ---- state transfer ---------------------------------------------------------------------------
static int put_paths(pigs_ctx* h, int chain0, int nchain, const double* Path) {
    const DevParams& P = h->P;
    size_t per = (size_t)P.S * P.Np * P.dim;
    int rc = ensure_stage(h, per * nchain);
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->d_stage, Path, per * nchain * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemsetAsync(h->d_flag, 0, sizeof(int), h->st));
    CK(launch_aos_to_soa(P, h->d_stage, h->d_path + (size_t)chain0 * P.chain_stride, nchain, h->st, h->d_flag));
    h->launches += 1;
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (bad) return fail(PIGS_E_ARG, "Path has coordinates outside the periodic box [-L/2, L/2] (or NaN): wrap them first "
                                     "(BoundaryConditions, pbc_mod.f90:11-25) -- the action takes one periodic image per component");
    return PIGS_OK;
}
static int get_paths(pigs_ctx* h, int chain0, int nchain, double* Path) {
    const DevParams& P = h->P;
    size_t per = (size_t)P.S * P.Np * P.dim;
    int rc = ensure_stage(h, per * nchain);
    if (rc) return rc;
    CK(launch_soa_to_aos(P, h->d_path + (size_t)chain0 * P.chain_stride, h->d_stage, nchain, h->st));
    h->launches += 1;
    CK(cudaMemcpyAsync(Path, h->d_stage, per * nchain * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    return PIGS_OK;
}
static int put_scalars(pigs_ctx* h, int chain0, int nchain, const double* xend, const int32_t* isopen, const int32_t* iworm) {
    const int dim = h->P.dim;
    std::vector<double> xe((size_t)nchain * 6, 0.0);
    for (int c = 0; c < nchain; ++c) for (int j = 0; j < 2; ++j) for (int k = 0; k < dim; ++k) xe[(size_t)c * 6 + j * 3 + k] = xend[((size_t)c * 2 + j) * dim + k];
    CK(cudaMemcpyAsync(h->d_xend + (size_t)chain0 * 6, xe.data(), xe.size() * sizeof(double), cudaMemcpyHostToDevice, h->st));
    std::vector<int> ist((size_t)nchain * IS_N);
    CK(cudaMemcpyAsync(ist.data(), h->d_istate + (size_t)chain0 * IS_N, ist.size() * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    for (int c = 0; c < nchain; ++c) {
        if (isopen[c] && (iworm[c] < 1 || iworm[c] > h->P.Np)) return fail(PIGS_E_ARG, "isopen needs 1 <= iworm <= Np");
        ist[(size_t)c * IS_N + IS_OPEN] = isopen[c] != 0;
        ist[(size_t)c * IS_N + IS_IWORM] = (iworm[c] >= 0 && iworm[c] <= h->P.Np) ? iworm[c] : 0;
    }
    CK(cudaMemcpyAsync(h->d_istate + (size_t)chain0 * IS_N, ist.data(), ist.size() * sizeof(int), cudaMemcpyHostToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));
    return PIGS_OK;
}
static int get_scalars(pigs_ctx* h, int chain0, int nchain, double* xend, int32_t* isopen, int32_t* iworm) {
    const int dim = h->P.dim;
    std::vector<double> xe((size_t)nchain * 6);
    std::vector<int> ist((size_t)nchain * IS_N);
    CK(cudaMemcpyAsync(xe.data(), h->d_xend + (size_t)chain0 * 6, xe.size() * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(ist.data(), h->d_istate + (size_t)chain0 * IS_N, ist.size() * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    for (int c = 0; c < nchain; ++c) {
        if (xend) for (int j = 0; j < 2; ++j) for (int k = 0; k < dim; ++k) xend[((size_t)c * 2 + j) * dim + k] = xe[(size_t)c * 6 + j * 3 + k];
        if (isopen) isopen[c] = ist[(size_t)c * IS_N + IS_OPEN];
        if (iworm) iworm[c] = ist[(size_t)c * IS_N + IS_IWORM];
    }
    return PIGS_OK;
}

extern "C" int pigs_set_state(pigs_handle h, int chain, const double* Path, const double* xend, int isopen, int iworm) {
    MULTI_CHAIN(h, chain, pigs_set_state(s_, lc_, Path, xend, isopen, iworm));
    NEED(h); NEED_CHAIN(h, chain);
    if (!Path || !xend) return fail(PIGS_E_ARG, "null state");
    int rc = put_paths(h, chain, 1, Path);
    if (rc) return rc;
    int32_t io = isopen, iw = iworm;
    return put_scalars(h, chain, 1, xend, &io, &iw);
}
extern "C" int pigs_get_state(pigs_handle h, int chain, double* Path, double* xend, int* isopen, int* iworm) {
    MULTI_CHAIN(h, chain, pigs_get_state(s_, lc_, Path, xend, isopen, iworm));
    NEED(h); NEED_CHAIN(h, chain);
    if (Path) { int rc = get_paths(h, chain, 1, Path); if (rc) return rc; }
    int32_t io = 0, iw = 0;
    int rc = get_scalars(h, chain, 1, xend, &io, &iw);
    if (isopen) *isopen = io;
    if (iworm) *iworm = iw;
    return rc;
}
extern "C" int pigs_set_state_all(pigs_handle h, const double* Path, const double* xend, const int32_t* isopen, const int32_t* iworm) {
    if (h && !h->sub.empty()) {
        if (!Path || !xend || !isopen || !iworm) return fail(PIGS_E_ARG, "null state");
        const size_t per = (size_t)h->P.S * h->P.Np * h->P.dim;
        for (size_t k = 0; k < h->sub.size(); ++k) {
            const size_t f = (size_t)h->first[k];
            int rc = pigs_set_state_all(h->sub[k], Path + f * per, xend + f * 2 * h->P.dim, isopen + f, iworm + f);
            if (rc) return rc;
        }
        return PIGS_OK;
    }
    NEED(h);
    if (!Path || !xend || !isopen || !iworm) return fail(PIGS_E_ARG, "null state");
    int rc = put_paths(h, 0, h->hp.n_chains, Path);
    if (rc) return rc;
    return put_scalars(h, 0, h->hp.n_chains, xend, isopen, iworm);
}
extern "C" int pigs_get_state_all(pigs_handle h, double* Path, double* xend, int32_t* isopen, int32_t* iworm) {
    if (h && !h->sub.empty()) {
        const size_t per = (size_t)h->P.S * h->P.Np * h->P.dim;
        for (size_t k = 0; k < h->sub.size(); ++k) {
            const size_t f = (size_t)h->first[k];
            int rc = pigs_get_state_all(h->sub[k], Path ? Path + f * per : nullptr, xend ? xend + f * 2 * h->P.dim : nullptr,
                                        isopen ? isopen + f : nullptr, iworm ? iworm + f : nullptr);
            if (rc) return rc;
        }
        return PIGS_OK;
    }
    NEED(h);
    if (Path) { int rc = get_paths(h, 0, h->hp.n_chains, Path); if (rc) return rc; }
    return get_scalars(h, 0, h->hp.n_chains, xend, isopen, iworm);
}

extern "C" int pigs_get_perm(pigs_handle h, int chain, int* iperm, int32_t* cycle, int32_t* hist, int* new_pc, int* end_pc) {
    MULTI_CHAIN(h, chain, pigs_get_perm(s_, lc_, iperm, cycle, hist, new_pc, end_pc));
    NEED(h); NEED_CHAIN(h, chain);
    int ist[IS_N];
    CK(cudaMemcpyAsync(ist, h->d_istate + (size_t)chain * IS_N, sizeof ist, cudaMemcpyDeviceToHost, h->st));
    if (cycle) CK(cudaMemcpyAsync(cycle, h->d_cyc + (size_t)chain * h->P.Np, sizeof(int) * h->P.Np, cudaMemcpyDeviceToHost, h->st));
    if (hist) CK(cudaMemcpyAsync(hist, h->d_hist + (size_t)chain * h->P.Np, sizeof(int) * h->P.Np, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (iperm) *iperm = ist[IS_IPERM];
    if (new_pc) *new_pc = ist[IS_NEWPC];
    if (end_pc) *end_pc = ist[IS_ENDPC];
    return PIGS_OK;
}
extern "C" int pigs_set_perm(pigs_handle h, int chain, int iperm, const int32_t* cycle, const int32_t* hist, int new_pc, int end_pc) {
    MULTI_CHAIN(h, chain, pigs_set_perm(s_, lc_, iperm, cycle, hist, new_pc, end_pc));
    NEED(h); NEED_CHAIN(h, chain);
    int ist[IS_N];
    CK(cudaMemcpyAsync(ist, h->d_istate + (size_t)chain * IS_N, sizeof ist, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (iperm < 0 || iperm > h->P.Np) return fail(PIGS_E_ARG, "iperm must be in 0..Np");
    ist[IS_IPERM] = iperm; ist[IS_NEWPC] = new_pc != 0; ist[IS_ENDPC] = end_pc != 0;
    CK(cudaMemcpyAsync(h->d_istate + (size_t)chain * IS_N, ist, sizeof ist, cudaMemcpyHostToDevice, h->st));
    if (cycle) CK(cudaMemcpyAsync(h->d_cyc + (size_t)chain * h->P.Np, cycle, sizeof(int) * h->P.Np, cudaMemcpyHostToDevice, h->st));
    if (hist) CK(cudaMemcpyAsync(h->d_hist + (size_t)chain * h->P.Np, hist, sizeof(int) * h->P.Np, cudaMemcpyHostToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));
    return PIGS_OK;
}

// ---- multi-chain checkpoint ---------------------------------------------------------------------
// The reference checkpoints ONE chain as text (CheckPoint, vpi_mod.f90:263-309) and its RNG state (mtsavef,
// random_mod.f90:125-158); the driver keeps writing those for chain 0.  This is the extension SURVEY 8(f2) asks
// for: every chain's path, xend, worm / permutation state, Philox counters and MT19937 state in one binary file,
// in GLOBAL chain order, so a run can be resumed on a different number of GPUs.  Little-endian, version 1:
//   "PIGSCKP1" | int32 dim, Np, Nb, n_chains, rng_mode, PCS, IS_N, reserved | uint64 seed |
//   Path[n][2Nb+1][Np][dim] f64 | xend[n][2][dim] f64 | istate[n][IS_N] i32 | cyc[n][Np] i32 | hist[n][Np] i32 |
//   pctr[n][PCS] u64 | mt[n][624] u32
static const char CKP_MAGIC[8] = {'P', 'I', 'G', 'S', 'C', 'K', 'P', '1'};
struct CkpHeader {
    char magic[8];
    int32_t dim, Np, Nb, n_chains, rng_mode, pcs, is_n, reserved;
    uint64_t seed;
};
template <class T>
static int ckp_array(pigs_ctx* h, FILE* f, bool save, T* dptr, size_t per_chain) {
    std::vector<T> buf((size_t)h->hp.n_chains * per_chain);
    if (save) {
        CK(cudaMemcpyAsync(buf.data(), dptr, buf.size() * sizeof(T), cudaMemcpyDeviceToHost, h->st));
        CK(cudaStreamSynchronize(h->st));
        if (fwrite(buf.data(), sizeof(T), buf.size(), f) != buf.size()) return fail(PIGS_E_STATE, "checkpoint: short write");
    } else {
        if (fread(buf.data(), sizeof(T), buf.size(), f) != buf.size()) return fail(PIGS_E_STATE, "checkpoint: short read");
        CK(cudaMemcpyAsync(dptr, buf.data(), buf.size() * sizeof(T), cudaMemcpyHostToDevice, h->st));
        CK(cudaStreamSynchronize(h->st));
    }
    return PIGS_OK;
}
static int ckp_paths(pigs_ctx* h, FILE* f, bool save) {
    const DevParams& P = h->P;
    const size_t per = (size_t)P.S * P.Np * P.dim;
    std::vector<double> buf((size_t)h->hp.n_chains * per), xe((size_t)h->hp.n_chains * 2 * P.dim);
    std::vector<int32_t> io(h->hp.n_chains), iw(h->hp.n_chains);
    if (save) {
        int rc = get_paths(h, 0, h->hp.n_chains, buf.data());
        if (rc) return rc;
        CK(cudaStreamSynchronize(h->st));
        if (fwrite(buf.data(), sizeof(double), buf.size(), f) != buf.size()) return fail(PIGS_E_STATE, "checkpoint: short write");
    } else {
        if (fread(buf.data(), sizeof(double), buf.size(), f) != buf.size()) return fail(PIGS_E_STATE, "checkpoint: short read");
        int rc = put_paths(h, 0, h->hp.n_chains, buf.data());
        if (rc) return rc;
    }
    return PIGS_OK;
}
// one pass over the file per array, shard after shard: the file is in global chain order
static int ckp_io(pigs_ctx* h, const char* path, bool save) {
    if (!path) return fail(PIGS_E_ARG, "null path");
    std::vector<pigs_ctx*> parts = h->sub.empty() ? std::vector<pigs_ctx*>{h} : h->sub;
    for (pigs_ctx* s : parts) { CK(cudaSetDevice(s->hp.device)); CK(cudaStreamSynchronize(s->st)); }
    FILE* f = fopen(path, save ? "wb" : "rb");
    if (!f) return fail(PIGS_E_STATE, std::string("cannot open ") + path);
    CkpHeader H;
    std::memset(&H, 0, sizeof H);
    const pigs_params& p = h->hp;
    int rc = PIGS_OK;
    if (save) {
        std::memcpy(H.magic, CKP_MAGIC, 8);
        H.dim = p.dim; H.Np = p.Np; H.Nb = p.Nb; H.n_chains = p.n_chains; H.rng_mode = p.rng_mode; H.pcs = PCS; H.is_n = IS_N;
        H.seed = p.seed;
        if (fwrite(&H, sizeof H, 1, f) != 1) rc = fail(PIGS_E_STATE, "checkpoint: short write");
    } else {
        if (fread(&H, sizeof H, 1, f) != 1 || std::memcmp(H.magic, CKP_MAGIC, 8) != 0) rc = fail(PIGS_E_STATE, "not a PIGSCKP1 checkpoint");
        else if (H.dim != p.dim || H.Np != p.Np || H.Nb != p.Nb || H.n_chains != p.n_chains || H.rng_mode != p.rng_mode || H.pcs != PCS || H.is_n != IS_N)
            rc = fail(PIGS_E_ARG, "checkpoint was written for another configuration (dim, Np, Nb, n_chains or rng_mode differ)");
    }
    for (int arr = 0; arr < 7 && rc == PIGS_OK; ++arr)
        for (pigs_ctx* s : parts) {
            if (cudaSetDevice(s->hp.device) != cudaSuccess) { rc = fail(PIGS_E_CUDA, "cudaSetDevice"); break; }
            const DevParams& P = s->P;
            switch (arr) {
            case 0: rc = ckp_paths(s, f, save); break;
            case 1: {       // xend travels as [2][dim] like the ABI, the device keeps [2][3]
                std::vector<double> x6((size_t)s->hp.n_chains * 6, 0.0), xd((size_t)s->hp.n_chains * 2 * P.dim);
                if (save) {
                    if (cudaMemcpy(x6.data(), s->d_xend, x6.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { rc = fail(PIGS_E_CUDA, "xend copy"); break; }
                    for (int c = 0; c < s->hp.n_chains; ++c) for (int j = 0; j < 2; ++j) for (int k = 0; k < P.dim; ++k) xd[((size_t)c * 2 + j) * P.dim + k] = x6[(size_t)c * 6 + j * 3 + k];
                    if (fwrite(xd.data(), 8, xd.size(), f) != xd.size()) rc = fail(PIGS_E_STATE, "checkpoint: short write");
                } else {
                    if (fread(xd.data(), 8, xd.size(), f) != xd.size()) { rc = fail(PIGS_E_STATE, "checkpoint: short read"); break; }
                    for (int c = 0; c < s->hp.n_chains; ++c) for (int j = 0; j < 2; ++j) for (int k = 0; k < P.dim; ++k) x6[(size_t)c * 6 + j * 3 + k] = xd[((size_t)c * 2 + j) * P.dim + k];
                    if (cudaMemcpy(s->d_xend, x6.data(), x6.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(PIGS_E_CUDA, "xend copy");
                }
                break;
            }
            case 2: rc = ckp_array(s, f, save, s->d_istate, (size_t)IS_N); break;
            case 3: rc = ckp_array(s, f, save, s->d_cyc, (size_t)P.Np); break;
            case 4: rc = ckp_array(s, f, save, s->d_hist, (size_t)P.Np); break;
            case 5: rc = ckp_array(s, f, save, s->d_pctr, (size_t)PCS); break;
            case 6: rc = ckp_array(s, f, save, s->d_mt, (size_t)624); break;
            }
            if (rc != PIGS_OK) break;
        }
    fclose(f);
    return rc;
}
extern "C" int pigs_save_checkpoint(pigs_handle h, const char* path) {
    if (!h) return fail(PIGS_E_ARG, "null handle");
    return ckp_io(h, path, true);
}
extern "C" int pigs_load_checkpoint(pigs_handle h, const char* path) {
    if (!h) return fail(PIGS_E_ARG, "null handle");
    return ckp_io(h, path, false);
}

// ---- random streams ----------------------------------------------------------------------------
extern "C" int pigs_sgrnd(pigs_handle h, int chain, int32_t seed) {
    if (h && !h->sub.empty()) {
        if (chain < 0) { for (pigs_ctx* s : h->sub) { int rc = pigs_sgrnd(s, -1, seed); if (rc) return rc; } return PIGS_OK; }
        MULTI_CHAIN(h, chain, pigs_sgrnd(s_, lc_, seed));
    }
    NEED(h);
    if (chain >= h->hp.n_chains) return fail(PIGS_E_ARG, "chain index out of range");
    SweepArgs A = base_args(h);
    A.op = OP_SEED; A.seed = seed; A.chain_only = chain < 0 ? -1 : chain;
    int rc = launch(h, A);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->st));
    return PIGS_OK;
}
extern "C" int pigs_get_mt(pigs_handle h, int chain, uint32_t* mt624, int32_t* mti) {
    MULTI_CHAIN(h, chain, pigs_get_mt(s_, lc_, mt624, mti));
    NEED(h); NEED_CHAIN(h, chain);
    int v = 0;
    if (mt624) CK(cudaMemcpyAsync(mt624, h->d_mt + (size_t)chain * 624, 624 * sizeof(unsigned), cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(&v, h->d_istate + (size_t)chain * IS_N + IS_MTI, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    if (mti) *mti = v;
    return PIGS_OK;
}
extern "C" int pigs_set_mt(pigs_handle h, int chain, const uint32_t* mt624, int32_t mti) {
    MULTI_CHAIN(h, chain, pigs_set_mt(s_, lc_, mt624, mti));
    NEED(h); NEED_CHAIN(h, chain);
    if (!mt624) return fail(PIGS_E_ARG, "null mt state");
    if (mti < 0 || mti > 625) return fail(PIGS_E_ARG, "mti must be in 0..625");
    int v = mti;
    CK(cudaMemcpyAsync(h->d_mt + (size_t)chain * 624, mt624, 624 * sizeof(unsigned), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->d_istate + (size_t)chain * IS_N + IS_MTI, &v, sizeof(int), cudaMemcpyHostToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));
    return PIGS_OK;
}
static int draws(pigs_ctx* h, int op, int chain, int n, double* out) {
    if (n < 0 || !out) return fail(PIGS_E_ARG, "bad draw request");
    if (n == 0) return PIGS_OK;
    int rc = ensure_stage(h, (size_t)n);
    if (rc) return rc;
    SweepArgs A = base_args(h);
    A.op = op; A.nstep = n; A.chain_only = chain; A.draws = h->d_stage;
    rc = launch(h, A);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, h->d_stage, sizeof(double) * n, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return PIGS_OK;
}
extern "C" int pigs_grnd(pigs_handle h, int chain, int n, double* out) { MULTI_CHAIN(h, chain, pigs_grnd(s_, lc_, n, out)); NEED(h); NEED_CHAIN(h, chain); return draws(h, OP_UNIFORM, chain, n, out); }
extern "C" int pigs_rangauss(pigs_handle h, int chain, int n, double* out) { MULTI_CHAIN(h, chain, pigs_rangauss(s_, lc_, n, out)); NEED(h); NEED_CHAIN(h, chain); return draws(h, OP_GAUSS, chain, n, out); }

// ---- production path ---------------------------------------------------------------------------
extern "C" int pigs_run_block_async(pigs_handle h, int Nstep) {
    if (h && !h->sub.empty()) {          // every GPU gets its launch before anybody waits
        for (pigs_ctx* s : h->sub) { int rc = pigs_run_block_async(s, Nstep); if (rc) return rc; }
        return PIGS_OK;
    }
    NEED(h);
    if (!h->tables_set) return fail(PIGS_E_STATE, "pigs_set_tables has not been called");
    if (Nstep < 0) return fail(PIGS_E_ARG, "Nstep < 0");
    CK(cudaEventRecord(h->ev0, h->st));
    CK(launch_zero_block(h->P, h->st));
    SweepArgs A = base_args(h);
    A.op = OP_BLOCK; A.nstep = Nstep;
    int rc = launch(h, A);
    if (rc) return rc;
    CK(launch_reduce_block(h->P, h->d_vec, h->st));
    h->launches += 2;
    CK(cudaEventRecord(h->ev1, h->st));
    return PIGS_OK;
}
extern "C" int pigs_sync(pigs_handle h) {
    if (h && !h->sub.empty()) {
        for (pigs_ctx* s : h->sub) { int rc = pigs_sync(s); if (rc) return rc; }
        return PIGS_OK;
    }
    NEED(h);
    CK(cudaStreamSynchronize(h->st));
    return PIGS_OK;
}
extern "C" int pigs_run_block(pigs_handle h, int Nstep) {
    if (h && !h->sub.empty()) {
        int rc = pigs_run_block_async(h, Nstep);
        if (rc) return rc;
        h->last_ms = 0.f;
        for (pigs_ctx* s : h->sub) {
            float ms = 0.f;
            rc = pigs_sync(s);
            if (rc) return rc;
            rc = pigs_last_block_ms(s, &ms);
            if (rc) return rc;
            if (ms > h->last_ms) h->last_ms = ms;      // the block ends when the slowest GPU ends
        }
        return PIGS_OK;
    }
    int rc = pigs_run_block_async(h, Nstep);
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->st));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    return PIGS_OK;
}
extern "C" int pigs_last_block_ms(pigs_handle h, float* ms) {
    if (h && !h->sub.empty()) {
        float m = 0.f;
        for (pigs_ctx* s : h->sub) { float x = 0.f; int rc = pigs_last_block_ms(s, &x); if (rc) return rc; if (x > m) m = x; }
        if (ms) *ms = m;
        return PIGS_OK;
    }
    NEED(h);
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    if (ms) *ms = h->last_ms;
    return PIGS_OK;
}
extern "C" int pigs_launch_plan(pigs_handle h, int* threads_per_chain, int* groups_per_cta, int* grid, int* team, int* table_mode) {
    MULTI(h, pigs_launch_plan(h->sub[0], threads_per_chain, groups_per_cta, grid, team, table_mode));
    NEED(h);
    if (threads_per_chain) *threads_per_chain = h->T;
    if (groups_per_cta) *groups_per_cta = h->G;
    if (grid) *grid = h->grid;
    if (team) *team = h->team;
    if (table_mode) *table_mode = h->var;
    return PIGS_OK;
}
extern "C" int pigs_stream(pigs_handle h, void** stream) { MULTI(h, pigs_stream(h->sub[0], stream)); NEED(h); if (stream) *stream = (void*)h->st; return PIGS_OK; }
extern "C" int pigs_launch_count(pigs_handle h, int64_t* n) {
    if (h && !h->sub.empty()) {
        int64_t t = 0;
        for (pigs_ctx* s : h->sub) { int64_t x = 0; pigs_launch_count(s, &x); t += x; }
        if (n) *n = t;
        return PIGS_OK;
    }
    NEED(h); if (n) *n = h->launches; return PIGS_OK; }

static void unpack(const pigs_ctx* h, const double* vec, pigs_block_result* out, double* gr, double* Sk, double* nrho) {
    const DevParams& P = h->P;
    if (out) {
        double* e = &out->sumE;
        for (int i = 0; i < NE; ++i) e[i] = vec[i];
        int64_t* c = &out->idiag_block;
        for (int i = 0; i < NCNT; ++i) c[i] = (int64_t)llround(vec[NE + i]);
    }
    const double* hst = vec + NE + NCNT;
    if (gr) std::memcpy(gr, hst, sizeof(double) * P.Nbin);
    if (Sk) std::memcpy(Sk, hst + P.Nbin, sizeof(double) * P.dim * P.Nk);
    if (nrho) std::memcpy(nrho, hst + P.Nbin + P.dim * P.Nk, sizeof(double) * P.Nbin * (P.Npw + 1));
}
extern "C" int pigs_get_block(pigs_handle h, pigs_block_result* out, double* gr, double* Sk, double* nrho) {
    if (h && !h->sub.empty()) {
        // the block-boundary reduction over GPUs (the reference's reduction point, vpi.f90:477-520): each GPU has
        // already summed its chains in a fixed order (k_reduce_block); the ~3 KB vectors are added here in GPU order
        std::vector<double> tot((size_t)h->nvec, 0.0), v((size_t)h->nvec);
        for (pigs_ctx* s : h->sub) {
            CK(cudaSetDevice(s->hp.device));
            CK(cudaMemcpyAsync(v.data(), s->d_vec, sizeof(double) * h->nvec, cudaMemcpyDeviceToHost, s->st));
            CK(cudaStreamSynchronize(s->st));
            for (int i = 0; i < h->nvec; ++i) tot[i] += v[i];
        }
        unpack(h->sub[0], tot.data(), out, gr, Sk, nrho);
        return PIGS_OK;
    }
    NEED(h);
    std::vector<double> v((size_t)h->nvec);
    CK(cudaMemcpyAsync(v.data(), h->d_vec, sizeof(double) * h->nvec, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    unpack(h, v.data(), out, gr, Sk, nrho);
    return PIGS_OK;
}
extern "C" int pigs_get_block_chain(pigs_handle h, int chain, pigs_block_result* out, double* gr, double* Sk, double* nrho) {
    MULTI_CHAIN(h, chain, pigs_get_block_chain(s_, lc_, out, gr, Sk, nrho));
    NEED(h); NEED_CHAIN(h, chain);
    const DevParams& P = h->P;
    std::vector<double> a((size_t)P.nacc), v((size_t)h->nvec);
    std::vector<long long> c(NCNT);
    CK(cudaMemcpyAsync(a.data(), h->d_acc + (size_t)chain * P.nacc, sizeof(double) * P.nacc, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(c.data(), h->d_cnt + (size_t)chain * NCNT, sizeof(long long) * NCNT, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    for (int i = 0; i < NE; ++i) v[i] = a[i];
    for (int i = 0; i < NCNT; ++i) v[NE + i] = (double)c[i];
    for (int i = NE; i < P.nacc; ++i) v[NCNT + i] = a[i];
    unpack(h, v.data(), out, gr, Sk, nrho);
    return PIGS_OK;
}
// per-chain results of the last block for chains [chain0, chain0 + n): ONE device->host copy of the accumulators
// and one of the counters (the statistics layer needs every chain's block; a getter per chain costs n round trips)
extern "C" int pigs_get_block_chains(pigs_handle h, int chain0, int n, pigs_block_result* out, double* gr, double* Sk, double* nrho) {
    if (h && !h->sub.empty()) {
        if (n < 0 || chain0 < 0 || chain0 + n > h->hp.n_chains) return fail(PIGS_E_ARG, "chain range out of bounds");
        const size_t ngr = h->P.Nbin, nsk = (size_t)h->P.dim * h->P.Nk, nnr = (size_t)h->P.Nbin * (h->P.Npw + 1);
        for (size_t k = 0; k < h->sub.size(); ++k) {
            const int lo = std::max(chain0, h->first[k]), hi = std::min(chain0 + n, h->first[k + 1]);
            if (lo >= hi) continue;
            const size_t o = (size_t)(lo - chain0);
            int rc = pigs_get_block_chains(h->sub[k], lo - h->first[k], hi - lo, out ? out + o : nullptr, gr ? gr + o * ngr : nullptr,
                                           Sk ? Sk + o * nsk : nullptr, nrho ? nrho + o * nnr : nullptr);
            if (rc) return rc;
        }
        return PIGS_OK;
    }
    NEED(h);
    if (n < 0 || chain0 < 0 || chain0 + n > h->hp.n_chains) return fail(PIGS_E_ARG, "chain range out of bounds");
    if (n == 0) return PIGS_OK;
    const DevParams& P = h->P;
    std::vector<double> a((size_t)n * P.nacc), v((size_t)h->nvec);
    std::vector<long long> c((size_t)n * NCNT);
    CK(cudaMemcpyAsync(a.data(), h->d_acc + (size_t)chain0 * P.nacc, sizeof(double) * a.size(), cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(c.data(), h->d_cnt + (size_t)chain0 * NCNT, sizeof(long long) * c.size(), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    const size_t ngr = P.Nbin, nsk = (size_t)P.dim * P.Nk, nnr = (size_t)P.Nbin * (P.Npw + 1);
    for (int k = 0; k < n; ++k) {
        const double* ak = a.data() + (size_t)k * P.nacc;
        for (int i = 0; i < NE; ++i) v[i] = ak[i];
        for (int i = 0; i < NCNT; ++i) v[NE + i] = (double)c[(size_t)k * NCNT + i];
        for (int i = NE; i < P.nacc; ++i) v[NCNT + i] = ak[i];
        unpack(h, v.data(), out ? out + k : nullptr, gr ? gr + k * ngr : nullptr, Sk ? Sk + k * nsk : nullptr, nrho ? nrho + k * nnr : nullptr);
    }
    return PIGS_OK;
}
extern "C" int pigs_block_vector(pigs_handle h, double** dev_ptr, int* n) {
    if (h && !h->sub.empty()) return fail(PIGS_E_ARG, "pigs_block_vector: a multi-GPU handle reduces by itself (pigs_get_block)");
    NEED(h);
    if (dev_ptr) *dev_ptr = h->d_vec;
    if (n) *n = h->nvec;
    return PIGS_OK;
}
extern "C" int pigs_unpack_block_vector(pigs_handle h, const double* vec, pigs_block_result* out, double* gr, double* Sk, double* nrho) {
    if (!h || !vec) return fail(PIGS_E_ARG, "null argument");
    if (!h->sub.empty()) h = h->sub[0];
    unpack(h, vec, out, gr, Sk, nrho);
    return PIGS_OK;
}

// ---- unit API ----------------------------------------------------------------------------------
extern "C" int pigs_move(pigs_handle h, int move, int ip, int half, int32_t* accepted, int32_t* aux) {
    if (h && !h->sub.empty()) {
        for (size_t k = 0; k < h->sub.size(); ++k) {
            int rc = pigs_move(h->sub[k], move, ip, half, accepted ? accepted + h->first[k] : nullptr, aux ? aux + h->first[k] : nullptr);
            if (rc) return rc;
        }
        return PIGS_OK;
    }
    NEED(h);
    if (!h->tables_set) return fail(PIGS_E_STATE, "pigs_set_tables has not been called");
    if (move < 0 || move > 13) return fail(PIGS_E_ARG, "unknown move");
    if (ip < 1 || ip > h->hp.Np) return fail(PIGS_E_ARG, "ip out of range");
    if (move >= 7 && move <= 10 && half != 1 && half != 2) return fail(PIGS_E_ARG, "half must be 1 or 2");
    SweepArgs A = base_args(h);
    A.op = OP_MOVE; A.move = move; A.ip0 = ip - 1; A.half = half;
    A.accepted = h->d_iout; A.aux = h->d_iout + h->hp.n_chains;
    int rc = launch(h, A);
    if (rc) return rc;
    if (accepted) CK(cudaMemcpyAsync(accepted, h->d_iout, sizeof(int) * h->hp.n_chains, cudaMemcpyDeviceToHost, h->st));
    if (aux) CK(cudaMemcpyAsync(aux, h->d_iout + h->hp.n_chains, sizeof(int) * h->hp.n_chains, cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return PIGS_OK;
}

// host AoS [n][Np][dim] -> device blocked SoA [n][NpS/32][3][32] (unit calls are test-sized; transposed on the host)
static void to_soa(const DevParams& P, long long nslice, const double* aos, std::vector<double>& soa) {
    soa.assign((size_t)nslice * 3 * P.NpS, 0.0);
    for (long long s = 0; s < nslice; ++s)
        for (int ip = 0; ip < P.Np; ++ip)
            for (int k = 0; k < P.dim; ++k) soa[(size_t)s * 3 * P.NpS + pidx(ip) + 32 * k] = aos[((size_t)s * P.Np + ip) * P.dim + k];
}
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};

extern "C" int pigs_update_action(pigs_handle h, int n, const double* R, const int32_t* ip, const int32_t* ib,
                                  const double* xnew, const double* xold, double* DeltaS) {
    MULTI(h, pigs_update_action(h->sub[0], n, R, ip, ib, xnew, xold, DeltaS));
    NEED(h);
    if (!h->tables_set) return fail(PIGS_E_STATE, "pigs_set_tables has not been called");
    if (n < 0) return fail(PIGS_E_ARG, "bad argument");
    if (n == 0) return PIGS_OK;                    // an empty request: nothing to read, nothing to launch
    if (!R || !ip || !ib || !xnew || !xold || !DeltaS) return fail(PIGS_E_ARG, "bad argument");
    const DevParams& P = h->P;
    for (int i = 0; i < n; ++i)
        if (ip[i] < 1 || ip[i] > P.Np || ib[i] < 0 || ib[i] > 2 * P.Nb) return fail(PIGS_E_ARG, "ip/ib out of range");
    std::vector<double> soa;
    to_soa(P, n, R, soa);
    DevBuf dR, dip, dib, dxn, dxo, dS;
    CK(cudaMalloc(&dR.p, soa.size() * sizeof(double))); CK(cudaMalloc(&dip.p, n * sizeof(int))); CK(cudaMalloc(&dib.p, n * sizeof(int)));
    CK(cudaMalloc(&dxn.p, (size_t)n * P.dim * sizeof(double))); CK(cudaMalloc(&dxo.p, (size_t)n * P.dim * sizeof(double)));
    CK(cudaMalloc(&dS.p, n * sizeof(double)));
    CK(cudaMemcpyAsync(dR.p, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(dip.p, ip, n * sizeof(int), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(dib.p, ib, n * sizeof(int), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(dxn.p, xnew, (size_t)n * P.dim * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(dxo.p, xold, (size_t)n * P.dim * sizeof(double), cudaMemcpyHostToDevice, h->st));
    { std::lock_guard<std::mutex> lk(g_launch_mutex);
    int rc = guard_enter(h);
    if (rc) return rc;
    // default: global-memory tables and the reference's roundings; a handle created with table_mode = 2 EXPLICITLY runs the
    // production (Philox) instance of the pair loop instead -- shared-memory tables, fused r^2, rint minimum image
    CK(launch_update_action(P.trap, h->hp.table_mode == 2, P, n, (const double*)dR.p, (const int*)dip.p, (const int*)dib.p, (const double*)dxn.p,
                            (const double*)dxo.p, (double*)dS.p, h->st));
    rc = guard_leave(h);
    if (rc) return rc; }
    h->launches += 1;
    CK(cudaMemcpyAsync(DeltaS, dS.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return PIGS_OK;
}

static int unit_call(pigs_ctx* h, int op, int n, const std::vector<double>& in, size_t out_per, const double* out_init, double* out) {
    DevBuf din, dout;
    CK(cudaMalloc(&din.p, in.size() * sizeof(double)));
    CK(cudaMalloc(&dout.p, (size_t)n * out_per * sizeof(double)));
    CK(cudaMemcpyAsync(din.p, in.data(), in.size() * sizeof(double), cudaMemcpyHostToDevice, h->st));
    if (out_init) CK(cudaMemcpyAsync(dout.p, out_init, (size_t)n * out_per * sizeof(double), cudaMemcpyHostToDevice, h->st));
    UnitArgs A; A.op = op; A.n = n; A.in = (const double*)din.p; A.out = (double*)dout.p;
    { std::lock_guard<std::mutex> lk(g_launch_mutex);
    int rc = guard_enter(h);
    if (rc) return rc;
    CK(launch_unit(h->P.trap, h->P, A, h->st));
    rc = guard_leave(h);
    if (rc) return rc; }
    h->launches += 1;
    CK(cudaMemcpyAsync(out, dout.p, (size_t)n * out_per * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return PIGS_OK;
}
extern "C" int pigs_local_energy(pigs_handle h, int n, const double* R, double* E, double* Kin, double* Pot) {
    MULTI(h, pigs_local_energy(h->sub[0], n, R, E, Kin, Pot));
    NEED(h);
    if (!h->tables_set) return fail(PIGS_E_STATE, "pigs_set_tables has not been called");
    if (n < 0) return fail(PIGS_E_ARG, "bad argument");
    if (n == 0) return PIGS_OK;                    // an empty request: nothing to read, nothing to launch
    if (!R) return fail(PIGS_E_ARG, "bad argument");
    std::vector<double> soa, out((size_t)3 * n);
    to_soa(h->P, n, R, soa);
    int rc = unit_call(h, U_LOCAL_ENERGY, n, soa, 3, nullptr, out.data());
    if (rc) return rc;
    for (int i = 0; i < n; ++i) { if (E) E[i] = out[3 * i]; if (Kin) Kin[i] = out[3 * i + 1]; if (Pot) Pot[i] = out[3 * i + 2]; }
    return PIGS_OK;
}
extern "C" int pigs_therm_energy(pigs_handle h, int n, const double* Path, double* E, double* Ec, double* Ep) {
    MULTI(h, pigs_therm_energy(h->sub[0], n, Path, E, Ec, Ep));
    NEED(h);
    if (!h->tables_set) return fail(PIGS_E_STATE, "pigs_set_tables has not been called");
    if (n < 0) return fail(PIGS_E_ARG, "bad argument");
    if (n == 0) return PIGS_OK;                    // an empty request: nothing to read, nothing to launch
    if (!Path) return fail(PIGS_E_ARG, "bad argument");
    std::vector<double> soa, out((size_t)3 * n);
    to_soa(h->P, (long long)n * h->P.S, Path, soa);
    int rc = unit_call(h, U_THERM_ENERGY, n, soa, 3, nullptr, out.data());
    if (rc) return rc;
    for (int i = 0; i < n; ++i) { if (E) E[i] = out[3 * i]; if (Ec) Ec[i] = out[3 * i + 1]; if (Ep) Ep[i] = out[3 * i + 2]; }
    return PIGS_OK;
}
extern "C" int pigs_pair_correlation(pigs_handle h, int n, const double* R, double* gr) {
    MULTI(h, pigs_pair_correlation(h->sub[0], n, R, gr));
    NEED(h);
    if (h->P.trap) return fail(PIGS_E_ARG, "PairCorrelation is not defined in trap mode (vpi.f90:466)");
    if (n < 0) return fail(PIGS_E_ARG, "bad argument");
    if (n == 0) return PIGS_OK;                    // an empty request: nothing to read, nothing to launch
    if (!R || !gr) return fail(PIGS_E_ARG, "bad argument");
    std::vector<double> soa;
    to_soa(h->P, n, R, soa);
    return unit_call(h, U_PAIR_CORR, n, soa, h->P.Nbin, gr, gr);
}
extern "C" int pigs_structure_factor(pigs_handle h, int n, const double* R, double* Sk) {
    MULTI(h, pigs_structure_factor(h->sub[0], n, R, Sk));
    NEED(h);
    if (h->P.trap) return fail(PIGS_E_ARG, "StructureFactor is not defined in trap mode (vpi.f90:466)");
    if (n < 0) return fail(PIGS_E_ARG, "bad argument");
    if (n == 0) return PIGS_OK;                    // an empty request: nothing to read, nothing to launch
    if (!R || !Sk) return fail(PIGS_E_ARG, "bad argument");
    std::vector<double> soa;
    to_soa(h->P, n, R, soa);
    return unit_call(h, U_SOFK, n, soa, (size_t)h->P.Nk * h->P.dim, Sk, Sk);
}
extern "C" int pigs_obdm(pigs_handle h, int n, const double* xend, double* nrho) {
    MULTI(h, pigs_obdm(h->sub[0], n, xend, nrho));
    NEED(h);
    if (h->P.trap) return fail(PIGS_E_ARG, "OBDM is not defined in trap mode (vpi.f90:400)");
    if (n < 0) return fail(PIGS_E_ARG, "bad argument");
    if (n == 0) return PIGS_OK;                    // an empty request: nothing to read, nothing to launch
    if (!xend || !nrho) return fail(PIGS_E_ARG, "bad argument");
    const int dim = h->P.dim;
    std::vector<double> xe((size_t)n * 6, 0.0);
    for (int c = 0; c < n; ++c) for (int j = 0; j < 2; ++j) for (int k = 0; k < dim; ++k) xe[(size_t)c * 6 + j * 3 + k] = xend[((size_t)c * 2 + j) * dim + k];
    return unit_call(h, U_OBDM, n, xe, (size_t)h->P.Nbin * (h->P.Npw + 1), nrho, nrho);
}

extern "C" int pigs_measure_fp64_peak(int device, double* tflops) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(PIGS_E_CUDA, "no CUDA device");
    if (device < 0 || device >= ndev) return fail(PIGS_E_ARG, "bad device ordinal");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double* sink = nullptr;
    CK(cudaMalloc((void**)&sink, sizeof(double)));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(a, 0));
        CK(launch_dfma_peak(blocks, threads, iters, sink, 0));
        CK(cudaEventRecord(b, 0));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        double fl = 2.0 * 64.0 * (double)iters * (double)blocks * threads;
        double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
    if (tflops) *tflops = best;
    return PIGS_OK;
}
