// ref_glue.cpp -- C interface over oracle/_ref/ref_gen.cpp, the MACHINE TRANSLATION of the reference's Fortran
// sources (oracle/f90toc/f90toc.py).  TEST INFRASTRUCTURE ONLY: it pins the hand-written oracle
// (oracle/pigs_oracle.cpp) to the reference text -- tests/test_ref_pin.py demands bit equality between the two.
//
// Hand-written here is only what the translation cannot carry: the five I/O procedures of the reference
// (ReadParameters, ReadSystemParameters: namelists on stdin -> a parameter struct; mtsavef, mtgetf, CheckPoint:
// files -> no-ops) and thin extern "C" wrappers that call the translated procedures with caller arrays.
// The translated code keeps the reference's module globals, so one configuration is live at a time.
#include "_ref/ref_gen.cpp"

#include <cstdint>

extern "C" {
typedef struct ref_params {
    int32_t dim, Np;
    double density;
    int32_t crystal, trap;
    double dt;
    int32_t Nb, seed;
    double delta_cm;
    int32_t CMFreq, sampling;            /* sampling: 0 'sta', 1 'bis' */
    int32_t Lstag, Nlev, Nstag, Nblock, Nstep, Nbin, Nk;
    int32_t swapping;
    double CWorm;
    int32_t Nobdm, Npw, Nmax, wf_table, v_table;
    double Rm;
    double a_ho[3];
} ref_params;
}

static ref_params g_par;

// vpi_mod.f90:14-80: defaults, then the namelists &system &samp &obdm &wavefun (here: the struct), then
// ReadSystemParameters.  Every value the reference reads from vpi.in comes from g_par.
void readparameters_(bool& resume_, bool& crystal_, bool& wf_table_, bool& v_table_, bool& swapping_, bool& trap_,
                     std::string& sampling_, double& density_, double& dt_, double& delta_cm_, double& rm_, int& dim_, int& np_,
                     int& nb_, int& seed_, int& cmfreq_, int& lstag_, int& nlev_, int& nstag_, int& nmax_, int& nobdm_,
                     int& nblock_, int& nstep_, int& nbin_, int& nk_) {
    dim_ = g_par.dim; np_ = g_par.Np; density_ = g_par.density; crystal_ = g_par.crystal != 0; trap_ = g_par.trap != 0;
    resume_ = false; dt_ = g_par.dt; nb_ = g_par.Nb; seed_ = g_par.seed; delta_cm_ = g_par.delta_cm; cmfreq_ = g_par.CMFreq;
    sampling_ = g_par.sampling ? "bis" : "sta";
    lstag_ = g_par.Lstag; nlev_ = g_par.Nlev; nstag_ = g_par.Nstag; nblock_ = g_par.Nblock; nstep_ = g_par.Nstep;
    nbin_ = g_par.Nbin; nk_ = g_par.Nk;
    swapping_ = g_par.swapping != 0; ::cworm_ = g_par.CWorm; nobdm_ = g_par.Nobdm; ::npw_ = g_par.Npw;
    nmax_ = g_par.Nmax; wf_table_ = g_par.wf_table != 0; v_table_ = g_par.v_table != 0;
    (void)rm_;
    readsystemparameters_(trap_);
}
// system_mod.f90:15-34: &extpot a_ho (trap only, allocates a_ho(dim)) and &jastrow Rm
void readsystemparameters_(bool& trap_) {
    if (trap_) {
        a_ho__l1 = 1; a_ho__n1 = ::dim_;
        a_ho__p = (double*)f90rt::alloc(sizeof(double) * (size_t)::dim_);
        for (int k = 0; k < ::dim_; ++k) a_ho__p[k] = g_par.a_ho[k];
    }
    ::rm_ = g_par.Rm;
}
void mtsavef_(std::string, std::string) {}
void mtgetf_(std::string, std::string) { throw std::runtime_error("resume is not wired in the pin"); }
void checkpoint_(bool&, double*, double*, bool&, int&) {}

extern "C" {

// ---- the whole program: `./vpi < vpi.in` (vpi.f90), output files captured as numbers
int ref_run_program(const ref_params* p) {
    g_par = *p;
    f90rt::files().clear();
    mti_ = 625;
    try { vpi_(); } catch (const std::exception& e) { std::fprintf(stderr, "ref_run_program: %s\n", e.what()); return -1; }
    return 0;
}
// UpdateAction calls so far (the metric's unit: one bead-update; counter injected by f90toc --count)
long long ref_bead_updates(void) { return calls_updateaction_; }
// queue one record for `read(unit,*)` statements (config_ini.in of crystal = .true.)
void ref_queue_read(int unit, const double* v, int n) { f90rt::unit(unit).to_read.emplace_back(v, v + n); }
void ref_clear_reads(int unit) { f90rt::unit(unit).to_read.clear(); f90rt::unit(unit).rpos = 0; }
// records of one captured file, flattened row by row; returns the number of records, *ncol = widest record
int ref_file(const char* name, double* buf, int cap, int* ncol) {
    auto& recs = f90rt::files().get(name);
    int w = 0, n = 0;
    for (auto& r : recs) if ((int)r.v.size() > w) w = (int)r.v.size();
    if (ncol) *ncol = w;
    for (auto& r : recs) {
        if (buf && (n + 1) * w <= cap) {
            for (int j = 0; j < w; ++j) buf[n * w + j] = j < (int)r.v.size() ? r.v[j] : 0.0;
        }
        ++n;
    }
    return n;
}
// globals as the program left them (vpi.f90:80-128)
void ref_geometry(double* Lbox3, double* rcut, double* dr, double* rbin) {
    for (int k = 0; k < 3; ++k) Lbox3[k] = (lbox__p && k < dim_) ? lbox__p[k] : 0.0;
    *rcut = rcut_; *dr = dr_; *rbin = rbin_;
}
void ref_tables(double* LogWF, double* VTable) {      // JastrowTable / PotentialTable (vpi_mod.f90:84-145)
    jastrowtable_(rcut_, LogWF);
    potentialtable_(rcut_, VTable);
}

// ---- leaves
double ref_interpolate(int opt, int N, double dx, double* F, double x) { return interpolate_(opt, N, dx, F, x); }
double ref_potential(double r) { double xij[3] = {r, 0, 0}; return potential_(xij, r); }
double ref_logpsi(int opt, double Rm, double r) { return logpsi_(opt, Rm, r); }
double ref_trappsi(int opt, double a, double x) { return trappsi_(opt, a, x); }
double ref_trappot(int opt, double a, double x) { return trappot_(opt, a, x); }
double ref_green(int opt, int ib, double dt, double Pot, double F2) { return greenfunction_(opt, ib, dt, Pot, F2); }
double ref_r8_gamma(double x) { return r8_gamma_(x); }
void ref_minimum_image(double* xij, double* r2) { minimumimage_(xij, *r2); }
double ref_boundary(int k, double x) { boundaryconditions_(k, x); return x; }
void ref_sgrnd(int seed) { sgrnd_(seed); }
double ref_grnd(void) { return grnd_(); }
double ref_rangauss(void) { double a, b; rangauss_(1.0, 0.0, a, b); return a; }
void ref_get_mt(uint32_t* mt624, int32_t* mti) { for (int i = 0; i < 624; ++i) mt624[i] = (uint32_t)mt__p[i]; *mti = mti_; }
void ref_set_mt(const uint32_t* mt624, int32_t mti) { for (int i = 0; i < 624; ++i) mt__p[i] = (int)mt624[i]; mti_ = mti; }

// ---- action and estimators on caller arrays (Fortran layouts)
double ref_update_action(double* LogWF, double* VTable, double* Path, int ip, int ib, double* xnew, double* xold, double dt, int trap) {
    double dS = 0.0;
    updateaction_(trap != 0, LogWF, VTable, Path, ip, ib, xnew, xold, dt, dS);
    return dS;
}
void ref_local_energy(double* LogWF, double* VTable, double* R, int trap, double* E, double* Kin, double* Pot) {
    localenergy_(trap != 0, LogWF, VTable, R, *E, *Kin, *Pot);
}
void ref_therm_energy(double* VTable, double* Path, double dt, int trap, double* E, double* Ec, double* Ep) {
    thermenergy_(trap != 0, VTable, Path, dt, *E, *Ec, *Ep);
}
void ref_pair_correlation(double* R, double* gr) { paircorrelation_(R, gr); }
void ref_structure_factor(int Nk, double* R, double* Sk) { structurefactor_(Nk, R, Sk); }
void ref_obdm(double* xend, double* nrho) { obdm_(xend, nrho); }
void ref_normalize_gr(double density, int ngr, double* gr) { normalizegr_(density, ngr, gr); }
void ref_normalize_sk(int Nk, int ngr, double* Sk) { normalizesk_(Nk, ngr, Sk); }
void ref_normalize_nr(double density, double zconf, int Nobdm, double* nrho) { normalizenr_(density, zconf, Nobdm, nrho); }
double ref_var(int n, double s, double s2) { return var_(n, s, s2); }

// ---- the 14 moves on caller arrays; numbering of oracle/pigs_oracle.h.  Returns the new value of `accepted`.
int ref_move(int move, int trap, double* LogWF, double* VTable, double dt, double delta_cm, double density, int Lstag, int Nlev,
             int ip, int half, double* Path, double* xend, int* isopen, int* aux) {
    const bool tr = trap != 0;
    int acc = 0;
    bool open = *isopen != 0, flag = false;
    switch (move) {
    case 0: translatechain_(tr, delta_cm, LogWF, VTable, dt, ip, Path, acc); break;
    case 1: staging_(tr, LogWF, VTable, dt, Lstag, ip, Path, acc); break;
    case 2: movehead_(tr, LogWF, VTable, dt, Lstag, ip, Path, acc); break;
    case 3: movetail_(tr, LogWF, VTable, dt, Lstag, ip, Path, acc); break;
    case 4: bisection_(tr, LogWF, VTable, dt, Nlev, ip, Path, acc); break;
    case 5: moveheadbisection_(tr, LogWF, VTable, dt, Nlev, ip, Path, acc); break;
    case 6: movetailbisection_(tr, LogWF, VTable, dt, Nlev, ip, Path, acc); break;
    case 7: translatehalfchain_(tr, half, delta_cm, LogWF, VTable, dt, ip, Path, xend, acc); break;
    case 8: staginghalfchain_(tr, half, LogWF, VTable, dt, Lstag, ip, Path, xend, acc); break;
    case 9: moveheadhalfchain_(tr, half, LogWF, VTable, dt, Lstag, ip, Path, xend, acc); break;
    case 10: movetailhalfchain_(tr, half, LogWF, VTable, dt, Lstag, ip, Path, xend, acc); break;
    case 11: openchain_(tr, LogWF, VTable, density, dt, Lstag, ip, Path, xend, open, acc, flag); break;
    case 12: closechain_(tr, LogWF, VTable, density, dt, Lstag, ip, Path, xend, open, acc, flag); break;
    case 13: {
        int ipar = 0;
        bool sw = false;
        swap_(tr, LogWF, VTable, dt, Lstag, ip, Path, xend, acc, ipar, sw);
        if (aux) *aux = sw ? ipar : 0;
        break;
    }
    default: return -1;
    }
    *isopen = open ? 1 : 0;
    return acc;
}

}  // extern "C"
