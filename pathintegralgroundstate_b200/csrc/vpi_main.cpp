// vpi_cuda -- `program vpi` (vpi.f90:76-653) on top of libpigs_cuda.
//
//     ./vpi_cuda < vpi.in          (the reference's command line: ./vpi < vpi.in)
//
// The reference driver's outer layers -- ReadParameters (vpi_mod.f90:14-80),
// geometry (vpi.f90:80-128), JastrowTable/PotentialTable (vpi_mod.f90:84-145),
// init incl. resume (vpi_mod.f90:149-259), the block loop with NormalizeAv /
// NormalizeGr / NormalizeSk / NormalizeNr / Var (vpi.f90:244-545,
// sample_mod.f90:598-732,921-932), the output files e_vpi.out, et_vpi.out
// ('(5g20.10e3)', vpi.f90:517-518), gr_vpi.out, sk_vpi.out, nr_vpi.out
// ('(20g20.10e3)', sample_mod.f90:794-870), jastrow.out, potential.out, fort.99
// (vpi.f90:590-592), checkpoint.dat (vpi_mod.f90:263-309) and rand_state
// (random_mod.f90:125-158, gfortran unformatted records) -- with the step loop
// (vpi.f90:297-475) replaced by pigs_run_block.  Compiled host code over the C
// ABI of include/pigs_cuda.h only; the Python module
// pathintegralgroundstate_b200/driver.py is the same program for scripting and
// the two are tested against each other file by file.
//
// With n_chains > 1 (&cuda group or --chains) every sum the reference keeps per
// block is the sum over chains: block averages become averages over chains x
// diagonal steps.
//
// Test hooks without a GPU: --tables-only (parse, geometry, jastrow.out,
// potential.out, then stop) and --format-test (reads "w d e x" lines from stdin,
// prints Gw.dEe).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "pigs_cuda.h"

namespace {

// ------------------------------------------------------------------ Fortran edit descriptors
std::string rjust(const std::string& s, int w) { return (int)s.size() >= w ? s : std::string(w - s.size(), ' ') + s; }
std::string fmt_f(double x, int d) {
    char buf[512];
    snprintf(buf, sizeof buf, "%.*f", d, x);
    return buf;
}
// Fw.d
std::string fortran_f(double x, int w, int d) {
    if (std::isnan(x)) return rjust("NaN", w);             // gfortran's spelling of non-finite values
    if (std::isinf(x)) {
        const std::string t = x > 0 ? (w >= 8 ? "Infinity" : "Inf") : (w >= 9 ? "-Infinity" : "-Inf");
        return (int)t.size() <= w ? rjust(t, w) : std::string(w, '*');
    }
    std::string s = fmt_f(x, d);
    if (d == 0) s += ".";                                  // Fortran always prints the decimal point
    if (s.rfind("0.", 0) == 0 && (int)s.size() > w) s = s.substr(1);
    else if (s.rfind("-0.", 0) == 0 && (int)s.size() > w) s = "-" + s.substr(2);
    return (int)s.size() <= w ? rjust(s, w) : std::string(w, '*');
}
// a / 10^ex without overflow or underflow of the power at the ends of the double range
double scaled(double a, int ex) {
    if (ex > 300) return a / 1e300 / std::pow(10.0, ex - 300);
    if (ex < -300) return a * 1e300 / std::pow(10.0, ex + 300);
    return a / std::pow(10.0, ex);
}
// Ew.dEe with scale factor 0: 0.dddddE+eee
std::string fortran_e(double x, int w, int d, int e) {
    if (std::isnan(x)) return rjust("NaN", w);
    if (std::isinf(x)) return rjust(x > 0 ? "Infinity" : "-Infinity", w);
    std::string sgn, mant;
    int ex = 0;
    if (x == 0.0) {
        mant = "0." + std::string(d, '0');
        sgn = std::signbit(x) ? "-" : "";
    } else {
        sgn = x < 0 ? "-" : "";
        const double a = std::fabs(x);
        ex = (int)std::floor(std::log10(a)) + 1;
        std::string ms = fmt_f(scaled(a, ex), d);
        if (ms.rfind("1.", 0) == 0) {                      // rounding carried into the next decade
            ex += 1;
            ms = fmt_f(scaled(a, ex), d);
        }
        mant = ms;
    }
    char eb[32];
    snprintf(eb, sizeof eb, "E%c%0*d", ex >= 0 ? '+' : '-', e, std::abs(ex));
    std::string s = sgn + mant + eb;
    if ((int)s.size() > w) {
        const size_t p = s.find("0.");
        const std::string stripped = (s[0] == '-') ? s.substr(1) : s;
        if (stripped.rfind("0.", 0) == 0 && p != std::string::npos) s.replace(p, 2, ".");
    }
    return (int)s.size() <= w ? rjust(s, w) : std::string(w, '*');
}
// Gw.dEe (Fortran 2003 10.6.4.1.2): F editing with e+2 trailing blanks when the
// magnitude fits d significant digits, E editing otherwise
std::string fortran_g(double x, int w = 20, int d = 10, int e = 3) {
    const int n = e + 2;
    if (!std::isfinite(x)) return fortran_e(x, w, d, e);
    const double a = std::fabs(x);
    if (a == 0.0) return fortran_f(x, w - n, d - 1) + std::string(n, ' ');
    if (a < 0.1 - 0.5 * std::pow(10.0, -d - 1) || a >= std::pow(10.0, d) - 0.5) return fortran_e(x, w, d, e);
    int k = (int)std::floor(std::log10(a)) + 1;
    if (a >= std::pow(10.0, k) - 0.5 * std::pow(10.0, k - d)) k += 1;
    else if (a < std::pow(10.0, k - 1) - 0.5 * std::pow(10.0, k - d - 1)) k -= 1;
    k = std::min(std::max(k, 0), d);
    return fortran_f(x, w - n, d - k) + std::string(n, ' ');
}
std::string g_line(const std::vector<double>& v, int w = 20, int d = 10, int e = 3) {
    std::string s;
    for (double x : v) s += fortran_g(x, w, d, e);
    return s + "\n";
}
// Var (sample_mod.f90:921-932)
double var(long long nitem, double s, double s2) {
    const double r = (s2 - s * s) / (double)nitem;
    return r >= 0 ? std::sqrt(r) : std::nan("");
}
double f32(double x) { return (double)(float)x; }      // a default-kind real() cast

// ------------------------------------------------------------------ vpi.in
struct Value {
    std::vector<std::string> toks;
    bool has() const { return !toks.empty(); }
};
typedef std::map<std::string, std::map<std::string, Value>> Namelists;       // lower-case group -> lower-case name

std::string lower(std::string s) {
    for (auto& c : s) c = (char)std::tolower((unsigned char)c);
    return s;
}
bool is_ident_char(char c) { return std::isalnum((unsigned char)c) || c == '_'; }

// All `&group ... /` blocks; '!' outside quotes starts a comment; names are case-insensitive.
Namelists parse_namelists(const std::string& text) {
    std::string body;
    {
        std::istringstream in(text);
        std::string ln;
        while (std::getline(in, ln)) {
            char q = 0;
            std::string buf;
            for (char ch : ln) {
                if (q) { buf += ch; if (ch == q) q = 0; }
                else if (ch == '\'' || ch == '"') { q = ch; buf += ch; }
                else if (ch == '!') break;
                else buf += ch;
            }
            body += buf + "\n";
        }
    }
    Namelists out;
    size_t i = 0;
    const size_t n = body.size();
    while (i < n) {
        if (body[i] != '&') { ++i; continue; }
        ++i;
        while (i < n && std::isspace((unsigned char)body[i])) ++i;
        std::string g;
        while (i < n && is_ident_char(body[i])) g += body[i++];
        g = lower(g);
        if (g == "end") continue;
        // group body up to an unquoted '/' or "&end"
        std::string gb;
        char q = 0;
        while (i < n) {
            const char ch = body[i];
            if (q) { gb += ch; if (ch == q) q = 0; ++i; continue; }
            if (ch == '\'' || ch == '"') { q = ch; gb += ch; ++i; continue; }
            if (ch == '/') { ++i; break; }
            if (ch == '&') { if (lower(body.substr(i + 1, 3)) == "end") i += 4; break; }
            gb += ch;
            ++i;
        }
        // name [ (index) ] = values ... up to the next "name ="
        std::map<std::string, Value> vals;
        size_t j = 0;
        const size_t m = gb.size();
        std::string cur;
        std::vector<std::string> pend;          // tokens since the last '='
        auto flush = [&](const std::string& next_name) {
            if (!cur.empty()) vals[cur].toks = pend;
            pend.clear();
            cur = next_name;
        };
        while (j < m) {
            const char ch = gb[j];
            if (std::isspace((unsigned char)ch) || ch == ',') { ++j; continue; }
            std::string tok;
            if (ch == '\'' || ch == '"') {
                const char qq = ch;
                ++j;
                while (j < m && gb[j] != qq) tok += gb[j++];
                ++j;
                pend.push_back(tok);
                continue;
            }
            while (j < m && !std::isspace((unsigned char)gb[j]) && gb[j] != ',' && gb[j] != '=' && gb[j] != '(') tok += gb[j++];
            // look ahead: [ (..) ] '=' makes tok a name
            size_t k = j;
            while (k < m && std::isspace((unsigned char)gb[k])) ++k;
            if (k < m && gb[k] == '(') { while (k < m && gb[k] != ')') ++k; if (k < m) ++k; while (k < m && std::isspace((unsigned char)gb[k])) ++k; }
            if (k < m && gb[k] == '=' && !tok.empty() && (std::isalpha((unsigned char)tok[0]) || tok[0] == '_')) {
                flush(lower(tok));
                j = k + 1;
            } else if (!tok.empty()) {
                pend.push_back(tok);
            } else {
                ++j;
            }
        }
        flush("");
        out[g] = vals;
    }
    return out;
}
double to_double(std::string t) {
    for (auto& c : t) if (c == 'd' || c == 'D') c = 'e';
    return std::strtod(t.c_str(), nullptr);
}
bool to_bool(const std::string& t) {
    std::string s = lower(t);
    while (!s.empty() && s[0] == '.') s = s.substr(1);
    return !s.empty() && s[0] == 't';
}

struct Config {
    // &system
    int dim = 0, Np = 0;
    double density = 0;
    bool crystal = false, trap = false;
    // &samp
    bool resume = false;
    double dt = 0, delta_cm = 0;
    int Nb = 0, seed = 1982, CMFreq = 0, Lstag = 2, Nlev = 1, Nstag = 0, Nblock = 0, Nstep = 0, Nbin = 0, Nk = 0;
    std::string sampling;
    // &obdm
    bool swapping = false;
    double CWorm = 0;
    int Nobdm = 0, Npw = 0;
    // &wavefun / &jastrow / &extpot
    int Nmax = 10000;
    bool wf_table = false, v_table = false;
    double Rm = 0;
    std::vector<double> a_ho;
    // &cuda (not in the reference)
    int n_chains = 1, threads_per_chain = 0, table_mode = -1, schedule = -1, gpus = 1, checkpoint_every = 1;
    std::string rng, action;
    // crystal: config_ini.in line 2
    std::vector<double> Lbox_in;
};

bool file_exists(const std::string& p) {
    FILE* f = fopen(p.c_str(), "rb");
    if (f) fclose(f);
    return f != nullptr;
}
void die(const std::string& msg);
void die(const std::string& msg) {
    std::cerr << "vpi_cuda: " << msg << std::endl;
    std::exit(2);
}

Config read_vpi_in(const std::string& text) {
    Namelists nl = parse_namelists(text);
    Config c;
    std::vector<std::string> missing;
    auto get = [&](const char* g, const char* name) -> const Value* {
        auto gi = nl.find(g);
        if (gi == nl.end()) return nullptr;
        auto vi = gi->second.find(lower(name));
        return (vi == gi->second.end() || !vi->second.has()) ? nullptr : &vi->second;
    };
    auto I = [&](const char* g, const char* name, int& dst, bool required) {
        if (const Value* v = get(g, name)) dst = (int)std::llround(to_double(v->toks[0]));
        else if (required) missing.push_back(name);
    };
    auto D = [&](const char* g, const char* name, double& dst, bool required) {
        if (const Value* v = get(g, name)) dst = to_double(v->toks[0]);
        else if (required) missing.push_back(name);
    };
    auto B = [&](const char* g, const char* name, bool& dst) { if (const Value* v = get(g, name)) dst = to_bool(v->toks[0]); };
    I("system", "dim", c.dim, true); I("system", "Np", c.Np, true); D("system", "density", c.density, true);
    B("system", "crystal", c.crystal); B("system", "trap", c.trap);
    B("samp", "resume", c.resume); D("samp", "dt", c.dt, true); I("samp", "Nb", c.Nb, true); I("samp", "seed", c.seed, false);
    D("samp", "delta_cm", c.delta_cm, true); I("samp", "CMFreq", c.CMFreq, true);
    if (const Value* v = get("samp", "sampling")) c.sampling = v->toks[0].substr(0, 3); else missing.push_back("sampling");
    I("samp", "Lstag", c.Lstag, false); I("samp", "Nlev", c.Nlev, false); I("samp", "Nstag", c.Nstag, true);
    I("samp", "Nblock", c.Nblock, true); I("samp", "Nstep", c.Nstep, true); I("samp", "Nbin", c.Nbin, true); I("samp", "Nk", c.Nk, true);
    B("obdm", "swapping", c.swapping); D("obdm", "CWorm", c.CWorm, false); I("obdm", "Nobdm", c.Nobdm, false); I("obdm", "Npw", c.Npw, false);
    I("wavefun", "Nmax", c.Nmax, false); B("wavefun", "wf_table", c.wf_table); B("wavefun", "v_table", c.v_table);
    D("jastrow", "Rm", c.Rm, true);
    if (c.trap) {
        const Value* v = get("extpot", "a_ho");
        if (!v) die("trap=T needs &extpot a_ho (system_mod.f90:24-28)");
        for (auto& t : v->toks) c.a_ho.push_back(to_double(t));
    }
    I("cuda", "n_chains", c.n_chains, false); I("cuda", "threads_per_chain", c.threads_per_chain, false);
    I("cuda", "table_mode", c.table_mode, false); I("cuda", "schedule", c.schedule, false);
    I("cuda", "gpus", c.gpus, false); I("cuda", "checkpoint_every", c.checkpoint_every, false);
    if (const Value* v = get("cuda", "rng")) c.rng = lower(v->toks[0]);
    if (const Value* v = get("cuda", "action")) c.action = lower(v->toks[0]);      // 'chin' | 'primitive' (global_mod.f90:48,67)
    if (!missing.empty()) {
        std::string s = "vpi.in lacks:";
        for (auto& m : missing) s += " " + m;
        die(s);
    }
    return c;
}

// ------------------------------------------------------------------ geometry (vpi.f90:80-128, vpi_mod.f90:94)
struct Geometry {
    double Lbox[3] = {0, 0, 0}, a_ho[3] = {1, 1, 1};
    double rcut = 0, rbin = 0, density = 0, delta_cm = 0, dr = 0, pi = 0;
};
Geometry derive_geometry(const Config& c) {
    Geometry g;
    g.pi = std::acos(-1.0);
    const int dim = c.dim;
    if (c.trap) {
        if ((int)c.a_ho.size() < dim) die("&extpot a_ho needs dim values");
        double rcut = 1.0, amin = c.a_ho[0];
        for (int k = 0; k < dim; ++k) { rcut = 3.0 * rcut * c.a_ho[k]; amin = std::min(amin, c.a_ho[k]); g.a_ho[k] = c.a_ho[k]; }
        g.density = f32(c.Np) / (std::pow(g.pi, 0.5 * dim) * rcut / std::tgamma(0.5 * dim + 1.0));
        rcut = std::pow(rcut, 1.0 / f32(dim));
        g.rcut = 10.0 * rcut;
        g.delta_cm = c.delta_cm * amin;
    } else {
        g.density = c.density;
        for (int k = 0; k < dim; ++k)
            g.Lbox[k] = c.crystal ? c.Lbox_in[k] : std::pow(f32(c.Np) / g.density, 1.0 / f32(dim));
        g.rcut = 0.5 * g.Lbox[0];
        for (int k = 1; k < dim; ++k) g.rcut = std::min(g.rcut, 0.5 * g.Lbox[k]);
        g.delta_cm = c.delta_cm / std::pow(g.density, 1.0 / f32(dim));
    }
    g.rbin = g.rcut / f32(c.Nbin);
    g.dr = g.rcut / f32(c.Nmax - 1);
    return g;
}

// ------------------------------------------------------------------ system_mod.f90
double aziz_hfdb(double r) {        // Aziz II HFD-B(HE), reduced units (system_mod.f90:136-182)
    const double E_0 = 10.948, rm = 2.963, A = 1.8443101e5, alpha = 10.43329537, beta = -2.27965105;
    const double C6 = 1.36745214, C8 = 0.42123807, C10 = 0.17473318, Dd = 1.4826;
    const double V0 = E_0 / 1.85505153154686;
    const double d = r * 2.556 / rm, d2 = d * d, d4 = d2 * d2, d6 = d4 * d2;
    const double t = Dd / d - 1.0;
    const double Hx = d <= Dd ? std::exp(-(t * t)) : 1.0;
    return V0 * (A * std::exp(-alpha * d + beta * d2) - (C6 + C8 / d2 + C10 / d4) * Hx / d6);
}
double aziz_hfdhe2(double r) {      // Aziz I HFDHE2 (the commented-out alternative, system_mod.f90:87-132)
    const double E_0 = 10.8, rm = 2.9673, A = 0.54485046e6, alpha = 13.353384;
    const double C6 = 1.3732412, C8 = 0.4253785, C10 = 0.1781, Dd = 1.241314;
    const double V0 = E_0 / 1.85505153154686;
    const double d = r * 2.556 / rm, d2 = d * d, d4 = d2 * d2, d6 = d4 * d2;
    const double t = Dd / d - 1.0;
    const double Hx = d <= Dd ? std::exp(-(t * t)) : 1.0;
    return V0 * (A * std::exp(-alpha * d) - (C6 + C8 / d2 + C10 / d4) * Hx / d6);
}
double mcmillan_logpsi(double r, double Rm) {      // LogPsi(0,Rm,r) (system_mod.f90:38-66)
    const double q = Rm / r, q2 = q * q;
    return -0.5 * ((q * q2) * q2);        // x**5 as gfortran expands it: (x * x^2) * x^2
}
// JastrowTable / PotentialTable (vpi_mod.f90:84-145): entry i holds f((i-1)*dr), pads F(0)=F(2), F(Nmax+1)=F(Nmax)
template <class F>
std::vector<double> make_table(F f, double dr, int Nmax) {
    std::vector<double> T(Nmax + 2, 0.0);
    for (int i = 1; i <= Nmax; ++i) T[i] = f((double)(i - 1) * dr);
    T[0] = T[2];
    T[Nmax + 1] = T[Nmax];
    return T;
}

// ------------------------------------------------------------------ checkpoint files
// CheckPoint (vpi_mod.f90:263-309): list-directed text, particle-major, bead-minor
void write_checkpoint(const std::string& path, bool trap, const std::vector<double>& P, const double* xend, int isopen, int iworm,
                      int S, int Np, int dim) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) die("cannot write " + path);
    fputs(trap ? " .True.\n" : " .False.\n", f);
    fputs(isopen ? " .True.\n" : " .False.\n", f);
    fprintf(f, " %11d\n", iworm);
    for (int ip = 0; ip < Np; ++ip)
        for (int ib = 0; ib < S; ++ib) {
            fputc(' ', f);
            for (int k = 0; k < dim; ++k) fprintf(f, "%s%24.16E", k ? " " : "", P[((size_t)ib * Np + ip) * dim + k]);
            fputc('\n', f);
        }
    fputs("\n\n", f);
    for (int j = 0; j < 2; ++j) {
        fputc(' ', f);
        for (int k = 0; k < dim; ++k) fprintf(f, "%s%24.16E", k ? " " : "", xend[j * dim + k]);
        fputc('\n', f);
    }
    fclose(f);
}
// init, resume branch (vpi_mod.f90:162-185)
void read_checkpoint(const std::string& path, int dim, int Np, int Nb, bool& trap, int& isopen, int& iworm, std::vector<double>& P,
                     double* xend) {
    std::ifstream in(path);
    if (!in) die("resume=T but " + path + " is missing");
    std::vector<std::string> t;
    std::string w;
    while (in >> w) t.push_back(w);
    const int S = 2 * Nb + 1;
    const size_t need = 3 + (size_t)Np * S * dim + 2 * dim;
    if (t.size() < need) die(path + " is too short");
    trap = to_bool(t[0]);
    isopen = to_bool(t[1]) ? 1 : 0;
    iworm = std::atoi(t[2].c_str());
    P.assign((size_t)S * Np * dim, 0.0);
    size_t q = 3;
    for (int ip = 0; ip < Np; ++ip)
        for (int ib = 0; ib < S; ++ib)
            for (int k = 0; k < dim; ++k) P[((size_t)ib * Np + ip) * dim + k] = to_double(t[q++]);
    for (int j = 0; j < 2 * dim; ++j) xend[j] = to_double(t[q++]);
}
// mtsavef(fname,'u') (random_mod.f90:125-158): two gfortran unformatted sequential records, APPENDED
void append_rand_state(const std::string& path, const uint32_t* mt, int32_t mti) {
    FILE* f = fopen(path.c_str(), "ab");
    if (!f) die("cannot write " + path);
    const int32_t four = 4, nb = 624 * 4;
    fwrite(&four, 4, 1, f); fwrite(&mti, 4, 1, f); fwrite(&four, 4, 1, f);
    fwrite(&nb, 4, 1, f); fwrite(mt, 4, 624, f); fwrite(&nb, 4, 1, f);
    fclose(f);
}
// mtgetf(fname,'u') (random_mod.f90:162-191): reads the FIRST (oldest) pair of records
void read_rand_state(const std::string& path, uint32_t* mt, int32_t& mti) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) die("resume=T but " + path + " is missing");
    int32_t h[4];
    if (fread(h, 4, 4, f) != 4 || h[0] != 4 || h[2] != 4 || h[3] != 624 * 4 || fread(mt, 4, 624, f) != 624) die(path + ": bad record");
    mti = h[1];
    fclose(f);
}

#define CK(call)                                                                         \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != 0) die(std::string(#call) + " -> " + std::to_string(rc_) + ": " + pigs_last_error()); \
    } while (0)

void say(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void say(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
    putchar('\n');
    fflush(stdout);
}
// stdout of `program vpi` (vpi.f90:161-194, 552-586, 620-634): list-directed `print *` starts a record with one blank
// and writes a default integer in 12 columns (gfortran); formats 101 (x,a,x,f7.2,x,a), 102 (a,x,G16.8e2,x,a,x,G16.8e2),
// 103 (x,a,x,i5), 104 (x,a,x,G13.6e2), 105 (x,a,x,3G13.6e2).  Same lines as driver.py (tests compare them).
double pct(double a, double t) { return t != 0 ? (double)(float)(100.0 * a) / t : std::nan(""); }                 // 100*real(acc)/try
double pct2(double a, double t) { return t != 0 ? 100.0 * (double)(float)a / (double)(float)t : std::nan(""); }   // 100.d0*real(acc)/real(try)
void ld(const std::string& text) { say("%s", (" " + text).c_str()); }
void ldi(const std::string& text, long long v) { say(" %s%12lld", text.c_str(), v); }
void f101(const std::string& text, double v, const char* tail) { say(" %s %s %s", text.c_str(), fortran_f(v, 7, 2).c_str(), tail); }
void f102(const std::string& text, double a, double b) { say("%s %s +/- %s", text.c_str(), fortran_g(a, 16, 8, 2).c_str(), fortran_g(b, 16, 8, 2).c_str()); }
void f103(const std::string& text, int v) { say(" %s %5d", text.c_str(), v); }

}  // namespace

int main(int argc, char** argv) {
    std::string workdir = ".", potential = "hfdb", rng_arg;
    int chains_arg = 0, device = 0, gpus_arg = 0;
    bool tables_only = false, format_test = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) die(a + " needs a value"); return argv[++i]; };
        if (a == "--workdir") workdir = next();
        else if (a == "--chains") chains_arg = std::atoi(next().c_str());
        else if (a == "--rng") rng_arg = next();
        else if (a == "--device") device = std::atoi(next().c_str());
        else if (a == "--gpus") gpus_arg = std::atoi(next().c_str());
        else if (a == "--potential") potential = next();
        else if (a == "--tables-only") tables_only = true;
        else if (a == "--format-test") format_test = true;
        else if (a == "-h" || a == "--help") {
            puts("usage: vpi_cuda [--workdir DIR] [--chains N] [--gpus G] [--rng philox|mt] [--device D] [--potential hfdb|hfdhe2|zero]\n"
                 "                [--tables-only] < vpi.in");
            return 0;
        } else die("unknown argument " + a);
    }
    if (format_test) {
        int w, d, e;
        std::string xs;
        while (std::cin >> w >> d >> e >> xs) fputs((fortran_g(std::strtod(xs.c_str(), nullptr), w, d, e) + "\n").c_str(), stdout);
        return 0;
    }
    if (potential != "hfdb" && potential != "hfdhe2" && potential != "zero") die("unknown potential " + potential);
    std::stringstream ss;
    ss << std::cin.rdbuf();
    Config c = read_vpi_in(ss.str());
    auto P = [&](const char* name) { return workdir + "/" + name; };
    if (system(("mkdir -p '" + workdir + "'").c_str()) != 0) die("cannot create " + workdir);
    std::vector<std::vector<double>> R0;          // crystal sites
    if (c.crystal) {                               // config_ini.in: Np / Lbox / density / positions (vpi.f90:101-107, vpi_mod.f90:220-228)
        std::ifstream in(P("config_ini.in"));
        if (!in) die("crystal=T but config_ini.in is missing");
        std::string t;
        in >> t; c.Np = std::atoi(t.c_str());
        for (int k = 0; k < c.dim; ++k) { in >> t; c.Lbox_in.push_back(to_double(t)); }
        in >> t; c.density = to_double(t);
        R0.assign(c.Np, std::vector<double>(c.dim));
        for (int ip = 0; ip < c.Np; ++ip)
            for (int k = 0; k < c.dim; ++k) { if (!(in >> t)) die("config_ini.in is too short"); R0[ip][k] = to_double(t); }
    }
    const Geometry g = derive_geometry(c);
    const int dim = c.dim, Np = c.Np, Nb = c.Nb, S = 2 * Nb + 1, Nmax = c.Nmax, Nbin = c.Nbin, Nk = c.Nk, Npw = c.Npw;

    // ---- tables + jastrow.out / potential.out (vpi_mod.f90:84-145)
    std::vector<double> W(Nmax + 2, 0.0), V(Nmax + 2, 0.0);
    if (potential != "zero") {
        W = make_table([&](double r) { return mcmillan_logpsi(r, c.Rm); }, g.dr, Nmax);
        V = potential == "hfdb" ? make_table(aziz_hfdb, g.dr, Nmax) : make_table(aziz_hfdhe2, g.dr, Nmax);
    }
    {
        std::ofstream fj(P("jastrow.out")), fp(P("potential.out"));
        for (int i = 0; i < Nmax; ++i) {
            const double r = (double)i * g.dr;
            fj << g_line({r, std::isfinite(W[i + 1]) ? std::exp(W[i + 1]) : 0.0, W[i + 1]});
            fp << g_line({r, V[i + 1]});
        }
    }
    if (tables_only) {
        printf("dim %d Np %d Nb %d Nmax %d sampling %s n_chains %d rng %s\n", dim, Np, Nb, Nmax, c.sampling.c_str(), c.n_chains, c.rng.c_str());
        printf("Lbox %.17g %.17g %.17g\nrcut %.17g\ndr %.17g\nrbin %.17g\ndensity %.17g\ndelta_cm %.17g\n", g.Lbox[0], g.Lbox[1], g.Lbox[2],
               g.rcut, g.dr, g.rbin, g.density, g.delta_cm);
        printf("config resume %d seed %d CMFreq %d Lstag %d Nlev %d Nstag %d Nblock %d Nstep %d Nbin %d Nk %d swapping %d Nobdm %d Npw %d "
               "trap %d crystal %d wf_table %d v_table %d threads_per_chain %d table_mode %d\n",
               (int)c.resume, c.seed, c.CMFreq, c.Lstag, c.Nlev, c.Nstag, c.Nblock, c.Nstep, c.Nbin, c.Nk, (int)c.swapping, c.Nobdm, c.Npw,
               (int)c.trap, (int)c.crystal, (int)c.wf_table, (int)c.v_table, c.threads_per_chain, c.table_mode);
        printf("reals dt %.17g CWorm %.17g Rm %.17g a_ho", c.dt, c.CWorm, c.Rm);
        for (double a : c.a_ho) printf(" %.17g", a);
        printf("\naction %s\n", c.action.empty() ? "chin" : c.action.c_str());
        return 0;
    }

    // ---- context
    const int n_chains = chains_arg > 0 ? chains_arg : std::max(1, c.n_chains);
    std::string rng = !rng_arg.empty() ? rng_arg : (!c.rng.empty() ? c.rng : (n_chains == 1 ? "mt" : "philox"));
    pigs_params p;
    std::memset(&p, 0, sizeof p);
    p.dim = dim; p.Np = Np; p.Nb = Nb; p.Nmax = Nmax; p.Nbin = Nbin; p.Nk = Nk; p.Npw = Npw; p.trap = c.trap ? 1 : 0;
    for (int k = 0; k < 3; ++k) { p.Lbox[k] = g.Lbox[k]; p.a_ho[k] = g.a_ho[k]; }
    p.rcut = g.rcut; p.dr = g.dr; p.density = g.density; p.dt = c.dt; p.delta_cm = g.delta_cm; p.CWorm = c.CWorm;
    p.CMFreq = c.CMFreq;
    p.sampling = lower(c.sampling).rfind("sta", 0) == 0 ? 0 : 1;
    p.Lstag = c.Lstag; p.Nlev = c.Nlev; p.Nstag = c.Nstag; p.Nobdm = c.Nobdm; p.swapping = c.swapping ? 1 : 0;
    p.n_chains = n_chains;
    p.rng_mode = rng.rfind("mt", 0) == 0 ? PIGS_RNG_MT_REPLAY : PIGS_RNG_PHILOX;
    p.seed = (uint64_t)c.seed;
    p.device = device; p.threads_per_chain = c.threads_per_chain; p.table_mode = c.table_mode; p.action = c.action.rfind("prim", 0) == 0 ? 1 : 0;
    p.schedule = c.schedule; p.chain_offset = 0; p.gpus = gpus_arg > 0 ? gpus_arg : c.gpus;
    pigs_handle h = nullptr;
    CK(pigs_create(&p, &h));
    CK(pigs_set_tables(h, W.data(), V.data()));

    // ---- init (vpi_mod.f90:149-259)
    std::vector<double> Path((size_t)S * Np * dim);
    double xend[6] = {0, 0, 0, 0, 0, 0};
    // resume = T with the extended checkpoint of an earlier run present: every chain and every accumulator comes
    // back and the run continues at the next block (what the reference cannot do, Q10); else the reference's own
    // resume: chain 0 from checkpoint.dat + rand_state, accumulators lost
    const bool full_resume = c.resume && file_exists(P("checkpoint_chains.bin")) && file_exists(P("checkpoint_driver.bin"));
    if (full_resume) {
        CK(pigs_load_checkpoint(h, P("checkpoint_chains.bin").c_str()));
    } else if (c.resume) {
        bool tr; int isopen, iworm;
        read_checkpoint(P("checkpoint.dat"), dim, Np, Nb, tr, isopen, iworm, Path, xend);
        std::vector<uint32_t> mt(624);
        int32_t mti;
        read_rand_state(P("rand_state"), mt.data(), mti);
        for (int ch = 0; ch < n_chains; ++ch) CK(pigs_set_state(h, ch, Path.data(), xend, isopen, iworm));
        CK(pigs_set_mt(h, 0, mt.data(), mti));
    } else {
        std::vector<double> u((size_t)Np * dim);
        for (int ch = 0; ch < n_chains; ++ch) {
            CK(pigs_sgrnd(h, ch, c.seed + ch));
            std::vector<std::vector<double>> R(Np, std::vector<double>(dim));
            if (c.crystal) R = R0;
            else {
                CK(pigs_grnd(h, ch, Np * dim, u.data()));                   // ip-major, k-minor draws
                for (int ip = 0; ip < Np; ++ip)
                    for (int k = 0; k < dim; ++k)
                        R[ip][k] = c.trap ? 2.0 * g.a_ho[k] * (u[ip * dim + k] - 0.5)       // vpi_mod.f90:208-214
                                          : g.Lbox[k] * (u[ip * dim + k] - 0.5);            // vpi_mod.f90:232-236
            }
            for (int ib = 0; ib < S; ++ib)
                for (int ip = 0; ip < Np; ++ip)
                    for (int k = 0; k < dim; ++k) Path[((size_t)ib * Np + ip) * dim + k] = R[ip][k];
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < dim; ++k) xend[j * dim + k] = R[Np - 1][k];             // vpi_mod.f90:250-254
            CK(pigs_set_state(h, ch, Path.data(), xend, 0, 0));
        }
    }

    // ---- banner (vpi.f90:161-194)
    const int Nblock = c.Nblock, Nstep = c.Nstep;
    {
        const bool sta = lower(c.sampling).rfind("sta", 0) == 0;
        ld(""); ld("=============================================================="); ld("                      VPI Monte Carlo                         ");
        ld("=============================================================="); ld(""); ld(" ");
        ld(std::string("# The Monte Carlo sampling will be performed using ") + (sta ? "STAGING" : "BISECTION")); ld("  algorithm");
        ld(c.swapping ? "# The Monte Carlo sampling will use swap updates" : "# The Monte Carlo sampling will not use swap updates");
        ld(" "); ld("# Simulation parameters:"); ld("");
        f103("  > Dimensions          :", dim); f103("  > Number of particles :", Np);
        auto g3 = [&](const double* v) { std::string t; for (int k = 0; k < dim; ++k) t += fortran_g(v[k], 13, 6, 2); return t; };
        if (c.trap) say(" %s %s", "  > Trapping length     :", g3(g.a_ho).c_str());
        else {
            say(" %s %s", "  > Density             :", fortran_g(g.density, 13, 6, 2).c_str());
            say(" %s %s", "  > Size of the box     :", g3(g.Lbox).c_str());
        }
        f103("  > Number of beads     :", Nb);
        say(" %s %s", "  > Time step           :", fortran_g(c.dt, 13, 6, 2).c_str());
        f103("  > Number of blocks    :", Nblock); f103("  > MC steps per block  :", Nstep);
        if (n_chains > 1) f103("  > Markov chains (GPU) :", n_chains);      // not in the reference: the replicas run at once
        ld("");
    }

    std::vector<double> Av(6, 0.0), Av2(6, 0.0), AvGr(Nbin, 0.0), AvGr2(Nbin, 0.0), AvSk((size_t)Nk * dim, 0.0), AvSk2((size_t)Nk * dim, 0.0);
    std::vector<double> AvNr((size_t)Nbin * (Npw + 1), 0.0), AvNr2((size_t)Nbin * (Npw + 1), 0.0), nrho((size_t)Nbin * (Npw + 1), 0.0);
    std::vector<double> gr(Nbin), Sk((size_t)std::max(Nk, 1) * dim), nr((size_t)Nbin * (Npw + 1)), rr(Nbin), nid(Nbin);
    const double k_n = std::pow(std::acos(-1.0), 0.5 * dim) / std::tgamma(0.5 * dim + 1.0);
    for (int j = 0; j < Nbin; ++j) {
        rr[j] = ((double)(j + 1) - 0.5) * g.rbin;
        nid[j] = g.density * k_n * (std::pow(rr[j] + 0.5 * g.rbin, dim) - std::pow(rr[j] - 0.5 * g.rbin, dim));
    }
    long long idiag_aux = 0;
    int obdm_bl = 0, diag_bl = 0, iblock0 = 0;
    // the driver's own state, in the order of checkpoint_driver.bin ("PIGSDRV1", count, doubles)
    auto driver_state = [&](bool save) {
        std::vector<std::vector<double>*> arrs = {&Av, &Av2, &AvGr, &AvGr2, &AvSk, &AvSk2, &AvNr, &AvNr2, &nrho};
        size_t n = 4;
        for (auto* a : arrs) n += a->size();
        std::vector<double> v;
        if (save) {
            v = {(double)iblock0, (double)idiag_aux, (double)obdm_bl, (double)diag_bl};
            for (auto* a : arrs) v.insert(v.end(), a->begin(), a->end());
            FILE* f = fopen(P("checkpoint_driver.bin").c_str(), "wb");
            if (!f) die("cannot write checkpoint_driver.bin");
            const uint64_t cnt = v.size();
            fwrite("PIGSDRV1", 1, 8, f); fwrite(&cnt, 8, 1, f); fwrite(v.data(), 8, v.size(), f);
            fclose(f);
        } else {
            FILE* f = fopen(P("checkpoint_driver.bin").c_str(), "rb");
            char mg[8]; uint64_t cnt = 0;
            if (!f || fread(mg, 1, 8, f) != 8 || std::memcmp(mg, "PIGSDRV1", 8) != 0 || fread(&cnt, 8, 1, f) != 1 || cnt != n)
                die("checkpoint_driver.bin does not belong to this configuration");
            v.resize(n);
            if (fread(v.data(), 8, n, f) != n) die("checkpoint_driver.bin is truncated");
            fclose(f);
            iblock0 = (int)v[0]; idiag_aux = (long long)v[1]; obdm_bl = (int)v[2]; diag_bl = (int)v[3];
            size_t o = 4;
            for (auto* a : arrs) { std::copy(v.begin() + o, v.begin() + o + a->size(), a->begin()); o += a->size(); }
        }
    };
    if (full_resume) driver_state(false);
    std::ofstream fe(P("e_vpi.out"), full_resume ? std::ios::app : std::ios::out), fet(P("et_vpi.out"), full_resume ? std::ios::app : std::ios::out);
    for (int iblock = iblock0 + 1; iblock <= Nblock; ++iblock) {
        const auto t0 = std::chrono::steady_clock::now();
        CK(pigs_run_block(h, Nstep));
        pigs_block_result b;
        CK(pigs_get_block(h, &b, gr.data(), Sk.data(), nr.data()));
        const long long nd = b.idiag_block;
        for (size_t i = 0; i < nrho.size(); ++i) nrho[i] += nr[i];
        idiag_aux += nd;
        double m[6], bvar[6];
        if (nd != 0) {                                                      // vpi.f90:477-520
            const double s1[6] = {b.sumE, b.sumK, b.sumV, b.sumEt, b.sumKt, b.sumVt};
            const double s2[6] = {b.sumE2, b.sumK2, b.sumV2, b.sumEt2, b.sumKt2, b.sumVt2};
            for (int i = 0; i < 6; ++i) {
                m[i] = s1[i] / (double)nd;
                bvar[i] = var(nd, m[i], s2[i] / (double)nd);
                Av[i] += m[i];
                Av2[i] += m[i] * m[i];
            }
            diag_bl += 1;
            if (!c.trap) {
                const double ngr = (double)b.ngr;
                for (int j = 0; j < Nbin; ++j) {                           // NormalizeGr (sample_mod.f90:656-679)
                    const double x = gr[j] / (nid[j] * ((double)Np * ngr));
                    AvGr[j] += x; AvGr2[j] += x * x;
                }
                for (size_t j = 0; j < (size_t)Nk * dim; ++j) {            // NormalizeSk (:683-702)
                    const double x = Sk[j] / ((double)Np * ngr);
                    AvSk[j] += x; AvSk2[j] += x * x;
                }
            }
            fe << g_line({(double)(float)iblock, m[0] / Np, m[1] / Np, m[2] / Np});
            fet << g_line({(double)(float)iblock, m[3] / Np, m[4] / Np, m[5] / Np});
            fe.flush(); fet.flush();
        } else {
            for (int i = 0; i < 6; ++i) m[i] = bvar[i] = std::nan("");
        }
        // OBDM block (vpi.f90:522-539): closes once >= Nstep diagonal configurations (per chain) accumulated
        if (idiag_aux / ((long long)Nstep * n_chains) >= 1) {
            obdm_bl += 1;
            if (!c.trap && c.CWorm > 0 && c.Nobdm > 0) {
                for (int j = 0; j < Nbin; ++j) {                           // NormalizeNr (:706-732)
                    const double den = c.CWorm * nid[j] * (double)idiag_aux * (double)c.Nobdm;
                    for (int q = 0; q <= Npw; ++q) {
                        const double x = nrho[(size_t)j * (Npw + 1) + q] / den;
                        AvNr[(size_t)j * (Npw + 1) + q] += x; AvNr2[(size_t)j * (Npw + 1) + q] += x * x;
                    }
                }
            }
            idiag_aux = 0;
            std::fill(nrho.begin(), nrho.end(), 0.0);
        }
        // CheckPoint of chain 0 every block (vpi.f90:541-545) in the reference's single-chain format
        {
            int isopen, iworm;
            CK(pigs_get_state(h, 0, Path.data(), xend, &isopen, &iworm));
            write_checkpoint(P("checkpoint.dat"), c.trap, Path, xend, isopen, iworm, S, Np, dim);
            std::vector<uint32_t> mt(624);
            int32_t mti;
            CK(pigs_get_mt(h, 0, mt.data(), &mti));
            append_rand_state(P("rand_state"), mt.data(), mti);
        }
        // ... and the extended checkpoint: all chains (library) + the accumulators above (checkpoint_driver.bin)
        if (c.checkpoint_every > 0 && (iblock % c.checkpoint_every == 0 || iblock == Nblock)) {
            CK(pigs_save_checkpoint(h, P("checkpoint_chains.bin").c_str()));
            iblock0 = iblock;
            driver_state(true);
        }
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        // the block report (vpi.f90:552-586)
        ld("-----------------------------------------------------------"); ldi("BLOCK NUMBER :", iblock); ld(" "); ld("# Block results:"); ld(" ");
        const char* lab2[6] = {"  > <E>  =", "  > <Ec> =", "  > <Ep> =", "  > <Et> =", "  > <Kt> =", "  > <Vt> ="};
        for (int i = 0; i < 3; ++i) f102(lab2[i], m[i] / Np, bvar[i] / Np);
        ld(" ");
        for (int i = 3; i < 6; ++i) f102(lab2[i], m[i] / Np, bvar[i] / Np);
        ld(""); ld("# Acceptance of diagonal movements:"); ld(" ");
        f101("> CM movements      =", pct((double)b.acc_cm, (double)b.try_cm), "%");
        f101("> Staging movements =", pct((double)b.acc_bd, (double)b.try_stag), "%");
        f101("> Head movements    =", pct((double)b.acc_head, (double)b.try_stag), "%");
        f101("> Tail movements    =", pct((double)b.acc_tail, (double)b.try_stag), "%");
        ld(" "); ld("# Acceptance of off-diagonal movements:"); ld(" ");
        f101("> CM movements      =", pct((double)b.acc_cm_half, (double)b.try_cm_half), "%");
        f101("> Staging movements =", pct((double)b.acc_bd_half, (double)b.try_stag_half), "%");
        f101("> Head movements    =", pct((double)b.acc_head_half, (double)b.try_stag_half), "%");
        f101("> Tail movements    =", pct((double)b.acc_tail_half, (double)b.try_stag_half), "%");
        ld(" "); ld("# Acceptance open/close updates:"); ld(" ");
        f101("> Diagonal conf.    =", pct2((double)nd, (double)Nstep * n_chains), "%");
        f101("> Open acc          =", pct2((double)b.acc_open, (double)b.try_open), "%");
        f101("> Close acc         =", pct2((double)b.acc_close, (double)b.try_close), "%");
        f101("> Swap acc          =", pct2((double)b.acc_swap, (double)b.try_swap), "%");
        say("  "); f101("# Time per block    =", dt, "seconds");
        say(" # GPU throughput    = %.4g bead-updates/s", (double)(b.bead_updates[0] + b.bead_updates[1] + b.bead_updates[2]) / dt);
    }
    fe.close();
    fet.close();
    // fort.99: permutation histogram (vpi.f90:590-592), summed over chains
    {
        std::vector<long long> hist(Np, 0);
        std::vector<int32_t> cyc(Np), hh(Np);
        for (int ch = 0; ch < n_chains; ++ch) {
            int ipm, a, bb;
            CK(pigs_get_perm(h, ch, &ipm, cyc.data(), hh.data(), &a, &bb));
            for (int ip = 0; ip < Np; ++ip) hist[ip] += hh[ip];
        }
        FILE* f = fopen(P("fort.99").c_str(), "w");
        for (int ip = 0; ip < Np; ++ip) fprintf(f, " %11d %11lld\n", ip + 1, hist[ip]);
        fclose(f);
    }
    // finals (vpi.f90:606-642)
    if (diag_bl) {
        ld("=============================================================="); ld("FINAL RESULTS:"); ld(""); ld("# Final averages:"); ld("");
        const char* lab2[6] = {"  > <E>  =", "  > <Ec> =", "  > <Ep> =", "  > <Et> =", "  > <Kt> =", "  > <Vt> ="};
        for (int i = 0; i < 6; ++i) {
            if (i == 3) ld("");
            const double A = Av[i] / diag_bl, A2 = Av2[i] / diag_bl;
            f102(lab2[i], A / Np, var(diag_bl, A, A2) / Np);
        }
        ld(""); ld("=============================================================="); ld("");
    }
    if (!c.trap && diag_bl) {
        {
            std::ofstream f(P("gr_vpi.out"));                               // NormAvGr (sample_mod.f90:794-816)
            for (int j = 0; j < Nbin; ++j) {
                const double a = AvGr[j] / diag_bl, a2 = AvGr2[j] / diag_bl;
                f << g_line({rr[j], a, var(diag_bl, a, a2)});
            }
        }
        {
            std::ofstream f(P("sk_vpi.out"));                               // NormAvSk (:820-842)
            for (int j = 0; j < Nk; ++j) {
                std::vector<double> row;
                for (int k = 0; k < dim; ++k) {
                    const double a = AvSk[(size_t)j * dim + k] / diag_bl, a2 = AvSk2[(size_t)j * dim + k] / diag_bl;
                    row.push_back((double)(j + 1) * (2.0 * g.pi / g.Lbox[k]));
                    row.push_back(a);
                    row.push_back(var(diag_bl, a, a2));
                }
                f << g_line(row);
            }
        }
        if (obdm_bl) {
            std::ofstream f(P("nr_vpi.out"));                               // NormAvNr (:846-870)
            for (int j = 0; j < Nbin; ++j) {
                std::vector<double> row{rr[j]};
                for (int q = 0; q <= Npw; ++q) {
                    const double a = AvNr[(size_t)j * (Npw + 1) + q] / obdm_bl, a2 = AvNr2[(size_t)j * (Npw + 1) + q] / obdm_bl;
                    row.push_back(a);
                    row.push_back(var(obdm_bl, a, a2));
                }
                f << g_line(row);
            }
        }
    }
    CK(pigs_destroy(h));
    return 0;
}
