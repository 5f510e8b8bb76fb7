// f90rt.h -- runtime support for the C++ that oracle/f90toc/f90toc.py generates from the reference's Fortran.
// TEST INFRASTRUCTURE ONLY.  Everything here restates what gfortran's middle end / libgfortran do for the few
// constructs the translation cannot express directly in C++.
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace f90rt {

// x**n, n a compile-time constant: gfortran (trans-expr.c gfc_conv_powi) and GCC (__builtin_powi expansion) both
// build x^n = x^(n - T[n]) * x^T[n] from this table of optimal addition chains -- x**5 is (x*x*x)*(x*x), not
// ((((x*x)*x)*x)*x).  (GCC 13 -O2 on `__builtin_powi(x,5)` emits exactly mulsd x,x; mulsd x2,x; mulsd x2,x3.)
constexpr unsigned char POWI_TABLE[33] = {
    0,  1,  1,  2,  2,  3,  3,  4,  4,  6,  5,  6,  6,  10, 7,  9,  8,  16, 9,  16, 10, 12, 11, 13, 12, 17, 13, 18, 14, 24, 15, 26,
    16};
template <int N>
inline double powi_c(double x) {
    static_assert(N >= 0 && N <= 32, "constant exponent out of the expansion table");
    if constexpr (N == 0) return 1.0;
    else if constexpr (N == 1) return x;
    else {
        const double a = powi_c<N - POWI_TABLE[N]>(x);      // each distinct power is computed once in GCC's expansion; the
        const double b = powi_c<POWI_TABLE[N]>(x);          // value of a power does not depend on how often it is formed
        return a * b;
    }
}
// x**n, n variable: libgfortran pow_r8_i4 (square-and-multiply from the low bit)
inline double powi_v(double x, int n) {
    double pow = 1.0;
    if (n != 0) {
        unsigned u;
        if (n < 0) { u = (unsigned)(-n); x = 1.0 / x; } else u = (unsigned)n;
        for (;;) {
            if (u & 1) pow *= x;
            u >>= 1;
            if (u) x *= x; else break;
        }
    }
    return pow;
}
inline int ipow_int(int b, int n) {           // integer**integer (libgfortran pow_i4_i4)
    if (n < 0) return (b == 1) ? 1 : ((b == -1) ? ((n & 1) ? -1 : 1) : 0);
    int r = 1;
    unsigned u = (unsigned)n;
    for (;;) {
        if (u & 1) r *= b;
        u >>= 1;
        if (u) b *= b; else break;
    }
    return r;
}
inline int ishft(int v, int s) {              // logical shift, zero fill
    if (s >= 32 || s <= -32) return 0;
    return s >= 0 ? (int)((uint32_t)v << s) : (int)((uint32_t)v >> (-s));
}
inline double sign(double a, double b) { return std::copysign(std::fabs(a), b); }
inline double minval(const double* p, int n) {
    double m = p[0];
    for (int i = 1; i < n; ++i) if (p[i] < m) m = p[i];
    return m;
}
inline bool streq(const std::string& a, const std::string& b) {      // Fortran pads the shorter operand with blanks
    size_t n = a.size() > b.size() ? a.size() : b.size();
    for (size_t i = 0; i < n; ++i) {
        char x = i < a.size() ? a[i] : ' ', y = i < b.size() ? b[i] : ' ';
        if (x != y) return false;
    }
    return true;
}
inline void* alloc(size_t bytes) { return std::calloc(bytes ? bytes : 1, 1); }
inline void dealloc(void* p) { std::free(p); }
[[noreturn]] inline void stop() { throw std::runtime_error("Fortran STOP"); }

// write(unit,...) list -> captured records; read(unit,...) list <- records queued by the glue
struct Rec {
    std::vector<double> v;
    std::vector<std::string> s;
    int n = 0;
};
struct Unit {
    std::string file;                                     // name given by the last open(); "" = never opened (fort.N)
    std::vector<std::vector<double>> to_read;
    size_t rpos = 0;
};
struct Files {                                            // every record written, by file name
    std::vector<std::pair<std::string, std::vector<Rec>>> f;
    std::vector<Rec>& get(const std::string& name) {
        for (auto& x : f) if (x.first == name) return x.second;
        f.emplace_back(name, std::vector<Rec>());
        return f.back().second;
    }
    void clear() { f.clear(); }
};
inline Files& files() {
    static Files F;
    return F;
}
inline Unit& unit(int u) {
    static Unit units[128];
    return units[(u >= 0 && u < 128) ? u : 127];
}
inline std::vector<double>& read_cur() {
    static std::vector<double> cur;
    return cur;
}
inline size_t& read_idx() {
    static size_t i = 0;
    return i;
}
inline void open_unit(int u, const char* name) {
    unit(u).file = name;
    unit(u).rpos = 0;
    files().get(name).clear();                            // open without position= rewinds/replaces
}
inline void write_rec(int u, const Rec& r) {
    const std::string name = unit(u).file.empty() ? "fort." + std::to_string(u) : unit(u).file;
    files().get(name).push_back(r);
}
inline void read_rec(int u, const Rec& r) {
    Unit& U = unit(u);
    if (U.rpos >= U.to_read.size()) throw std::runtime_error("read past the records queued for unit " + std::to_string(u));
    read_cur() = U.to_read[U.rpos++];
    read_idx() = 0;
    if ((int)read_cur().size() < r.n) throw std::runtime_error("record too short for the read list");
}
inline double read_next() { return read_cur()[read_idx()++]; }

}  // namespace f90rt
