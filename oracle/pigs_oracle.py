"""ctypes binding of the CPU ORACLE (oracle/libpigs_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by the product
package.  See oracle/pigs_oracle.h for what the oracle is and how it is pinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

MOVES = dict(
    translate_chain=0, staging=1, move_head=2, move_tail=3, bisection=4,
    move_head_bisection=5, move_tail_bisection=6, translate_half=7, staging_half=8,
    move_head_half=9, move_tail_half=10, open=11, close=12, swap=13,
)


class OrcParams(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("Np", C.c_int32), ("density", C.c_double),
        ("crystal", C.c_int32), ("trap", C.c_int32), ("dt", C.c_double),
        ("Nb", C.c_int32), ("seed", C.c_int32), ("delta_cm", C.c_double),
        ("CMFreq", C.c_int32), ("sampling", C.c_int32),
        ("Lstag", C.c_int32), ("Nlev", C.c_int32), ("Nstag", C.c_int32), ("Nbin", C.c_int32), ("Nk", C.c_int32),
        ("swapping", C.c_int32), ("CWorm", C.c_double), ("Nobdm", C.c_int32), ("Npw", C.c_int32),
        ("Nmax", C.c_int32), ("wf_table", C.c_int32), ("v_table", C.c_int32), ("Rm", C.c_double),
        ("a_ho", C.c_double * 3), ("Lbox_crystal", C.c_double * 3), ("action", C.c_int32), ("pad_", C.c_int32),
    ]


class OrcBlock(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("sumE", "sumK", "sumV", "sumEt", "sumKt", "sumVt",
                 "sumE2", "sumK2", "sumV2", "sumEt2", "sumKt2", "sumVt2")] + \
               [("idiag_block", C.c_int32), ("ngr", C.c_int32)] + \
               [(n, C.c_double) for n in ("try_cm", "try_stag", "try_cm_half", "try_stag_half")] + \
               [(n, C.c_int32) for n in
                ("acc_cm", "acc_bd", "acc_head", "acc_tail", "acc_cm_half", "acc_bd_half", "acc_head_half",
                 "acc_tail_half", "try_open", "acc_open", "try_close", "acc_close", "try_swap", "acc_swap",
                 "idiag_aux", "pad_")] + \
               [("bead_updates", C.c_uint64 * 3)]

    def as_dict(self):
        d = {}
        for n, _ in self._fields_:
            v = getattr(self, n)
            d[n] = list(v) if n == "bead_updates" else v
        d.pop("pad_")
        return d


def build(native: bool = False) -> str:
    """make the oracle library if needed; returns its path"""
    name = "libpigs_oracle_native.so" if native else "libpigs_oracle.so"
    path = os.path.join(_HERE, name)
    src = os.path.join(_HERE, "pigs_oracle.cpp")
    hdr = os.path.join(_HERE, "pigs_oracle.h")
    if (not os.path.exists(path)) or any(os.path.getmtime(f) > os.path.getmtime(path) for f in (src, hdr)):
        subprocess.check_call(["make", "-C", _HERE, name], stdout=subprocess.DEVNULL)
    return path


_LIBS = {}


def _lib(native: bool = False):
    if native in _LIBS:
        return _LIBS[native]
    L = C.CDLL(build(native))
    dp, ip_, u32p, u64p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    vp = C.c_void_p
    sig = {
        "orc_create": (vp, [C.POINTER(OrcParams)]),
        "orc_destroy": (None, [vp]),
        "orc_get_geometry": (None, [vp, dp, dp, dp, dp, dp, dp]),
        "orc_fill_tables": (None, [vp]),
        "orc_set_tables": (None, [vp, dp, dp]),
        "orc_get_tables": (None, [vp, dp, dp]),
        "orc_init": (None, [vp, dp]),
        "orc_set_state": (None, [vp, dp, dp, C.c_int, C.c_int]),
        "orc_get_state": (None, [vp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "orc_set_perm": (None, [vp, C.c_int, ip_, ip_, C.c_int, C.c_int]),
        "orc_get_perm": (None, [vp, C.POINTER(C.c_int), ip_, ip_, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
        "orc_sgrnd": (None, [vp, C.c_int32]),
        "orc_grnd": (C.c_double, [vp]),
        "orc_rangauss": (C.c_double, [vp]),
        "orc_mt_raw": (C.c_uint32, [vp]),
        "orc_get_mt": (None, [vp, u32p, ip_]),
        "orc_set_mt": (None, [vp, u32p, C.c_int32]),
        "orc_interpolate": (C.c_double, [C.c_int, C.c_int, C.c_double, dp, C.c_double]),
        "orc_potential": (C.c_double, [C.c_double]),
        "orc_logpsi": (C.c_double, [C.c_int, C.c_double, C.c_double]),
        "orc_green_function": (C.c_double, [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
        "orc_minimum_image": (None, [vp, dp, dp]),
        "orc_boundary_conditions": (C.c_double, [vp, C.c_int, C.c_double]),
        "orc_update_action": (C.c_double, [vp, C.c_int, C.c_int, dp, dp]),
        "orc_update_action_R": (C.c_double, [vp, dp, C.c_int, C.c_int, dp, dp]),
        "orc_move": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
        "orc_potential_energy": (None, [vp, dp, C.c_int, dp, dp]),
        "orc_local_energy": (None, [vp, dp, dp, dp, dp]),
        "orc_therm_energy": (None, [vp, dp, dp, dp]),
        "orc_therm_energy_P": (None, [vp, dp, dp, dp, dp]),
        "orc_pair_correlation": (None, [vp, dp, dp]),
        "orc_structure_factor": (None, [vp, dp, dp]),
        "orc_obdm": (None, [vp, dp, dp]),
        "orc_normalize_gr": (None, [vp, C.c_int, dp]),
        "orc_normalize_sk": (None, [vp, C.c_int, dp]),
        "orc_normalize_nr": (None, [vp, C.c_double, dp]),
        "orc_var": (C.c_double, [C.c_int, C.c_double, C.c_double]),
        "orc_run_block": (None, [vp, C.c_int, C.POINTER(OrcBlock), dp, dp, dp]),
        "orc_bead_updates": (None, [vp, u64p]),
    }
    for k, (res, args) in sig.items():
        f = getattr(L, k)
        f.restype, f.argtypes = res, args
    _LIBS[native] = L
    return L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


# stateless primitives ---------------------------------------------------
def interpolate(opt, F, dx, x):
    F = _f64(F)
    return _lib().orc_interpolate(opt, len(F) - 2, dx, _dp(F), x)


def potential(r):
    return _lib().orc_potential(r)


def logpsi(opt, Rm, r):
    return _lib().orc_logpsi(opt, Rm, r)


def var(n, s, s2):
    return _lib().orc_var(n, s, s2)


class Oracle:
    """One reference simulation (the Fortran program's global state made an object).

    `cfg` is a dict with the vpi.in variable names (see OrcParams)."""

    DEFAULTS = dict(dim=3, Np=64, density=0.365, crystal=0, trap=0, dt=5e-3, Nb=15, seed=1982, delta_cm=0.12,
                    CMFreq=1, sampling="bis", Lstag=2, Nlev=1, Nstag=5, Nbin=100, Nk=50, swapping=0, CWorm=0.0,
                    Nobdm=0, Npw=0, Nmax=10000, wf_table=1, v_table=1, Rm=1.2, a_ho=(1.0, 1.0, 1.0),
                    Lbox_crystal=(0.0, 0.0, 0.0), action=0)

    def __init__(self, cfg: dict, native: bool = False):
        self.L = _lib(native)
        c = dict(self.DEFAULTS)
        c.update({k: v for k, v in cfg.items() if k in self.DEFAULTS})
        self.cfg = c
        p = OrcParams()
        for k, v in c.items():
            if k == "sampling":
                p.sampling = 0 if str(v).strip().lower().startswith("sta") else 1
            elif k in ("a_ho", "Lbox_crystal"):
                vv = list(v) + [1.0] * 3
                setattr(p, k, (C.c_double * 3)(*vv[:3]))
            elif k in ("density", "dt", "delta_cm", "CWorm", "Rm"):
                setattr(p, k, float(v))
            else:
                setattr(p, k, int(v))
        self.p = p
        self.h = C.c_void_p(self.L.orc_create(C.byref(p)))
        self.dim, self.Np, self.Nb = p.dim, p.Np, p.Nb
        self.Nmax, self.Nbin, self.Nk, self.Npw = p.Nmax, p.Nbin, p.Nk, p.Npw
        Lb = np.zeros(3)
        sc = [C.c_double() for _ in range(5)]
        self.L.orc_get_geometry(self.h, _dp(Lb), *[C.byref(x) for x in sc])
        self.Lbox = Lb
        self.rcut, self.dr, self.rbin, self.density, self.delta_cm = [x.value for x in sc]

    def __del__(self):
        try:
            if self.h:
                self.L.orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # tables / state
    def fill_tables(self):
        self.L.orc_fill_tables(self.h)

    def set_tables(self, LogWF, VTable):
        a, b = _f64(LogWF), _f64(VTable)
        assert a.size == self.Nmax + 2 and b.size == self.Nmax + 2
        self.L.orc_set_tables(self.h, _dp(a), _dp(b))

    def get_tables(self):
        a, b = np.zeros(self.Nmax + 2), np.zeros(self.Nmax + 2)
        self.L.orc_get_tables(self.h, _dp(a), _dp(b))
        return a, b

    def init(self, R0=None):
        if R0 is None:
            self.L.orc_init(self.h, None)
        else:
            r = _f64(R0)
            self.L.orc_init(self.h, _dp(r))

    def path_shape(self):
        # C-order view of the Fortran Path(dim,Np,0:2Nb): [ib][ip][k]
        return (2 * self.Nb + 1, self.Np, self.dim)

    def get_state(self):
        P = np.zeros(self.path_shape())
        xe = np.zeros((2, self.dim))
        io, iw = C.c_int(), C.c_int()
        self.L.orc_get_state(self.h, _dp(P), _dp(xe), C.byref(io), C.byref(iw))
        return P, xe, io.value, iw.value

    def set_state(self, Path, xend, isopen=0, iworm=0):
        P, xe = _f64(Path), _f64(xend)
        assert P.size == np.prod(self.path_shape()) and xe.size == 2 * self.dim
        self.L.orc_set_state(self.h, _dp(P), _dp(xe), int(isopen), int(iworm))

    def get_perm(self):
        cyc = np.zeros(self.Np, dtype=np.int32)
        hist = np.zeros(self.Np, dtype=np.int32)
        ipm, a, b = C.c_int(), C.c_int(), C.c_int()
        ipt = C.POINTER(C.c_int32)
        self.L.orc_get_perm(self.h, C.byref(ipm), cyc.ctypes.data_as(ipt), hist.ctypes.data_as(ipt), C.byref(a), C.byref(b))
        return ipm.value, cyc, hist, a.value, b.value

    def set_perm(self, iperm, cyc, hist, new_pc=0, end_pc=0):
        cyc = np.ascontiguousarray(cyc, dtype=np.int32)
        hist = np.ascontiguousarray(hist, dtype=np.int32)
        ipt = C.POINTER(C.c_int32)
        self.L.orc_set_perm(self.h, int(iperm), cyc.ctypes.data_as(ipt), hist.ctypes.data_as(ipt), int(new_pc), int(end_pc))

    # rng
    def sgrnd(self, seed):
        self.L.orc_sgrnd(self.h, int(seed))

    def grnd(self):
        return self.L.orc_grnd(self.h)

    def rangauss(self):
        return self.L.orc_rangauss(self.h)

    def mt_raw(self):
        return self.L.orc_mt_raw(self.h)

    def get_mt(self):
        mt = np.zeros(624, dtype=np.uint32)
        mti = C.c_int32()
        self.L.orc_get_mt(self.h, mt.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(mti))
        return mt, mti.value

    def set_mt(self, mt, mti):
        mt = np.ascontiguousarray(mt, dtype=np.uint32)
        self.L.orc_set_mt(self.h, mt.ctypes.data_as(C.POINTER(C.c_uint32)), int(mti))

    # action / moves
    def green_function(self, opt, ib, dt, Pot, F2):
        return self.L.orc_green_function(self.h, opt, ib, dt, Pot, F2)

    def minimum_image(self, xij):
        x = _f64(xij).copy()
        r2 = C.c_double()
        self.L.orc_minimum_image(self.h, _dp(x), C.byref(r2))
        return x, r2.value

    def boundary_conditions(self, k, x):
        return self.L.orc_boundary_conditions(self.h, k, x)

    def update_action(self, ip, ib, xnew, xold, R=None):
        a, b = _f64(xnew), _f64(xold)
        if R is None:
            return self.L.orc_update_action(self.h, ip, ib, _dp(a), _dp(b))
        r = _f64(R)
        return self.L.orc_update_action_R(self.h, _dp(r), ip, ib, _dp(a), _dp(b))

    def move(self, name, ip, half=0):
        aux = C.c_int(0)
        acc = self.L.orc_move(self.h, MOVES[name], int(ip), int(half), C.byref(aux))
        return acc, aux.value

    # estimators
    def potential_energy(self, R, want_f2):
        r = _f64(R)
        a, b = C.c_double(), C.c_double()
        self.L.orc_potential_energy(self.h, _dp(r), int(want_f2), C.byref(a), C.byref(b))
        return a.value, b.value

    def local_energy(self, R):
        r = _f64(R)
        e, k, p = C.c_double(), C.c_double(), C.c_double()
        self.L.orc_local_energy(self.h, _dp(r), C.byref(e), C.byref(k), C.byref(p))
        return e.value, k.value, p.value

    def therm_energy(self, Path=None):
        e, k, p = C.c_double(), C.c_double(), C.c_double()
        if Path is None:
            self.L.orc_therm_energy(self.h, C.byref(e), C.byref(k), C.byref(p))
        else:
            P = _f64(Path)
            self.L.orc_therm_energy_P(self.h, _dp(P), C.byref(e), C.byref(k), C.byref(p))
        return e.value, k.value, p.value

    def pair_correlation(self, R, gr=None):
        r = _f64(R)
        g = np.zeros(self.Nbin) if gr is None else gr
        self.L.orc_pair_correlation(self.h, _dp(r), _dp(g))
        return g

    def structure_factor(self, R, Sk=None):
        r = _f64(R)
        s = np.zeros((self.Nk, self.dim)) if Sk is None else Sk
        self.L.orc_structure_factor(self.h, _dp(r), _dp(s))
        return s

    def obdm(self, xend, nrho=None):
        x = _f64(xend)
        n = np.zeros((self.Nbin, self.Npw + 1)) if nrho is None else nrho
        self.L.orc_obdm(self.h, _dp(x), _dp(n))
        return n

    def normalize_gr(self, ngr, gr):
        self.L.orc_normalize_gr(self.h, int(ngr), _dp(gr))
        return gr

    def normalize_sk(self, ngr, Sk):
        self.L.orc_normalize_sk(self.h, int(ngr), _dp(Sk))
        return Sk

    def normalize_nr(self, zconf, nrho):
        self.L.orc_normalize_nr(self.h, float(zconf), _dp(nrho))
        return nrho

    # driver
    def run_block(self, Nstep, nrho=None):
        b = OrcBlock()
        gr = np.zeros(self.Nbin)
        Sk = np.zeros((self.Nk, self.dim))
        nr = np.zeros((self.Nbin, self.Npw + 1)) if nrho is None else nrho
        self.L.orc_run_block(self.h, int(Nstep), C.byref(b), _dp(gr), _dp(Sk), _dp(nr))
        return b.as_dict(), gr, Sk, nr

    def bead_updates(self):
        a = (C.c_uint64 * 3)()
        self.L.orc_bead_updates(self.h, a)
        return list(a)
