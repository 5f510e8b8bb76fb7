"""ctypes binding of oracle/_ref/libpigs_ref.so -- the reference's own Fortran sources, machine-translated
to C++ by oracle/f90toc/f90toc.py (see oracle/Makefile, target `ref`).  TEST INFRASTRUCTURE ONLY: it exists to
pin the hand-written oracle to the reference text.  The translated code keeps the reference's module globals:
ONE configuration is live per process (`Ref(cfg)` re-runs the program prologue and replaces it)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libpigs_ref.so")
REFERENCE = os.environ.get("PIGS_REFERENCE", "/root/reference")


class RefParams(C.Structure):
    _fields_ = [("dim", C.c_int32), ("Np", C.c_int32), ("density", C.c_double), ("crystal", C.c_int32), ("trap", C.c_int32),
                ("dt", C.c_double), ("Nb", C.c_int32), ("seed", C.c_int32), ("delta_cm", C.c_double), ("CMFreq", C.c_int32),
                ("sampling", C.c_int32), ("Lstag", C.c_int32), ("Nlev", C.c_int32), ("Nstag", C.c_int32), ("Nblock", C.c_int32),
                ("Nstep", C.c_int32), ("Nbin", C.c_int32), ("Nk", C.c_int32), ("swapping", C.c_int32), ("CWorm", C.c_double),
                ("Nobdm", C.c_int32), ("Npw", C.c_int32), ("Nmax", C.c_int32), ("wf_table", C.c_int32), ("v_table", C.c_int32),
                ("Rm", C.c_double), ("a_ho", C.c_double * 3)]


def available() -> bool:
    return os.path.exists(LIB) or os.path.isdir(REFERENCE)


def build() -> str:
    """translate + compile when the reference sources are present; otherwise use the prebuilt library"""
    if os.path.isdir(REFERENCE):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"REF={REFERENCE}"])
    if not os.path.exists(LIB):
        raise FileNotFoundError(f"{LIB} is missing and {REFERENCE} does not exist: cannot build the translated reference")
    return LIB


_L = None


def lib():
    global _L
    if _L is None:
        L = C.CDLL(build())
        D, I, P = C.c_double, C.c_int, C.POINTER(C.c_double)
        L.ref_run_program.argtypes = [C.POINTER(RefParams)]
        L.ref_file.argtypes = [C.c_char_p, P, I, C.POINTER(I)]
        L.ref_geometry.argtypes = [P, P, P, P]
        L.ref_tables.argtypes = [P, P]
        for name, args in dict(ref_interpolate=[I, I, D, P, D], ref_potential=[D], ref_logpsi=[I, D, D], ref_trappsi=[I, D, D],
                               ref_trappot=[I, D, D], ref_green=[I, I, D, D, D], ref_r8_gamma=[D], ref_boundary=[I, D],
                               ref_grnd=[], ref_rangauss=[], ref_var=[I, D, D],
                               ref_update_action=[P, P, P, I, I, P, P, D, I]).items():
            f = getattr(L, name)
            f.argtypes, f.restype = args, D
        L.ref_minimum_image.argtypes = [P, P]
        L.ref_sgrnd.argtypes = [I]
        L.ref_get_mt.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]
        L.ref_set_mt.argtypes = [C.POINTER(C.c_uint32), C.c_int32]
        L.ref_local_energy.argtypes = [P, P, P, I, P, P, P]
        L.ref_therm_energy.argtypes = [P, P, D, I, P, P, P]
        L.ref_pair_correlation.argtypes = [P, P]
        L.ref_structure_factor.argtypes = [I, P, P]
        L.ref_obdm.argtypes = [P, P]
        L.ref_normalize_gr.argtypes = [D, I, P]
        L.ref_normalize_sk.argtypes = [I, I, P]
        L.ref_normalize_nr.argtypes = [D, D, I, P]
        L.ref_move.argtypes = [I, I, P, P, D, D, D, I, I, I, I, P, P, C.POINTER(I), C.POINTER(I)]
        L.ref_move.restype = I
        L.ref_bead_updates.argtypes, L.ref_bead_updates.restype = [], C.c_longlong
        L.ref_queue_read.argtypes = [I, P, I]
        L.ref_clear_reads.argtypes = [I]
        _L = L
    return _L


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Ref:
    """The translated reference configured like `./vpi < vpi.in` with the given namelist values."""

    DEFAULTS = dict(dim=3, Np=64, density=0.365, crystal=0, trap=0, dt=5e-3, Nb=15, seed=1982, delta_cm=0.12, CMFreq=1,
                    sampling="bis", Lstag=2, Nlev=1, Nstag=5, Nblock=0, Nstep=1, Nbin=100, Nk=50, swapping=0, CWorm=0.0,
                    Nobdm=0, Npw=0, Nmax=10000, wf_table=1, v_table=1, Rm=1.2, a_ho=(1.0, 1.0, 1.0))

    def __init__(self, cfg: dict, Nblock=0, Nstep=1, lattice=None):
        """lattice = (R[Np][dim], Lbox[dim]): run with crystal = .true.; the program then reads Np, Lbox, density and the
        starting positions from config_ini.in (vpi.f90:101-107, vpi_mod.f90:218-228) -- served from memory here"""
        self.L = lib()
        if lattice is not None:
            R, Lb = np.asarray(lattice[0], float), np.asarray(lattice[1], float)
            cfg = dict(cfg, crystal=1)
            self.L.ref_clear_reads(2)
            for rec in [[float(len(R))], list(Lb), [float(cfg["density"])]] + [list(x) for x in R]:
                a = np.asarray(rec, float)
                self.L.ref_queue_read(2, _dp(a), len(a))
        c = dict(self.DEFAULTS)
        c.update({k: v for k, v in cfg.items() if k in self.DEFAULTS})
        c["Nblock"], c["Nstep"] = Nblock, Nstep
        p = RefParams()
        for k, v in c.items():
            if k == "sampling":
                p.sampling = 0 if str(v).strip().lower().startswith("sta") else 1
            elif k == "a_ho":
                vv = list(v) + [1.0] * 3
                p.a_ho = (C.c_double * 3)(*vv[:3])
            elif k in ("density", "dt", "delta_cm", "CWorm", "Rm"):
                setattr(p, k, float(v))
            else:
                setattr(p, k, int(v))
        self.p, self.cfg = p, c
        if self.L.ref_run_program(C.byref(p)) != 0:
            raise RuntimeError("the translated program stopped")
        Lb = np.zeros(3)
        sc = [C.c_double() for _ in range(3)]
        self.L.ref_geometry(_dp(Lb), *[C.byref(x) for x in sc])
        self.Lbox = Lb
        self.rcut, self.dr, self.rbin = [x.value for x in sc]
        self.dim, self.Np, self.Nb, self.Nmax = p.dim, p.Np, p.Nb, p.Nmax

    def bead_updates(self):
        """UpdateAction calls since the library was loaded (the metric's unit)"""
        return int(self.L.ref_bead_updates())

    def file(self, name):
        """numeric records the program wrote to `name` (e_vpi.out, et_vpi.out, gr_vpi.out, sk_vpi.out, nr_vpi.out, fort.99)"""
        nc = C.c_int(0)
        n = self.L.ref_file(name.encode(), None, 0, C.byref(nc))
        buf = np.zeros((n, max(nc.value, 1)))
        self.L.ref_file(name.encode(), _dp(buf), buf.size, C.byref(nc))
        return buf

    def sgrnd(self, seed):
        self.L.ref_sgrnd(int(seed))

    def grnd(self):
        return self.L.ref_grnd()

    def rangauss(self):
        return self.L.ref_rangauss()

    def get_mt(self):
        mt = np.zeros(624, dtype=np.uint32)
        mti = C.c_int32()
        self.L.ref_get_mt(mt.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(mti))
        return mt, mti.value

    def set_mt(self, mt, mti):
        mt = np.ascontiguousarray(mt, dtype=np.uint32)
        self.L.ref_set_mt(mt.ctypes.data_as(C.POINTER(C.c_uint32)), int(mti))

    def tables(self):
        W, V = np.zeros(self.Nmax + 2), np.zeros(self.Nmax + 2)
        self.L.ref_tables(_dp(W), _dp(V))
        return W, V

    def update_action(self, W, V, Path, ip, ib, xnew, xold, dt=None):
        Path, a, b = _f64(Path), _f64(xnew), _f64(xold)
        return self.L.ref_update_action(_dp(_f64(W)), _dp(_f64(V)), _dp(Path), ip, ib, _dp(a), _dp(b),
                                        self.cfg["dt"] if dt is None else dt, int(self.cfg["trap"]))

    def local_energy(self, W, V, R):
        e = [C.c_double() for _ in range(3)]
        self.L.ref_local_energy(_dp(_f64(W)), _dp(_f64(V)), _dp(_f64(R)), int(self.cfg["trap"]), *[C.byref(x) for x in e])
        return tuple(x.value for x in e)

    def therm_energy(self, V, Path):
        e = [C.c_double() for _ in range(3)]
        self.L.ref_therm_energy(_dp(_f64(V)), _dp(_f64(Path)), self.cfg["dt"], int(self.cfg["trap"]), *[C.byref(x) for x in e])
        return tuple(x.value for x in e)

    def pair_correlation(self, R):
        gr = np.zeros(self.cfg["Nbin"])
        self.L.ref_pair_correlation(_dp(_f64(R)), _dp(gr))
        return gr

    def structure_factor(self, R):
        Sk = np.zeros((self.cfg["Nk"], self.dim))
        self.L.ref_structure_factor(self.cfg["Nk"], _dp(_f64(R)), _dp(Sk))
        return Sk

    def obdm(self, xend):
        nr = np.zeros((self.cfg["Nbin"], self.cfg["Npw"] + 1))
        self.L.ref_obdm(_dp(_f64(xend)), _dp(nr))
        return nr

    def move(self, move, W, V, Path, xend, ip, half=0, isopen=0, delta_cm=None, density=None):
        """runs one move of the reference on copies; returns (accepted, Path, xend, isopen, aux)"""
        P, xe = _f64(Path).copy(), _f64(xend).copy()
        io, aux = C.c_int(int(isopen)), C.c_int(0)
        acc = self.L.ref_move(int(move), int(self.cfg["trap"]), _dp(_f64(W)), _dp(_f64(V)), self.cfg["dt"], float(delta_cm),
                              float(density), int(self.cfg["Lstag"]), int(self.cfg["Nlev"]), int(ip), int(half), _dp(P), _dp(xe),
                              C.byref(io), C.byref(aux))
        return acc, P, xe, io.value, aux.value
