"""Host side of libpigs_cuda: a Python mirror of the reference's module-level
interface (vpi_mod / sample_mod / random_mod procedures the driver vpi.f90
calls) on top of the C ABI in ``include/pigs_cuda.h``.

Nothing here computes on the CPU what the library computes on the GPU: the
class is ctypes plumbing plus the *driver-side* arithmetic the reference keeps
in ``vpi.f90`` (geometry, table fill, normalisation), with its float32 casts.
If ``libpigs_cuda.so`` is missing or no CUDA device exists, construction fails
loudly -- there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PIGS_LIB selects an alternative build of the same library (tuning experiments only)
LIB_PATH = os.environ.get("PIGS_LIB") or os.path.join(_HERE, "libpigs_cuda.so")

PIGS_RNG_PHILOX, PIGS_RNG_MT_REPLAY = 0, 1
MOVES = dict(
    TranslateChain=0, Staging=1, MoveHead=2, MoveTail=3, Bisection=4, MoveHeadBisection=5, MoveTailBisection=6,
    TranslateHalfChain=7, StagingHalfChain=8, MoveHeadHalfChain=9, MoveTailHalfChain=10,
    OpenChain=11, CloseChain=12, Swap=13,
)


class PigsError(RuntimeError):
    pass


class PigsParams(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("Np", C.c_int32), ("Nb", C.c_int32), ("Nmax", C.c_int32),
        ("Nbin", C.c_int32), ("Nk", C.c_int32), ("Npw", C.c_int32), ("trap", C.c_int32),
        ("Lbox", C.c_double * 3), ("a_ho", C.c_double * 3),
        ("rcut", C.c_double), ("dr", C.c_double), ("density", C.c_double), ("dt", C.c_double),
        ("delta_cm", C.c_double), ("CWorm", C.c_double),
        ("CMFreq", C.c_int32), ("sampling", C.c_int32), ("Lstag", C.c_int32), ("Nlev", C.c_int32),
        ("Nstag", C.c_int32), ("Nobdm", C.c_int32), ("swapping", C.c_int32),
        ("n_chains", C.c_int32), ("rng_mode", C.c_int32), ("seed", C.c_uint64),
        ("device", C.c_int32), ("threads_per_chain", C.c_int32), ("table_mode", C.c_int32), ("action", C.c_int32),
        ("schedule", C.c_int32), ("chain_offset", C.c_int32), ("gpus", C.c_int32),
    ]


_BLOCK_F = ("sumE", "sumK", "sumV", "sumEt", "sumKt", "sumVt", "sumE2", "sumK2", "sumV2", "sumEt2", "sumKt2", "sumVt2")
_BLOCK_I = ("idiag_block", "ngr", "try_cm", "try_stag", "try_cm_half", "try_stag_half",
            "acc_cm", "acc_bd", "acc_head", "acc_tail", "acc_cm_half", "acc_bd_half", "acc_head_half", "acc_tail_half",
            "try_open", "acc_open", "try_close", "acc_close", "try_swap", "acc_swap")


class PigsBlockResult(C.Structure):
    _fields_ = [(n, C.c_double) for n in _BLOCK_F] + [(n, C.c_int64) for n in _BLOCK_I] + \
               [("bead_updates", C.c_int64 * 3), ("n_open_chains", C.c_int64)]

    def as_dict(self):
        d = {n: getattr(self, n) for n in _BLOCK_F + _BLOCK_I}
        d["bead_updates"] = list(self.bead_updates)
        d["n_open_chains"] = self.n_open_chains
        return d


_lib = None


def load_library() -> C.CDLL:
    """dlopen libpigs_cuda.so (built in-tree by __graft_entry__.build / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("PIGS_CUDA_LIB", LIB_PATH)      # tuning builds (e.g. another launch bound) for A/B runs
    if not os.path.exists(path):
        raise PigsError(f"{path} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                        "(there is no CPU fallback)")
    L = C.CDLL(path)
    dp, i32p, u32p = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32)
    H = C.c_void_p
    ip = C.POINTER(C.c_int)
    sig = {
        "pigs_last_error": (C.c_char_p, []),
        "pigs_version": (C.c_int, []),
        "pigs_create": (C.c_int, [C.POINTER(PigsParams), C.POINTER(H)]),
        "pigs_destroy": (C.c_int, [H]),
        "pigs_set_tables": (C.c_int, [H, dp, dp]),
        "pigs_set_state": (C.c_int, [H, C.c_int, dp, dp, C.c_int, C.c_int]),
        "pigs_get_state": (C.c_int, [H, C.c_int, dp, dp, ip, ip]),
        "pigs_set_state_all": (C.c_int, [H, dp, dp, i32p, i32p]),
        "pigs_get_state_all": (C.c_int, [H, dp, dp, i32p, i32p]),
        "pigs_get_perm": (C.c_int, [H, C.c_int, ip, i32p, i32p, ip, ip]),
        "pigs_set_perm": (C.c_int, [H, C.c_int, C.c_int, i32p, i32p, C.c_int, C.c_int]),
        "pigs_sgrnd": (C.c_int, [H, C.c_int, C.c_int32]),
        "pigs_get_mt": (C.c_int, [H, C.c_int, u32p, i32p]),
        "pigs_set_mt": (C.c_int, [H, C.c_int, u32p, C.c_int32]),
        "pigs_grnd": (C.c_int, [H, C.c_int, C.c_int, dp]),
        "pigs_rangauss": (C.c_int, [H, C.c_int, C.c_int, dp]),
        "pigs_run_block": (C.c_int, [H, C.c_int]),
        "pigs_run_block_async": (C.c_int, [H, C.c_int]),
        "pigs_sync": (C.c_int, [H]),
        "pigs_get_block": (C.c_int, [H, C.POINTER(PigsBlockResult), dp, dp, dp]),
        "pigs_get_block_chain": (C.c_int, [H, C.c_int, C.POINTER(PigsBlockResult), dp, dp, dp]),
        "pigs_get_block_chains": (C.c_int, [H, C.c_int, C.c_int, C.POINTER(PigsBlockResult), dp, dp, dp]),
        "pigs_block_vector": (C.c_int, [H, C.POINTER(C.c_void_p), ip]),
        "pigs_unpack_block_vector": (C.c_int, [H, dp, C.POINTER(PigsBlockResult), dp, dp, dp]),
        "pigs_last_block_ms": (C.c_int, [H, C.POINTER(C.c_float)]),
        "pigs_stream": (C.c_int, [H, C.POINTER(C.c_void_p)]),
        "pigs_launch_count": (C.c_int, [H, C.POINTER(C.c_int64)]),
        "pigs_launch_plan": (C.c_int, [H, ip, ip, ip, ip, ip]),
        "pigs_save_checkpoint": (C.c_int, [H, C.c_char_p]),
        "pigs_load_checkpoint": (C.c_int, [H, C.c_char_p]),
        "pigs_move": (C.c_int, [H, C.c_int, C.c_int, C.c_int, i32p, i32p]),
        "pigs_update_action": (C.c_int, [H, C.c_int, dp, i32p, i32p, dp, dp, dp]),
        "pigs_local_energy": (C.c_int, [H, C.c_int, dp, dp, dp, dp]),
        "pigs_therm_energy": (C.c_int, [H, C.c_int, dp, dp, dp, dp]),
        "pigs_pair_correlation": (C.c_int, [H, C.c_int, dp, dp]),
        "pigs_structure_factor": (C.c_int, [H, C.c_int, dp, dp]),
        "pigs_obdm": (C.c_int, [H, C.c_int, dp, dp]),
        "pigs_measure_fp64_peak": (C.c_int, [C.c_int, dp]),
    }
    for k, (res, args) in sig.items():
        f = getattr(L, k)
        f.restype, f.argtypes = res, args
    L._pigs_symbols = tuple(sig)
    _lib = L
    return L


def _f32(x) -> float:
    return float(np.float32(x))


# ------------------------------------------------------------------ driver-side arithmetic (vpi.f90)
def derive_geometry(cfg: dict) -> dict:
    """Box, cutoff, table step, density and CM step exactly as the driver derives
    them (vpi.f90:80-128, vpi_mod.f90:94) including the default-kind real() casts."""
    dim, Np = int(cfg["dim"]), int(cfg["Np"])
    pi = math.acos(-1.0)
    trap = bool(cfg.get("trap", False))
    delta_cm = float(cfg["delta_cm"])
    out = dict(dim=dim, Np=Np, trap=trap)
    if trap:
        a = [float(x) for x in cfg["a_ho"]][:dim]
        rcut = 1.0
        for k in range(dim):
            rcut = 3.0 * rcut * a[k]
        density = _f32(Np) / (pi ** (0.5 * dim) * rcut / math.gamma(0.5 * dim + 1.0))
        rcut = rcut ** (1.0 / _f32(dim))
        rcut = 10.0 * rcut
        delta_cm = delta_cm * min(a)
        Lbox = [0.0, 0.0, 0.0]
        out["a_ho"] = a + [1.0] * (3 - dim)
    else:
        density = float(cfg["density"])
        if cfg.get("crystal", False):
            Lbox = [float(x) for x in cfg["Lbox"]][:dim]       # config_ini.in line 2 (vpi.f90:104)
        else:
            Lbox = [(_f32(Np) / density) ** (1.0 / _f32(dim)) for _ in range(dim)]
        rcut = min(0.5 * L for L in Lbox)
        delta_cm = delta_cm / density ** (1.0 / _f32(dim))
        Lbox = Lbox + [0.0] * (3 - dim)
        out["a_ho"] = [1.0, 1.0, 1.0]
    Nmax = int(cfg.get("Nmax", 10000))
    out.update(Lbox=Lbox, rcut=rcut, rcut2=rcut * rcut, rbin=rcut / _f32(int(cfg["Nbin"])), density=density,
               delta_cm=delta_cm, dr=rcut / _f32(Nmax - 1), Nmax=Nmax, pi=pi)
    return out


def aziz_hfdb(r):
    """Aziz II HFD-B(HE) pair potential in the reference's reduced units (system_mod.f90:136-182)."""
    E_0, rm, A, alpha, beta = 10.948, 2.963, 1.8443101e5, 10.43329537, -2.27965105
    C6, C8, C10, D = 1.36745214, 0.42123807, 0.17473318, 1.4826
    V0 = E_0 / 1.85505153154686
    r = np.asarray(r, dtype=np.float64)
    with np.errstate(all="ignore"):
        d = r * 2.556 / rm
        d2 = d * d
        d4 = d2 * d2
        d6 = d4 * d2
        Hx = np.where(d <= D, np.exp(-(D / d - 1.0) ** 2), 1.0)
        return V0 * (A * np.exp(-alpha * d + beta * d2) - (C6 + C8 / d2 + C10 / d4) * Hx / d6)


def aziz_hfdhe2(r):
    """Aziz I HFDHE2 (the commented-out alternative, system_mod.f90:87-132)."""
    E_0, rm, A, alpha = 10.8, 2.9673, 0.54485046e6, 13.353384
    C6, C8, C10, D = 1.3732412, 0.4253785, 0.1781, 1.241314
    V0 = E_0 / 1.85505153154686
    r = np.asarray(r, dtype=np.float64)
    with np.errstate(all="ignore"):
        d = r * 2.556 / rm
        d2 = d * d
        d4 = d2 * d2
        d6 = d4 * d2
        Hx = np.where(d <= D, np.exp(-(D / d - 1.0) ** 2), 1.0)
        return V0 * (A * np.exp(-alpha * d) - (C6 + C8 / d2 + C10 / d4) * Hx / d6)


def lennard_jones(r):
    """Lennard-Jones in reduced units, V0 (1/r^6 - 1)/r^6 with V0 = 22.0228 (the first commented-out
    alternative, system_mod.f90:70-83)."""
    r = np.asarray(r, dtype=np.float64)
    with np.errstate(all="ignore"):
        r6 = r ** 6
        return 22.0228 * (1.0 / r6 - 1.0) / r6


def mcmillan_logpsi(r, Rm):
    """LogPsi(0,Rm,r) = -0.5 (Rm/r)**5 (system_mod.f90:38-66)."""
    r = np.asarray(r, dtype=np.float64)
    with np.errstate(all="ignore"):
        q = Rm / r
        q2 = q * q
        return -0.5 * ((q * q2) * q2)        # x**5 as gfortran expands it: (x * x^2) * x^2


def make_table(f, rmax: float, Nmax: int, shift_free: bool = False) -> np.ndarray:
    """JastrowTable / PotentialTable (vpi_mod.f90:84-145): entry i holds f((i-1)*dr),
    i = 1..Nmax, pads F(0)=F(2), F(Nmax+1)=F(Nmax).  (The lookup then evaluates
    f(r-dr); that shift is the reference's and is inherited on purpose.)

    shift_free=True is an input the reference lacks: entry i holds f(i*dr), so that
    Interpolate(0,...) returns f(r) itself; the kernels are unchanged."""
    dr = rmax / _f32(Nmax - 1)
    F = np.zeros(Nmax + 2)
    if shift_free:
        i = np.arange(0, Nmax + 2)
        F[:] = f(i * dr)
        return F
    i = np.arange(1, Nmax + 1)
    F[1:Nmax + 1] = f((i - 1) * dr)
    F[0] = F[2]
    F[Nmax + 1] = F[Nmax]
    return F


def write_config_ini(path: str, R, Lbox, density: float):
    """config_ini.in as the driver and `init` read it (vpi.f90:101-107, vpi_mod.f90:220-228):
    line 1 Np, line 2 Lbox(1:dim), line 3 density, then Np position lines."""
    R = np.asarray(R, dtype=np.float64)
    with open(path, "w") as f:
        f.write(f"{R.shape[0]}\n")
        f.write(" ".join(f"{x:.16e}" for x in Lbox[:R.shape[1]]) + "\n")
        f.write(f"{density:.16e}\n")
        for r in R:
            f.write(" ".join(f"{x:.16e}" for x in r) + "\n")


def read_config_ini(path: str, dim: int = 3):
    with open(path) as f:
        Np = int(f.readline().split()[0])
        Lbox = [float(t.replace("d", "e").replace("D", "e")) for t in f.readline().split()[:dim]]
        density = float(f.readline().split()[0].replace("d", "e").replace("D", "e"))
        R = np.array([[float(t.replace("d", "e").replace("D", "e")) for t in f.readline().split()[:dim]] for _ in range(Np)])
    return Np, Lbox, density, R


def kn_ball(dim: int) -> float:
    """pi**(dim/2)/Gamma(dim/2+1) (sample_mod.f90:669,721)"""
    return math.acos(-1.0) ** (0.5 * dim) / math.gamma(0.5 * dim + 1.0)


def normalize_gr(gr, geo: dict, Np: int, ngr_total: float):
    """NormalizeGr (sample_mod.f90:656-679); ngr_total = ngr summed over chains."""
    dim, rbin, density = geo["dim"], geo["rbin"], geo["density"]
    k_n = kn_ball(dim)
    j = np.arange(1, len(gr) + 1, dtype=np.float64)
    r = (j - 0.5) * rbin
    nid = density * k_n * ((r + 0.5 * rbin) ** dim - (r - 0.5 * rbin) ** dim)
    return np.asarray(gr) / (nid * (float(Np) * float(ngr_total)))


def normalize_sk(Sk, Np: int, ngr_total: float):
    """NormalizeSk (sample_mod.f90:683-702)"""
    return np.asarray(Sk) / (float(Np) * float(ngr_total))


def normalize_nr(nrho, geo: dict, CWorm: float, zconf: float, Nobdm: int):
    """NormalizeNr (sample_mod.f90:706-732); nrho[Nbin][Npw+1]"""
    dim, rbin, density = geo["dim"], geo["rbin"], geo["density"]
    k_n = kn_ball(dim)
    nrho = np.asarray(nrho, dtype=np.float64)
    j = np.arange(1, nrho.shape[0] + 1, dtype=np.float64)
    r = (j - 0.5) * rbin
    nid = density * k_n * ((r + 0.5 * rbin) ** dim - (r - 0.5 * rbin) ** dim)
    return nrho / (CWorm * nid * zconf * float(Nobdm))[:, None]


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class PigsCuda:
    """n_chains replicas of the reference program's state on one B200.

    ``cfg`` uses the variable names of vpi.in (see vpi_in.read_vpi_in).  Array
    arguments use the C-order views of the Fortran arrays:
    Path[2Nb+1][Np][dim], xend[2][dim], R[Np][dim]."""

    def __init__(self, cfg: dict, n_chains: int = 1, rng: str = "philox", seed: int | None = None, device: int = 0,
                 threads_per_chain: int = 0, table_mode: int = -1, schedule: int = -1, chain_offset: int = 0,
                 gpus: int = 1):
        self.L = load_library()
        self.cfg = dict(cfg)
        self.geo = derive_geometry(cfg)
        g = self.geo
        p = PigsParams()
        p.dim, p.Np, p.Nb, p.Nmax = g["dim"], g["Np"], int(cfg["Nb"]), g["Nmax"]
        p.Nbin, p.Nk, p.Npw = int(cfg["Nbin"]), int(cfg.get("Nk", 0)), int(cfg.get("Npw", 0))
        p.trap = 1 if g["trap"] else 0
        p.Lbox = (C.c_double * 3)(*g["Lbox"])
        p.a_ho = (C.c_double * 3)(*g["a_ho"])
        p.rcut, p.dr, p.density, p.dt = g["rcut"], g["dr"], g["density"], float(cfg["dt"])
        p.delta_cm, p.CWorm = g["delta_cm"], float(cfg.get("CWorm", 0.0))
        p.CMFreq = int(cfg["CMFreq"])
        p.sampling = 0 if str(cfg["sampling"]).strip().lower().startswith("sta") else 1
        p.Lstag, p.Nlev = int(cfg.get("Lstag", 2)), int(cfg.get("Nlev", 1))
        p.Nstag, p.Nobdm = int(cfg["Nstag"]), int(cfg.get("Nobdm", 0))
        p.swapping = 1 if cfg.get("swapping", False) else 0
        p.n_chains = int(n_chains)
        p.rng_mode = PIGS_RNG_MT_REPLAY if str(rng).lower().startswith("mt") else PIGS_RNG_PHILOX
        p.seed = int(cfg.get("seed", 1982) if seed is None else seed)
        p.device, p.threads_per_chain, p.table_mode = int(device), int(threads_per_chain), int(table_mode)
        p.action = 1 if str(cfg.get("action", "chin")).lower().startswith("prim") else 0
        p.schedule, p.chain_offset, p.gpus = int(schedule), int(chain_offset), int(gpus)
        self.p = p
        self.dim, self.Np, self.Nb, self.Nmax = p.dim, p.Np, p.Nb, p.Nmax
        self.Nbin, self.Nk, self.Npw, self.n_chains = p.Nbin, p.Nk, p.Npw, p.n_chains
        self.h = C.c_void_p()
        self._ck(self.L.pigs_create(C.byref(p), C.byref(self.h)))

    # -- plumbing
    def _ck(self, rc):
        if rc != 0:
            raise PigsError(f"libpigs_cuda error {rc}: {self.L.pigs_last_error().decode()}")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.pigs_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def path_shape(self):
        return (2 * self.Nb + 1, self.Np, self.dim)

    # -- JastrowTable / PotentialTable (vpi_mod.f90:84-145)
    def fill_tables(self, potential="hfdb", Rm=None, shift_free=False):
        Rm = float(self.cfg["Rm"]) if Rm is None else Rm
        pot = dict(hfdb=aziz_hfdb, hfdhe2=aziz_hfdhe2, lj=lennard_jones, zero=lambda r: np.zeros_like(r))[potential]
        W = make_table(lambda r: mcmillan_logpsi(r, Rm), self.geo["rcut"], self.Nmax, shift_free)
        V = make_table(pot, self.geo["rcut"], self.Nmax, shift_free)
        self.set_tables(W, V)
        return W, V

    def set_tables(self, LogWF, VTable):
        a, b = _f64(LogWF), _f64(VTable)
        if a.size != self.Nmax + 2 or b.size != self.Nmax + 2:
            raise ValueError("tables must have Nmax+2 entries (0:Nmax+1)")
        self._ck(self.L.pigs_set_tables(self.h, _dp(a), _dp(b)))

    # -- init / CheckPoint state
    def set_state(self, chain, Path, xend, isopen=0, iworm=0):
        P, xe = _f64(Path), _f64(xend)
        if P.size != int(np.prod(self.path_shape())) or xe.size != 2 * self.dim:
            raise ValueError("bad Path/xend shape")
        self._ck(self.L.pigs_set_state(self.h, int(chain), _dp(P), _dp(xe), int(isopen), int(iworm)))

    def get_state(self, chain):
        P = np.zeros(self.path_shape())
        xe = np.zeros((2, self.dim))
        io, iw = C.c_int(), C.c_int()
        self._ck(self.L.pigs_get_state(self.h, int(chain), _dp(P), _dp(xe), C.byref(io), C.byref(iw)))
        return P, xe, io.value, iw.value

    def set_state_all(self, Path, xend, isopen=None, iworm=None):
        P, xe = _f64(Path), _f64(xend)
        n = self.n_chains
        if P.size != n * int(np.prod(self.path_shape())) or xe.size != n * 2 * self.dim:
            raise ValueError("bad Path/xend shape")
        io = np.zeros(n, dtype=np.int32) if isopen is None else np.ascontiguousarray(isopen, dtype=np.int32)
        iw = np.zeros(n, dtype=np.int32) if iworm is None else np.ascontiguousarray(iworm, dtype=np.int32)
        self._ck(self.L.pigs_set_state_all(self.h, _dp(P), _dp(xe), _i32p(io), _i32p(iw)))

    def get_state_all(self, want_path=True, out=None):
        """All chains' state.  `out` = (Path, xend, isopen, iworm) caller buffers (e.g. pinned host memory)
        to download into without intermediate copies."""
        n = self.n_chains
        if out is not None:
            P, xe, io, iw = out
            assert P.flags.c_contiguous and P.dtype == np.float64 and P.size == n * int(np.prod(self.path_shape()))
            assert xe.dtype == np.float64 and xe.size == n * 2 * self.dim and io.dtype == np.int32 and iw.dtype == np.int32
        else:
            P = np.zeros((n,) + self.path_shape()) if want_path else None
            xe = np.zeros((n, 2, self.dim))
            io = np.zeros(n, dtype=np.int32)
            iw = np.zeros(n, dtype=np.int32)
        self._ck(self.L.pigs_get_state_all(self.h, _dp(P) if P is not None else None, _dp(xe), _i32p(io), _i32p(iw)))
        return P, xe, io, iw

    def get_perm(self, chain):
        cyc = np.zeros(self.Np, dtype=np.int32)
        hist = np.zeros(self.Np, dtype=np.int32)
        ipm, a, b = C.c_int(), C.c_int(), C.c_int()
        self._ck(self.L.pigs_get_perm(self.h, int(chain), C.byref(ipm), _i32p(cyc), _i32p(hist), C.byref(a), C.byref(b)))
        return ipm.value, cyc, hist, a.value, b.value

    def set_perm(self, chain, iperm, cyc, hist, new_pc=0, end_pc=0):
        cyc = np.ascontiguousarray(cyc, dtype=np.int32)
        hist = np.ascontiguousarray(hist, dtype=np.int32)
        self._ck(self.L.pigs_set_perm(self.h, int(chain), int(iperm), _i32p(cyc), _i32p(hist), int(new_pc), int(end_pc)))

    # -- random_mod
    def sgrnd(self, seed, chain=-1):
        self._ck(self.L.pigs_sgrnd(self.h, int(chain), int(seed)))

    def get_mt(self, chain):
        mt = np.zeros(624, dtype=np.uint32)
        mti = C.c_int32()
        self._ck(self.L.pigs_get_mt(self.h, int(chain), mt.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(mti)))
        return mt, mti.value

    def set_mt(self, chain, mt, mti):
        mt = np.ascontiguousarray(mt, dtype=np.uint32)
        self._ck(self.L.pigs_set_mt(self.h, int(chain), mt.ctypes.data_as(C.POINTER(C.c_uint32)), int(mti)))

    def grnd(self, n, chain=0):
        out = np.zeros(int(n))
        self._ck(self.L.pigs_grnd(self.h, int(chain), int(n), _dp(out)))
        return out

    def rangauss(self, n, chain=0):
        out = np.zeros(int(n))
        self._ck(self.L.pigs_rangauss(self.h, int(chain), int(n), _dp(out)))
        return out

    # -- the driver's step loop
    def run_block(self, Nstep, sync=True):
        if sync:
            self._ck(self.L.pigs_run_block(self.h, int(Nstep)))
        else:
            self._ck(self.L.pigs_run_block_async(self.h, int(Nstep)))

    def sync(self):
        self._ck(self.L.pigs_sync(self.h))

    def last_block_ms(self):
        ms = C.c_float()
        self._ck(self.L.pigs_last_block_ms(self.h, C.byref(ms)))
        return ms.value

    def get_block(self, chain=None):
        b = PigsBlockResult()
        gr = np.zeros(self.Nbin)
        Sk = np.zeros((max(self.Nk, 1), self.dim))
        nr = np.zeros((self.Nbin, self.Npw + 1))
        if chain is None:
            self._ck(self.L.pigs_get_block(self.h, C.byref(b), _dp(gr), _dp(Sk), _dp(nr)))
        else:
            self._ck(self.L.pigs_get_block_chain(self.h, int(chain), C.byref(b), _dp(gr), _dp(Sk), _dp(nr)))
        return b.as_dict(), gr, Sk[:self.Nk], nr

    def get_block_chains(self, chain0=0, n=None):
        """per-chain results of the last block in one transfer: (dict of arrays [n], gr[n][Nbin], Sk[n][Nk][dim],
        nrho[n][Nbin][Npw+1])"""
        n = self.n_chains - chain0 if n is None else int(n)
        out = (PigsBlockResult * n)()
        gr = np.zeros((n, self.Nbin))
        Sk = np.zeros((n, max(self.Nk, 1), self.dim))
        nr = np.zeros((n, self.Nbin, self.Npw + 1))
        self._ck(self.L.pigs_get_block_chains(self.h, int(chain0), n, out, _dp(gr), _dp(Sk), _dp(nr)))
        d = {k: np.array([getattr(o, k) for o in out]) for k in _BLOCK_F + _BLOCK_I}
        d["bead_updates"] = np.array([list(o.bead_updates) for o in out])
        d["n_open_chains"] = np.array([o.n_open_chains for o in out])
        return d, gr, Sk[:, :self.Nk], nr

    def save_checkpoint(self, path):
        """every chain's complete device state (paths, worm/permutation state, RNG) in one binary file"""
        self._ck(self.L.pigs_save_checkpoint(self.h, str(path).encode()))

    def load_checkpoint(self, path):
        self._ck(self.L.pigs_load_checkpoint(self.h, str(path).encode()))

    def launch_plan(self):
        v = [C.c_int() for _ in range(5)]
        self._ck(self.L.pigs_launch_plan(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("threads_per_chain", "groups_per_cta", "grid", "team", "table_mode"), (x.value for x in v)))

    def schedule_name(self):
        p = self.launch_plan()
        if self.p.rng_mode == PIGS_RNG_MT_REPLAY:
            return f"reference order (MT19937 replay), {p['threads_per_chain']} threads per chain"
        return ("team: 4 window workers per chain" if p["team"] else f"window order, {p['threads_per_chain']} threads per chain")

    def block_vector(self):
        """(device pointer, length in doubles) of the chain-summed accumulator vector"""
        ptr, n = C.c_void_p(), C.c_int()
        self._ck(self.L.pigs_block_vector(self.h, C.byref(ptr), C.byref(n)))
        return ptr.value, n.value

    def unpack_block_vector(self, vec):
        v = _f64(vec)
        b = PigsBlockResult()
        gr = np.zeros(self.Nbin)
        Sk = np.zeros((max(self.Nk, 1), self.dim))
        nr = np.zeros((self.Nbin, self.Npw + 1))
        self._ck(self.L.pigs_unpack_block_vector(self.h, _dp(v), C.byref(b), _dp(gr), _dp(Sk), _dp(nr)))
        return b.as_dict(), gr, Sk[:self.Nk], nr

    def stream(self):
        s = C.c_void_p()
        self._ck(self.L.pigs_stream(self.h, C.byref(s)))
        return s.value

    def launch_count(self):
        n = C.c_int64()
        self._ck(self.L.pigs_launch_count(self.h, C.byref(n)))
        return n.value

    # -- one reference procedure per call (unit API)
    def move(self, name, ip, half=0):
        acc = np.zeros(self.n_chains, dtype=np.int32)
        aux = np.zeros(self.n_chains, dtype=np.int32)
        self._ck(self.L.pigs_move(self.h, MOVES[name], int(ip), int(half), _i32p(acc), _i32p(aux)))
        return acc, aux

    def update_action(self, R, ip, ib, xnew, xold):
        R = _f64(R)
        ip = np.atleast_1d(np.ascontiguousarray(ip, dtype=np.int32))
        ib = np.atleast_1d(np.ascontiguousarray(ib, dtype=np.int32))
        n = ip.size
        xn, xo = _f64(xnew), _f64(xold)
        if R.size != n * self.Np * self.dim or xn.size != n * self.dim or xo.size != n * self.dim or ib.size != n:
            raise ValueError("bad shapes")
        out = np.zeros(n)
        self._ck(self.L.pigs_update_action(self.h, n, _dp(R), _i32p(ip), _i32p(ib), _dp(xn), _dp(xo), _dp(out)))
        return out

    def local_energy(self, R):
        R = _f64(R)
        n = R.size // (self.Np * self.dim)
        E, K, V = np.zeros(n), np.zeros(n), np.zeros(n)
        self._ck(self.L.pigs_local_energy(self.h, n, _dp(R), _dp(E), _dp(K), _dp(V)))
        return E, K, V

    def therm_energy(self, Path):
        P = _f64(Path)
        n = P.size // int(np.prod(self.path_shape()))
        E, Ec, Ep = np.zeros(n), np.zeros(n), np.zeros(n)
        self._ck(self.L.pigs_therm_energy(self.h, n, _dp(P), _dp(E), _dp(Ec), _dp(Ep)))
        return E, Ec, Ep

    def pair_correlation(self, R, gr=None):
        R = _f64(R)
        n = R.size // (self.Np * self.dim)
        g = np.zeros((n, self.Nbin)) if gr is None else _f64(gr)
        self._ck(self.L.pigs_pair_correlation(self.h, n, _dp(R), _dp(g)))
        return g

    def structure_factor(self, R, Sk=None):
        R = _f64(R)
        n = R.size // (self.Np * self.dim)
        s = np.zeros((n, self.Nk, self.dim)) if Sk is None else _f64(Sk)
        self._ck(self.L.pigs_structure_factor(self.h, n, _dp(R), _dp(s)))
        return s

    def obdm(self, xend, nrho=None):
        x = _f64(xend)
        n = x.size // (2 * self.dim)
        r = np.zeros((n, self.Nbin, self.Npw + 1)) if nrho is None else _f64(nrho)
        self._ck(self.L.pigs_obdm(self.h, n, _dp(x), _dp(r)))
        return r


def measure_fp64_peak(device: int = 0) -> float:
    L = load_library()
    t = C.c_double()
    rc = L.pigs_measure_fp64_peak(int(device), C.byref(t))
    if rc != 0:
        raise PigsError(L.pigs_last_error().decode())
    return t.value
