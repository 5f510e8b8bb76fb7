"""The BASELINE.json configurations and their synthetic
inputs.  Reduced units of the reference: lengths in sigma = 2.556 A, energies in
hbar^2/(m sigma^2) = 1.85505 K; rho = 0.02186 A^-3 == 0.365 sigma^-3."""
from __future__ import annotations

import numpy as np

from .host import derive_geometry

_COMMON = dict(dim=3, Nmax=10000, wf_table=True, v_table=True, Rm=1.2, swapping=True, seed=1982, Npw=0, Nbin=100)

CONFIGS = {
    # configs[0]: 3D harmonic oscillator, N=8 non-interacting bosons (zero tables), 2M=40 beads, tau=0.05
    "C1": dict(_COMMON, Np=8, density=1.0, trap=True, a_ho=[1.0, 1.0, 1.0], dt=0.05, Nb=20, delta_cm=0.5, CMFreq=1,
               sampling="bis", Lstag=16, Nlev=3, Nstag=5, Nk=10, CWorm=0.0, Nobdm=1, tables="zero"),
    # configs[1]: liquid He-4 N=64 at rho=0.02186 A^-3, Aziz + McMillan, 2M=30 beads, Chin, worm off (CWorm=0)
    "C2": dict(_COMMON, Np=64, density=0.365, trap=False, dt=5e-3, Nb=15, delta_cm=0.12, CMFreq=1, sampling="bis",
               Lstag=14, Nlev=3, Nstag=5, Nk=50, CWorm=0.0, Nobdm=1, tables="hfdb"),
    # configs[2]: liquid He-4 N=256, worm algorithm on, OBDM + g(r) + S(k)       <- the metric's configuration
    "C3": dict(_COMMON, Np=256, density=0.365, trap=False, dt=5e-3, Nb=15, delta_cm=0.12, CMFreq=1, sampling="bis",
               Lstag=14, Nlev=3, Nstag=5, Nk=50, CWorm=0.5, Nobdm=10, tables="hfdb"),
    # configs[3]: solid hcp He-4 N=180 at rho=0.029 A^-3 (0.48426 sigma^-3), worm OBDM
    "C4": dict(_COMMON, Np=180, density=0.48426, trap=False, crystal=True, dt=5e-3, Nb=15, delta_cm=0.12, CMFreq=1,
               sampling="bis", Lstag=14, Nlev=3, Nstag=5, Nk=50, CWorm=0.5, Nobdm=10, tables="hfdb"),
}
# configs[4] is C2 with 4096 chains over 1/2/4/8 GPUs
C5_CHAINS = 4096


def hcp_lattice(nx=5, ny=3, nz=3, density=0.48426):
    """hcp as an orthorhombic 4-atom cell a x sqrt(3)a x sqrt(8/3)a; returns (R[N][3] centred, Lbox[3])."""
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 1.0 / 6.0, 0.5], [0.0, 2.0 / 3.0, 0.5]])
    n = 4 * nx * ny * nz
    vol = n / density
    a = (vol / (nx * ny * nz * np.sqrt(3.0) * np.sqrt(8.0 / 3.0))) ** (1.0 / 3.0)
    cell = np.array([a, np.sqrt(3.0) * a, np.sqrt(8.0 / 3.0) * a])
    L = cell * np.array([nx, ny, nz])
    R = np.array([(np.array([i, j, k]) + b) * cell for i in range(nx) for j in range(ny) for k in range(nz) for b in basis])
    return R - L / 2 + 1e-9, L


def config(name: str) -> dict:
    cfg = dict(CONFIGS[name])
    if cfg.get("crystal"):
        _, L = hcp_lattice(density=cfg["density"])
        cfg["Lbox"] = list(L)
    return cfg


def lattice_sites(cfg: dict) -> np.ndarray:
    """Starting sites: simple-cubic/fcc-like lattice filling the box (liquids), hcp (C4), uniform in +-a (trap)."""
    geo = derive_geometry(cfg)
    Np = cfg["Np"]
    if cfg.get("trap"):
        rng = np.random.default_rng(12345)
        return rng.uniform(-1.0, 1.0, size=(Np, 3)) * np.asarray(geo["a_ho"])
    if cfg.get("crystal"):
        R, _ = hcp_lattice(density=cfg["density"])
        return R[:Np]
    L = np.asarray(geo["Lbox"])
    n = int(np.ceil(Np ** (1.0 / 3.0) - 1e-9))
    g = (np.arange(n) + 0.5) / n - 0.5
    R = np.array([[x, y, z] for x in g for y in g for z in g])[:Np] * L
    return R


def synthetic_paths(cfg: dict, n_chains: int, seed: int = 20260101, jitter: float = 0.05, spread: float = 0.03):
    """Path[n_chains][2Nb+1][Np][3], xend[n_chains][2][3]: lattice + per-particle jitter (as `init` puts all
    beads of a particle on one point, vpi_mod.f90:242-248) + a small per-bead spread, wrapped into the box."""
    geo = derive_geometry(cfg)
    rng = np.random.default_rng(seed)
    S, Np = 2 * cfg["Nb"] + 1, cfg["Np"]
    R0 = lattice_sites(cfg)
    P = np.empty((n_chains, S, Np, 3))
    P[:] = R0[None, None]
    P += rng.uniform(-jitter, jitter, size=(n_chains, 1, Np, 3))
    P += rng.normal(0.0, spread, size=P.shape)
    if not cfg.get("trap"):
        L = np.asarray(geo["Lbox"])
        P = (P + L / 2) % L - L / 2
    xend = np.repeat(P[:, cfg["Nb"], -1][:, None, :], 2, axis=1).copy()
    return P, xend


def flops_per_bead_update(Np: int) -> tuple[float, float, float]:
    """Algorithmic FP64 flops per UpdateAction by slice class (even, odd, end): BASELINE.md section 2,
    FMA = 2, compare/select = div = sqrt = 1:  2 (N-1) c + 40 with c = 28 / 46 / 37."""
    return tuple(2.0 * (Np - 1) * c + 40.0 for c in (28.0, 46.0, 37.0))
