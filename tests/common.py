"""Shared fixtures of the parity tests: the BASELINE.json configurations at
test size, synthetic He-4 configurations, and glue that feeds the SAME inputs
to the CPU oracle (checker) and to libpigs_cuda (product)."""
import numpy as np

from oracle.pigs_oracle import Oracle
from pathintegralgroundstate_b200 import PigsCuda

# C1: harmonic trap, N=8 non-interacting bosons (zero tables)      -- BASELINE.json configs[0]
C1 = dict(dim=3, Np=8, density=1.0, trap=True, a_ho=[1.0, 1.0, 1.0], dt=0.05, Nb=20, seed=1982, delta_cm=0.5,
          CMFreq=1, sampling="bis", Lstag=16, Nlev=3, Nstag=5, Nbin=100, Nk=10, swapping=True, CWorm=0.0, Nobdm=1,
          Npw=0, Nmax=10000, wf_table=True, v_table=True, Rm=1.2)
# C2: liquid He-4 N=64, worm "off" (CWorm=0)                       -- configs[1]
C2 = dict(dim=3, Np=64, density=0.365, trap=False, dt=5e-3, Nb=15, seed=1982, delta_cm=0.12, CMFreq=1,
          sampling="bis", Lstag=14, Nlev=3, Nstag=5, Nbin=100, Nk=50, swapping=True, CWorm=0.0, Nobdm=1, Npw=0,
          Nmax=10000, wf_table=True, v_table=True, Rm=1.2)
# the input file the reference ships (vpi.in): N=64, 2M=64 beads, 16-link bisection, 32-link worm moves, worm on
CREF = dict(C2, Nb=32, Lstag=32, Nlev=4, CWorm=0.5, Nobdm=10)
# C3: liquid He-4 N=256, worm on                                   -- configs[2]
C3 = dict(C2, Np=256, CWorm=0.5, Nobdm=10)
# small worm configuration for fast replay tests
# CWorm chosen so that open and close are both accepted often (open ~ CWorm*rho*(2 pi Ls dt)^1.5)
CW = dict(C2, Np=16, Nb=8, Lstag=6, Nlev=2, CWorm=5.0, Nobdm=3, Nstag=2, Nk=8, Nbin=40)
# a large time step makes swap (exchange) moves acceptable: exercises Swap + permutation bookkeeping
CWX = dict(CW, dt=0.04, Lstag=8, CWorm=3.0)
# staging flavour
CS = dict(CWX, sampling="sta")


def oracle_cfg(cfg):
    c = dict(cfg)
    if "action" in c:
        c["action"] = 1 if str(c["action"]).lower().startswith("prim") else 0
    if c.get("crystal") and "Lbox" in c:
        c["Lbox_crystal"] = list(c["Lbox"])
    for k in ("trap", "swapping", "wf_table", "v_table", "crystal"):
        if k in c:
            c[k] = int(bool(c[k]))
    return c


def lattice(Np, L, jitter, rng, dim=3):
    """simple-cubic lattice with >= Np sites centred in [-L/2,L/2), first Np sites, + uniform jitter"""
    n = int(np.ceil(Np ** (1.0 / dim) - 1e-9))
    g = (np.arange(n) + 0.5) * (L / n) - L / 2
    grids = np.meshgrid(*([g] * dim), indexing="ij")
    R = np.stack([x.ravel() for x in grids], axis=1)[:Np]
    return R + rng.uniform(-jitter, jitter, size=R.shape)


def synthetic_path(cfg, rng, spread=0.08, jitter=0.05):
    """a He-4-like configuration: lattice + per-particle jitter, beads scattered around it and wrapped"""
    o = Oracle(oracle_cfg(cfg))
    S = 2 * cfg["Nb"] + 1
    dim = cfg["dim"]
    if cfg.get("trap"):
        R0 = rng.uniform(-1.0, 1.0, size=(cfg["Np"], dim))
        P = R0[None] + rng.normal(0, spread, size=(S, cfg["Np"], dim))
        return P
    L = np.asarray(o.Lbox[:dim])
    if cfg.get("crystal"):
        from pathintegralgroundstate_b200.workloads import hcp_lattice
        R0 = hcp_lattice(density=cfg["density"])[0][:cfg["Np"]] + rng.uniform(-jitter, jitter, size=(cfg["Np"], 3))
    else:
        R0 = lattice(cfg["Np"], L[0], jitter, rng, dim)
    P = R0[None] + rng.normal(0, spread, size=(S, cfg["Np"], dim))
    P = (P + L / 2) % L - L / 2
    return P


def make_pair(cfg, n_chains=1, rng="mt", seed=None, tables="reference", **kw):
    """(oracle, gpu) on the same configuration and the same tables"""
    o = Oracle(oracle_cfg(cfg))
    g = PigsCuda(cfg, n_chains=n_chains, rng=rng, seed=seed, **kw)
    if tables == "zero":
        W = np.zeros(cfg["Nmax"] + 2)
        V = np.zeros(cfg["Nmax"] + 2)
        o.set_tables(W, V)
    else:
        o.fill_tables()
        W, V = o.get_tables()
    g.set_tables(W, V)
    return o, g


def rel_err(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.max(np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor))
