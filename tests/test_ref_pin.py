"""The oracle's pin.  oracle/_ref/libpigs_ref.so is the reference's own Fortran, machine-translated to C++ from
/root/reference/*.f90 by oracle/f90toc/f90toc.py (no Fortran compiler exists in the image) and compiled with the
oracle's flags.  Every test here demands BIT EQUALITY between that translation and the hand-written oracle
(oracle/pigs_oracle.cpp) on the same inputs and the same MT19937 stream: leaf functions, tables, the RNG,
UpdateAction, the estimators, each of the 14 moves, and the whole program `./vpi < vpi.in` block by block.

CPU tests (the reference sources live only in the build container).  Where neither the sources nor a prebuilt
library exist, the tests that need the translation skip, and test_oracle_matches_reference_goldens checks the
oracle against tests/golden/ref_golden.json, the vectors tests/golden/make_ref_golden.py took from the translation.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import pigs_oracle as po
from oracle import pigs_ref
from tests.common import C1, C2, CW, CWX, CS, oracle_cfg, synthetic_path

HAVE_REF = pigs_ref.available()
need_ref = pytest.mark.skipif(not HAVE_REF, reason="neither /root/reference nor oracle/_ref/libpigs_ref.so is present")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_golden.json")


def pair(cfg):
    c = oracle_cfg(cfg)
    r = pigs_ref.Ref(c)
    o = po.Oracle(c)
    o.fill_tables()
    W, V = o.get_tables()
    return c, r, o, W, V


def same(a, b):
    return np.array_equal(np.asarray(a, float), np.asarray(b, float), equal_nan=True)


# ------------------------------------------------------------------ leaves (system_mod, global_mod, interpolate, pbc_mod)
@need_ref
def test_leaf_functions_bit_equal():
    c, r, o, W, V = pair(C2)
    L = r.L
    assert same(r.Lbox, o.Lbox) and r.rcut == o.rcut and r.dr == o.dr and r.rbin == o.rbin
    rng = np.random.default_rng(1)
    x = rng.uniform(0.05, 4.5, 30000)
    assert same([L.ref_potential(v) for v in x], [po.potential(v) for v in x])
    for opt in (0, 1, 2):
        assert same([L.ref_logpsi(opt, 1.2, v) for v in x], [po.logpsi(opt, 1.2, v) for v in x])
    Wr, Vr = r.tables()
    assert same(Wr, W) and same(Vr, V)            # incl. F(0)=F(2), F(Nmax+1)=F(Nmax), -inf / NaN at r = 0
    xs = rng.uniform(2 * r.dr, r.rcut, 100000)
    for opt in (0, 1, 2):
        for T in (V, W):
            assert same([L.ref_interpolate(opt, r.Nmax, r.dr, pigs_ref._dp(T), v) for v in xs[:40000]],
                        [po.interpolate(opt, T, r.dr, v) for v in xs[:40000]])
    for opt in (0, 1):
        for ib in (0, 1, 2, 7, 2 * c["Nb"] - 1, 2 * c["Nb"]):
            for pot, f2 in rng.normal(0, 30, size=(50, 2)):
                assert L.ref_green(opt, ib, c["dt"], pot, f2) == o.green_function(opt, ib, c["dt"], pot, f2)
    for d in rng.uniform(-1.4, 1.4, size=(4000, 3)) * r.Lbox[0]:
        xr = np.array(d)
        r2 = np.zeros(1)
        L.ref_minimum_image(pigs_ref._dp(xr), pigs_ref._dp(r2))
        xo, r2o = o.minimum_image(d)
        assert same(xr, xo) and r2[0] == r2o
        assert L.ref_boundary(1, d[0]) == o.boundary_conditions(1, d[0])
    # the d-ball constant: Cody's Gamma (r8_gamma.f90) against the libm tgamma the oracle uses
    for dim in (1, 2, 3):
        assert abs(L.ref_r8_gamma(0.5 * dim + 1.0) / math.gamma(0.5 * dim + 1.0) - 1.0) < 4e-16


@need_ref
def test_trap_leaves_bit_equal():
    c, r, o, W, V = pair(C1)
    rng = np.random.default_rng(2)
    # TrapPsi / TrapPot enter through UpdateAction / LocalEnergy / ThermEnergy (the oracle exports no separate entry)
    P = synthetic_path(C1, rng)
    for ib in (0, 1, 2, 2 * c["Nb"]):
        for ip in (1, 4, c["Np"]):
            xo = P[ib, ip - 1].copy()
            xn = xo + rng.normal(0, 0.2, 3)
            assert r.update_action(W, V, P, ip, ib, xn, xo) == o.update_action(ip, ib, xn, xo, R=P[ib])
    assert r.local_energy(W, V, P[0]) == o.local_energy(P[0])
    assert r.therm_energy(V, P) == o.therm_energy(P)


# ------------------------------------------------------------------ random_mod.f90
@need_ref
def test_mt19937_and_rangauss_bit_equal():
    c, r, o, W, V = pair(CW)
    for seed in (1982, 4357, 20260101):
        r.sgrnd(seed)
        o.sgrnd(seed)
        assert same([r.grnd() for _ in range(2000)], [o.grnd() for _ in range(2000)])      # across three refills
        assert same([r.rangauss() for _ in range(1500)], [o.rangauss() for _ in range(1500)])
        mr, ir = r.get_mt()
        mo, io = o.get_mt()
        assert np.array_equal(mr, mo) and ir == io
    # the TRANSLATED generator reproduces the published 1998 mt19937int.out head
    r.sgrnd(4357)
    assert [round(r.grnd() * 4294967295.0) for _ in range(2)] == [3510405877, 4290933890]


# ------------------------------------------------------------------ UpdateAction (vpi_mod.f90:2491-2841)
@need_ref
@pytest.mark.parametrize("name,cfg", [("C2", C2), ("CWX", CWX)])
def test_update_action_bit_equal(name, cfg):
    c, r, o, W, V = pair(cfg)
    rng = np.random.default_rng(3)
    P = synthetic_path(cfg, rng)
    S = 2 * c["Nb"] + 1
    L = r.Lbox[0]
    n = 0
    for ib in list(range(S)) * 3:
        ip = int(rng.integers(1, c["Np"] + 1))
        xo = P[ib, ip - 1].copy()
        xn = xo + rng.normal(0, 0.25, 3)
        xn = (xn + L / 2) % L - L / 2
        a, b = r.update_action(W, V, P, ip, ib, xn, xo), o.update_action(ip, ib, xn, xo, R=P[ib])
        assert a == b, (ib, ip, a, b)
        n += 1
    assert n == 3 * S


# ------------------------------------------------------------------ estimators (sample_mod.f90)
@need_ref
def test_estimators_bit_equal():
    c, r, o, W, V = pair(dict(C2, Npw=2))
    rng = np.random.default_rng(4)
    for _ in range(3):
        P = synthetic_path(C2, rng)
        assert r.local_energy(W, V, P[0]) == o.local_energy(P[0])
        assert r.local_energy(W, V, P[-1]) == o.local_energy(P[-1])
        assert r.therm_energy(V, P) == o.therm_energy(P)
        assert same(r.pair_correlation(P[c["Nb"]]), o.pair_correlation(P[c["Nb"]]))
        assert same(r.structure_factor(P[c["Nb"]]), o.structure_factor(P[c["Nb"]]))
    L = r.Lbox[0]
    for _ in range(200):
        xe = rng.uniform(-L / 2, L / 2, size=(2, 3))
        xe[1] = xe[0] + rng.normal(0, 0.9, 3)
        assert same(r.obdm(xe), o.obdm(xe))
    # normalisers (sample_mod.f90:656-732) and Var (:921-932)
    gr = rng.integers(0, 50, c["Nbin"]).astype(float) * 2
    g1, g2 = gr.copy(), gr.copy()
    r.L.ref_normalize_gr(o.density, 7, pigs_ref._dp(g1))
    o.normalize_gr(7, g2)
    assert same(g1, g2)
    Sk = rng.uniform(0, 900, size=(c["Nk"], 3))
    s1, s2 = Sk.copy(), Sk.copy()
    r.L.ref_normalize_sk(c["Nk"], 7, pigs_ref._dp(s1))
    o.normalize_sk(7, s2)
    assert same(s1, s2)
    for n, s, s2 in ((9, 3.3, 12.9), (4, -7.0, 49.5), (1, 2.0, 4.0)):
        assert same(r.L.ref_var(n, s, s2), po.var(n, s, s2))


# ------------------------------------------------------------------ the 14 moves, one call each from identical state + stream
@need_ref
@pytest.mark.parametrize("name,cfg", [("CWX", CWX), ("CS", CS)])
def test_moves_bit_equal(name, cfg):
    c, r, o, W, V = pair(cfg)
    rng = np.random.default_rng(5)
    Nb, Np = c["Nb"], c["Np"]
    naccept = 0
    for mv_name, mv in po.MOVES.items():
        for trial in range(12):
            P = synthetic_path(cfg, rng, spread=0.03)
            ip = int(rng.integers(1, Np + 1))
            half = int(rng.integers(1, 3))
            isopen = 1 if mv >= 7 and mv != 11 else 0
            xe = np.stack([P[Nb, ip - 1], P[Nb, ip - 1] + rng.normal(0, 0.05, 3)]) if isopen else np.stack([P[Nb, -1], P[Nb, -1]])
            seed = int(rng.integers(1, 10 ** 6))
            o.set_state(P, xe, isopen, ip if isopen else 0)
            o.sgrnd(seed)
            r.sgrnd(seed)
            acc_o, aux_o = o.move(mv_name, ip, half)
            acc_r, Pr, xer, io_r, aux_r = r.move(mv, W, V, P, xe, ip, half, isopen, delta_cm=o.delta_cm, density=o.density)
            Po, xeo, io_o, _ = o.get_state()
            assert acc_r == acc_o and aux_r == aux_o and io_r == io_o, (mv_name, trial)
            assert same(Pr, Po) and same(xer, xeo), (mv_name, trial)
            mr, ir = r.get_mt()
            mo, imo = o.get_mt()
            assert np.array_equal(mr, mo) and ir == imo, (mv_name, trial)      # same number of draws consumed
            naccept += acc_o
    assert naccept >= 8          # accepted and rejected branches both ran


# ------------------------------------------------------------------ the whole program
def oracle_program(c, Nblock, Nstep, lattice=None):
    """what `program vpi` writes to e_vpi.out / et_vpi.out, from the oracle's block sums (vpi.f90:477-518)"""
    o = po.Oracle(c)
    o.fill_tables()
    if lattice is None:
        o.init()
    else:                                                       # crystal = T: positions from config_ini.in (vpi_mod.f90:218-228)
        R = np.asarray(lattice, float)
        P = np.broadcast_to(R, (2 * c["Nb"] + 1,) + R.shape).copy()
        o.set_state(P, np.stack([P[c["Nb"], -1]] * 2), 0, 0)
        o.sgrnd(c["seed"])
    Np = c["Np"]
    e, et, nr, events = [], [], None, dict(open=0, close=0, swap=0)
    for ib in range(1, Nblock + 1):
        b, gr, Sk, nr = o.run_block(Nstep, nr)
        for k in events:
            events[k] += b["acc_" + k]
        n = b["idiag_block"]
        if n:
            f = np.float64(np.float32(n))                       # NormalizeAv divides by real(Nitem): single precision
            e.append([float(np.float32(ib))] + [(b[k] / f) / Np for k in ("sumE", "sumK", "sumV")])
            et.append([float(np.float32(ib))] + [(b[k] / f) / Np for k in ("sumEt", "sumKt", "sumVt")])
        if b["idiag_aux"] // Nstep >= 1:
            nr = None
    return np.array(e), np.array(et), events, o


@need_ref
@pytest.mark.parametrize("name,cfg,Nblock,Nstep", [("CW", CW, 4, 25), ("CWX", CWX, 5, 25), ("CS", CS, 3, 20), ("C1", C1, 2, 10),
                                                   ("C2", C2, 2, 2)])
def test_whole_program_bit_equal(name, cfg, Nblock, Nstep):
    c = oracle_cfg(cfg)
    r = pigs_ref.Ref(c, Nblock=Nblock, Nstep=Nstep)
    e, et, events, o = oracle_program(c, Nblock, Nstep)
    re_, ret = r.file("e_vpi.out"), r.file("et_vpi.out")
    assert re_.shape == e.shape and e.shape[0] >= 1
    assert np.array_equal(re_, e) and np.array_equal(ret, et)
    if name == "CWX":
        assert events["open"] > 0 and events["close"] > 0 and events["swap"] > 0      # the worm sector really ran
        hist = r.file("fort.99")[:, 1]
        assert np.array_equal(hist, np.asarray(o.get_perm()[2], float)) and hist.sum() > 0


@need_ref
def test_driver_files_match_reference_program(tmp_path):
    """the drop-in driver layer (driver.py over the oracle) writes, byte for byte, what the translated `program vpi`
    hands to its write statements -- formatted with the same G20.10E3 editing"""
    from pathintegralgroundstate_b200.driver import VpiDriver, g_line
    from tests.oracle_backend import OracleBackend
    Nblock, Nstep = 5, 25
    cfg = dict(CWX, Nblock=Nblock, Nstep=Nstep)
    r = pigs_ref.Ref(oracle_cfg(cfg), Nblock=Nblock, Nstep=Nstep)
    d = VpiDriver(cfg, OracleBackend(cfg), workdir=str(tmp_path), quiet=True)
    d.run(Nblock, Nstep)
    for fname in ("e_vpi.out", "et_vpi.out", "gr_vpi.out", "sk_vpi.out", "nr_vpi.out"):
        want = "".join(g_line(row[:n]) + "\n" for row, n in ((row, len(row)) for row in r.file(fname)))
        got = open(os.path.join(str(tmp_path), fname)).read()
        # records of these files have a fixed width; compare the numbers the text carries
        gw = np.array([[float(x) for x in line.split()] for line in want.splitlines() if line.strip()])
        gg = np.array([[float(x) for x in line.split()] for line in got.splitlines() if line.strip()])
        assert gw.shape == gg.shape and np.array_equal(gw, gg), fname


# ------------------------------------------------------------------ goldens taken from the translation (travel without it)
def test_oracle_matches_reference_goldens():
    G = json.load(open(GOLDEN))
    h = lambda s: float.fromhex(s)
    for case in G["leaf"]:
        f, args, want = case["f"], case["args"], h(case["want"])
        got = {"potential": lambda a: po.potential(h(a[0])),
               "logpsi": lambda a: po.logpsi(a[0], h(a[1]), h(a[2]))}[f](args)
        assert got == want or (math.isnan(got) and math.isnan(want)), case
    for case in G["program"]:
        c = case["cfg"]
        lat = [[h(x) for x in row] for row in case["lattice"]] if "lattice" in case else None
        e, et, _, _ = oracle_program(c, case["Nblock"], case["Nstep"], lat)
        assert [[x.hex() for x in row] for row in e.tolist()] == case["e_vpi"], case["name"]
        assert [[x.hex() for x in row] for row in et.tolist()] == case["et_vpi"], case["name"]
    o = po.Oracle(G["stream"]["cfg"])
    o.sgrnd(G["stream"]["seed"])
    assert [o.grnd().hex() for _ in range(len(G["stream"]["grnd"]))] == G["stream"]["grnd"]
    assert [o.rangauss().hex() for _ in range(len(G["stream"]["rangauss"]))] == G["stream"]["rangauss"]
