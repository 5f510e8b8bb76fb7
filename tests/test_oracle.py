"""The CPU oracle against the only anchors that exist (the reference ships no
tests, golden vectors or fixtures -- SURVEY.md section 8(c)):
published MT19937 outputs, numpy's independent MT19937 implementation, analytic
values of the potential / Jastrow / trap energy, and self-consistency."""
import json
import os

import numpy as np
import pytest

from oracle.pigs_oracle import Oracle, interpolate, potential, logpsi, var
from tests.common import C1, C2, CW, oracle_cfg, synthetic_path

GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")


def test_mt19937_published_head_and_numpy_cross_check():
    o = Oracle(oracle_cfg(CW))
    o.sgrnd(4357)          # the 1998 reference output mt19937int.out starts 3510405877 4290933890 ...
    assert [o.mt_raw() for _ in range(5)] == [3510405877, 4290933890, 2191955339, 564929546, 152112058]
    o.sgrnd(1982)          # vpi.in default seed
    assert [o.mt_raw() for _ in range(5)] == [1660960932, 2371444531, 3587904794, 3514745772, 2424620290]
    # numpy's MT19937 core on the SAME state words (69069-LCG seeding done here): 5000 outputs, 8 refills
    key = np.zeros(624, dtype=np.uint32)
    key[0] = 1982
    for i in range(1, 624):
        key[i] = (69069 * int(key[i - 1])) & 0xFFFFFFFF
    bg = np.random.MT19937()
    bg.state = {"bit_generator": "MT19937", "state": {"key": key, "pos": 624}}
    ref = bg.random_raw(5000)
    o.sgrnd(1982)
    got = np.array([o.mt_raw() for _ in range(5000)], dtype=np.uint64)
    assert np.array_equal(got, ref)
    # grnd() = y / (2^32 - 1), [0,1] inclusive (random_mod.f90:108-112)
    o.sgrnd(1982)
    u = np.array([o.grnd() for _ in range(100)])
    assert np.array_equal(u, ref[:100].astype(np.float64) / 4294967295.0)


def test_rangauss_is_polar_method_on_the_same_stream():
    o = Oracle(oracle_cfg(CW))
    o.sgrnd(77)
    us = [o.grnd() for _ in range(400)]
    o.sgrnd(77)
    i, out = 0, []
    for _ in range(50):
        while True:
            u1, u2 = 2 * us[i] - 1, 2 * us[i + 1] - 1
            i += 2
            w = u1 * u1 + u2 * u2
            if w <= 1.0:
                break
        out.append(u1 * np.sqrt(-2 * np.log(w) / w))
    got = [o.rangauss() for _ in range(50)]
    assert np.allclose(got, out, rtol=1e-15, atol=0)


def test_aziz_hfdb_and_mcmillan_anchors():
    # SURVEY Appendix E; V(r_m) = -epsilon = -10.948 K / 1.85505 K
    assert potential(1.0) == pytest.approx(5.406438354381778, rel=1e-14)
    assert potential(0.9) == pytest.approx(53.590308124941856, rel=1e-14)
    assert potential(2.963 / 2.556) == pytest.approx(-10.948 / 1.85505153154686, rel=2e-6)
    assert potential(2.0) == pytest.approx(-0.3420735711435952, rel=1e-14)
    r = np.linspace(1.05, 1.3, 200)
    v = np.array([potential(x) for x in r])
    assert abs(r[np.argmin(v)] - 2.963 / 2.556) < 2e-3          # minimum at r_m
    assert [logpsi(i, 1.2, 1.5) for i in range(3)] == pytest.approx([-0.16384, 0.546133333333, -2.184533333333], rel=1e-11)
    h = 1e-5      # derivatives are consistent
    assert logpsi(1, 1.2, 1.5) == pytest.approx((logpsi(0, 1.2, 1.5 + h) - logpsi(0, 1.2, 1.5 - h)) / (2 * h), rel=1e-8)


def test_table_lookup_has_the_reference_one_step_shift():
    o = Oracle(oracle_cfg(C2))
    assert o.Lbox[0] == pytest.approx(5.597091024810214, rel=1e-15)
    assert o.rcut == pytest.approx(2.798545512405107, rel=1e-15)
    assert o.dr == pytest.approx(2.7988253949446013e-4, rel=1e-15)
    o.fill_tables()
    W, V = o.get_tables()
    assert np.isnan(V[1]) and np.isneginf(W[1]) and V[0] == V[2] and V[-1] == V[-2]
    # Interpolate(0, VTable, r) evaluates V(r - dr), not V(r)  (Q1)
    for r in (1.0, 1.5, 2.0):
        got = interpolate(0, V, o.dr, r)
        assert got == pytest.approx(potential(r - o.dr), rel=1e-5)          # linear-interpolation error only
        assert abs(got - potential(r)) > 20 * abs(got - potential(r - o.dr))  # ... and clearly not V(r)
    assert interpolate(0, V, o.dr, 1.0) == pytest.approx(5.467276343025026, rel=1e-13)
    assert interpolate(1, V, o.dr, 1.0) == pytest.approx(-217.80531233252734, rel=1e-10)


def test_trap_local_energy_is_zero_variance():
    o = Oracle(oracle_cfg(C1))
    o.set_tables(np.zeros(C1["Nmax"] + 2), np.zeros(C1["Nmax"] + 2))
    rng = np.random.default_rng(0)
    for _ in range(5):
        R = rng.normal(0, 1.3, size=(C1["Np"], 3))
        E, K, V = o.local_energy(R)
        assert E == pytest.approx(3 * C1["Np"] / 2.0, abs=1e-12)      # dim*N/(2 a^2)
        assert V == pytest.approx(0.5 * (R ** 2).sum(), rel=1e-13)


def test_update_action_equals_difference_of_slice_potentials():
    """even interior slice: DeltaS = (2 dt/3) [V(R') - V(R)], end slice: dt/3 DeltaV - Delta ln Psi"""
    cfg = C2
    o = Oracle(oracle_cfg(cfg))
    o.fill_tables()
    W, _ = o.get_tables()
    rng = np.random.default_rng(4)
    P = synthetic_path(cfg, rng)
    L = o.Lbox[0]

    def lnpsi(R):
        s = 0.0
        for i in range(len(R) - 1):
            d = R[i] - R[i + 1:]
            d = d - L * np.round(d / L)
            r = np.sqrt((d ** 2).sum(1))
            s += sum(interpolate(0, W, o.dr, x) for x in r[r <= o.rcut])
        return s

    for ib in (2, 14, 0, 30):
        R = P[ib].copy()
        ip = 7
        xold = R[ip - 1].copy()
        xnew = (xold + rng.normal(0, 0.1, 3) + L / 2) % L - L / 2
        Rn = R.copy()
        Rn[ip - 1] = xnew
        dV = o.potential_energy(Rn, False)[0] - o.potential_energy(R, False)[0]
        dS = o.update_action(ip, ib, xnew, xold, R=Rn)
        if ib in (0, 30):
            want = cfg["dt"] * dV / 3 - (lnpsi(Rn) - lnpsi(R))
        else:
            want = 2 * cfg["dt"] * dV / 3
        assert dS == pytest.approx(want, rel=1e-9, abs=1e-11)
    # Chin weights (global_mod.f90:19-72)
    assert o.green_function(0, 1, 0.1, 2.0, 3.0) == pytest.approx(4 * 0.1 * (2.0 + 0.01 * 3.0 / 6) / 3)
    assert o.green_function(1, 1, 0.1, 2.0, 3.0) == pytest.approx(4 * (2.0 + 0.01 * 3.0 / 2) / 3)
    assert o.green_function(0, 0, 0.1, 2.0, 3.0) == pytest.approx(0.1 * 2.0 / 3)


def test_moves_restore_path_on_reject_and_translate_rigidly():
    cfg = CW
    o = Oracle(oracle_cfg(cfg))
    o.fill_tables()
    rng = np.random.default_rng(8)
    P = synthetic_path(cfg, rng)
    xe = np.stack([P[cfg["Nb"], -1]] * 2)
    o.set_state(P, xe, 0, 0)
    o.sgrnd(5)
    L = o.Lbox[0]
    seen = set()
    for it in range(200):
        name = ["translate_chain", "staging", "move_head", "move_tail", "bisection", "move_head_bisection",
                "move_tail_bisection"][it % 7]
        before = o.get_state()[0]
        acc, _ = o.move(name, 1 + it % cfg["Np"])
        after = o.get_state()[0]
        seen.add((name, acc))
        if not acc:
            assert np.array_equal(before, after)
        else:
            ch = np.abs(after - before).max(axis=(0, 2)) > 0
            assert ch.sum() == 1 and ch[it % cfg["Np"]]
            if name == "translate_chain":
                d = (after - before)[:, it % cfg["Np"]]
                d = (d + L / 2) % L - L / 2
                assert np.allclose(d, d[0], atol=1e-13)
    assert ("bisection", 1) in seen and ("translate_chain", 1) in seen


def test_histograms_and_normalisation():
    cfg = C2
    o = Oracle(oracle_cfg(cfg))
    rng = np.random.default_rng(1)
    R = rng.uniform(-o.Lbox[0] / 2, o.Lbox[0] / 2, size=(cfg["Np"], 3))
    gr = o.pair_correlation(R)
    assert gr.sum() % 2 == 0 and 0 < gr.sum() <= cfg["Np"] * (cfg["Np"] - 1)
    # an ideal gas has g(r) ~ 1: average many configurations
    acc = np.zeros(cfg["Nbin"])
    n = 300
    for _ in range(n):
        o.pair_correlation(rng.uniform(-o.Lbox[0] / 2, o.Lbox[0] / 2, size=(cfg["Np"], 3)), acc)
    g = o.normalize_gr(n, acc.copy())
    assert abs(g[30:].mean() - (cfg["Np"] - 1) / cfg["Np"]) < 0.02
    Sk = o.structure_factor(R)
    assert Sk.shape == (cfg["Nk"], 3) and np.all(Sk >= 0)
    assert var(4, 2.0, 5.0) == pytest.approx(0.5)


def test_golden_regression_vectors():
    """Outputs of THIS oracle frozen by tests/golden/make_golden.py (regression guard; not reference-pinned)."""
    g = json.load(open(GOLD))
    cfg = g["cfg"]
    o = Oracle(oracle_cfg(cfg))
    o.fill_tables()
    P = np.array(g["path"])
    xe = np.array(g["xend"])
    o.set_state(P, xe, 0, 0)
    o.sgrnd(g["seed"])
    for ua in g["update_action"]:
        got = o.update_action(ua["ip"], ua["ib"], ua["xnew"], ua["xold"])
        assert got == pytest.approx(ua["dS"], rel=1e-12, abs=1e-14)
    b, gr, Sk, nr = o.run_block(g["nstep"])
    for k, v in g["block_int"].items():
        assert int(b[k]) == v, k
    assert list(b["bead_updates"]) == g["bead_updates"]
    assert b["sumE"] == pytest.approx(g["sumE"], rel=1e-11)
    assert b["sumEt"] == pytest.approx(g["sumEt"], rel=1e-11)
    assert gr.tolist() == g["gr"]
