"""Cross-chain statistics: with thousands of independent
Markov chains per block the reference's single-chain `Var` (sample_mod.f90:921-932,
sqrt((<x^2>-<x>^2)/n) over correlated steps) is replaced by estimators that use
the independence of the chains and the blocking of the time series."""
from __future__ import annotations

import numpy as np


def chain_means(sim, keys=("sumE", "sumEt", "sumV"), chains=None):
    """Per-chain block means of the last block: {key: array[n]} divided by each chain's diagonal count.
    One transfer for all chains (pigs_get_block_chains)."""
    if chains is None:
        b = sim.get_block_chains()[0]
    else:
        chains = list(chains)
        lo, hi = min(chains), max(chains) + 1
        full = sim.get_block_chains(lo, hi - lo)[0]
        idx = np.asarray(chains) - lo
        b = {k: v[idx] for k, v in full.items()}
    nd = np.maximum(b["idiag_block"], 0)
    out = {}
    with np.errstate(all="ignore"):
        for k in keys:
            out[k] = np.where(nd > 0, b[k] / np.maximum(nd, 1), np.nan)
    out["idiag_block"] = nd
    return out


class BlockSeries:
    """Per-chain block series: call add(sim) after every block.  series(key)[block, chain] feeds `blocking` (one chain
    over time) or `mean_and_error` (one block over chains); `ratio` gives ratio-of-sums estimators with a jackknife
    over chains (acceptance ratios, histogram fractions)."""

    def __init__(self, keys=("sumE", "sumK", "sumV", "sumEt", "sumKt", "sumVt", "idiag_block", "ngr")):
        self.keys, self.rows, self.gr, self.nr = tuple(keys), [], [], []

    def add(self, sim):
        b, gr, Sk, nr = sim.get_block_chains()
        self.rows.append({k: np.asarray(b[k], dtype=float) for k in self.keys})
        self.gr.append(gr)
        self.nr.append(nr)

    def series(self, key):
        return np.stack([r[key] for r in self.rows])

    def energy_per_particle(self, Np, key="sumE"):
        """(mean, standard error) of key / (diagonal configurations * Np) from per-chain totals over all blocks"""
        num, den = self.series(key).sum(axis=0), self.series("idiag_block").sum(axis=0)
        return ratio(num / Np, den)


def ratio(num, den):
    """ratio of sums over chains and its jackknife standard error"""
    num, den = np.asarray(num, dtype=float), np.asarray(den, dtype=float)
    n = num.size
    R = num.sum() / den.sum()
    loo = (num.sum() - num) / (den.sum() - den)
    return float(R), float(np.sqrt((n - 1) / n * np.sum((loo - loo.mean()) ** 2)))


def permutation_cycles(sim):
    """Distribution of permutation-cycle lengths over all chains: the reference's Perm_histogram (filled by
    PermutationSampling, sample_mod.f90:530-594, written to fort.99 at vpi.f90:590-592), summed over chains.
    Returns (hist[Np] counts of closed cycles by length, P(l) normalised, mean cycle length, chains that closed one)."""
    Np = sim.Np
    hist = np.zeros(Np, dtype=np.int64)
    contributing = 0
    for c in range(sim.n_chains):
        h = np.asarray(sim.get_perm(c)[2], dtype=np.int64)
        hist += h
        contributing += int(h.sum() > 0)
    tot = hist.sum()
    P = hist / tot if tot else np.zeros(Np)
    mean_len = float((np.arange(1, Np + 1) * hist).sum() / tot) if tot else float("nan")
    return hist, P, mean_len, contributing


def mean_and_error(x, weights=None):
    """Mean over independent chains and its standard error (weights = diagonal counts)."""
    x = np.asarray(x, dtype=float)
    ok = np.isfinite(x)
    x = x[ok]
    if weights is None:
        w = np.ones_like(x)
    else:
        w = np.asarray(weights, dtype=float)[ok]
    if x.size < 2 or w.sum() == 0:
        return (float(x.mean()) if x.size else np.nan), np.nan
    m = np.sum(w * x) / np.sum(w)
    neff = np.sum(w) ** 2 / np.sum(w * w)
    var = np.sum(w * (x - m) ** 2) / np.sum(w)
    return float(m), float(np.sqrt(var / max(neff - 1.0, 1.0)))


def blocking(series, min_blocks=8):
    """Flyvbjerg-Petersen blocking of one time series: [(block_len, std_err), ...]; the plateau is the error."""
    x = np.asarray(series, dtype=float)
    out, L = [], 1
    while x.size >= min_blocks:
        out.append((L, float(x.std(ddof=1) / np.sqrt(x.size))))
        n = x.size // 2
        x = 0.5 * (x[:2 * n:2] + x[1:2 * n:2])
        L *= 2
    return out


def jackknife(samples, f=np.mean):
    """Jackknife estimate and error of f over independent samples (e.g. a ratio of chain sums)."""
    s = np.asarray(samples, dtype=float)
    n = s.shape[0]
    full = f(s)
    loo = np.array([f(np.delete(s, i, axis=0)) for i in range(n)])
    return float(full), float(np.sqrt((n - 1) / n * np.sum((loo - loo.mean()) ** 2)))
