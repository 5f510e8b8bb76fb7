"""Cross-chain statistics: with thousands of independent
Markov chains per block the reference's single-chain `Var` (sample_mod.f90:921-932,
sqrt((<x^2>-<x>^2)/n) over correlated steps) is replaced by estimators that use
the independence of the chains and the blocking of the time series."""
from __future__ import annotations

import numpy as np


def chain_means(sim, keys=("sumE", "sumEt", "sumV"), chains=None):
    """Per-chain block means of the last block: {key: array[n]} divided by each chain's diagonal count."""
    chains = range(sim.n_chains) if chains is None else chains
    out = {k: [] for k in keys}
    nd = []
    for c in chains:
        b = sim.get_block(chain=c)[0]
        n = max(int(b["idiag_block"]), 0)
        nd.append(n)
        for k in keys:
            out[k].append(b[k] / n if n else np.nan)
    out = {k: np.asarray(v) for k, v in out.items()}
    out["idiag_block"] = np.asarray(nd)
    return out


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
