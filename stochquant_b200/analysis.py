"""Ensemble analysis over the C-ABI (the part with a well-defined meaning): error bars for the
running observables a context hands back.

The reference plots log|<x(t) x(t_mid)> - <x(t)><x(t_mid)>| of ONE chain and reads the gap off the slope
(taumain.py:31-41,137); it has no error analysis.  With `nchains` independent chains per context (and more per
GPU) the natural estimators are over chains and, within a chain, over bins of tau-steps.  Host-side numpy on the
few numbers `sq_measure` / `sq_measure_chains` return -- not part of the hot path, no device code.

    series = collect_series(ctx, dtau, loops, nframes)       # [nframes][nchains] of <phi^2> per frame
    mean, err = binned_error(series.mean(axis=1), nbins=20)
    mean, err = jackknife(lambda s: s.mean(), per_chain_means)
"""
from __future__ import annotations

import numpy as np


def binning(x, nbins: int):
    """Means of `nbins` consecutive, equally long bins of a series (a tail that does not fill a bin is dropped)."""
    x = np.asarray(x, dtype=np.float64)
    if nbins < 1 or x.shape[0] < nbins:
        raise ValueError("need at least one sample per bin")
    n = (x.shape[0] // nbins) * nbins
    return x[:n].reshape((nbins, n // nbins) + x.shape[1:]).mean(axis=1)


def binned_error(x, nbins: int = 20):
    """(mean, standard error) of an autocorrelated series from the scatter of its bin means."""
    b = binning(x, nbins)
    return b.mean(axis=0), b.std(axis=0, ddof=1) / np.sqrt(b.shape[0])


def jackknife(estimator, samples):
    """Delete-one jackknife of `estimator(samples_without_i)` over the first axis: (bias-corrected value, error).
    For derived quantities (connected correlator, effective mass) whose error does not follow from a plain mean."""
    s = np.asarray(samples, dtype=np.float64)
    n = s.shape[0]
    if n < 2:
        raise ValueError("jackknife needs at least two samples")
    full = np.asarray(estimator(s), dtype=np.float64)
    loo = np.array([estimator(np.delete(s, i, axis=0)) for i in range(n)], dtype=np.float64)
    mean_loo = loo.mean(axis=0)
    err = np.sqrt((n - 1) / n * ((loo - mean_loo) ** 2).sum(axis=0))
    return n * full - (n - 1) * mean_loo, err


def tau_int(x, window_c: float = 6.0):
    """Integrated autocorrelation time of a series (in samples) with Sokal's automatic window W >= c tau_int."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    d = x - x.mean()
    var = float(d @ d) / n
    if var == 0.0 or n < 4:
        return 0.5
    t = 0.5
    for w in range(1, n // 2):
        t += float(d[:-w] @ d[w:]) / (n - w) / var
        if w >= window_c * t:
            break
    return max(t, 0.5)


def connected_correlator(slice_x, slice_xx0, tmid: int):
    """C(t) = <Phi(t) Phi(t_mid)> - <Phi(t)><Phi(t_mid)> from the running means `sq_measure` returns (the quantity the
    reference plots, tauhost.c:519-521)."""
    sx, sxx = np.asarray(slice_x, dtype=np.float64), np.asarray(slice_xx0, dtype=np.float64)
    return sxx - sx * sx[..., tmid:tmid + 1]


def effective_mass(corr, tmid: int):
    """log(C(t)/C(t+1)) on the far side of t_mid (lattice units); NaN where the ratio is not positive."""
    c = np.asarray(corr, dtype=np.float64)[..., tmid:]
    with np.errstate(divide="ignore", invalid="ignore"):
        r = c[..., :-1] / c[..., 1:]
        return np.where(r > 0, np.log(r), np.nan)


def collect_series(ctx, dtau: float, loops: int, nframes: int):
    """Advance `ctx` by `nframes` frames of `loops` tau-steps; returns ([nframes][nchains] <phi>, same for <phi^2>) as
    sq_measure_chains reports them after each frame (the last step's global means per chain)."""
    m1, m2 = [], []
    for _ in range(nframes):
        if not ctx.step(dtau, loops):
            raise RuntimeError("frame rejected")
        a, b, _ = ctx.measure_chains()
        m1.append(np.array(a, dtype=np.float64))
        m2.append(np.array(b, dtype=np.float64))
    return np.array(m1), np.array(m2)
