"""Host-side ensemble analysis (stochquant_b200/analysis.py): estimators checked on synthetic series
with known answers.  CPU only."""
import numpy as np
import pytest

from stochquant_b200 import analysis as an


def ar1(n, rho, rng):
    x = np.empty(n)
    x[0] = rng.normal()
    e = rng.normal(size=n) * np.sqrt(1 - rho * rho)
    for i in range(1, n):
        x[i] = rho * x[i - 1] + e[i]
    return x


def test_binning_shapes_and_mean():
    x = np.arange(103, dtype=np.float64)
    b = an.binning(x, 10)
    assert b.shape == (10,) and np.isclose(b.mean(), x[:100].mean())
    with pytest.raises(ValueError):
        an.binning(x[:3], 10)


def test_binned_error_sees_autocorrelation():
    rng = np.random.default_rng(1)
    x = ar1(200000, 0.9, rng)
    _, naive = x.mean(), x.std(ddof=1) / np.sqrt(x.size)
    _, err = an.binned_error(x, 50)
    want = naive * np.sqrt((1 + 0.9) / (1 - 0.9))  # sqrt(2 tau_int)
    assert 0.75 * want < err < 1.3 * want
    assert abs(an.tau_int(x) - 0.5 * (1 + 0.9) / (1 - 0.9)) < 1.5


def test_jackknife_equals_standard_error_for_the_mean_and_handles_ratios():
    rng = np.random.default_rng(2)
    s = rng.normal(3.0, 0.5, size=400)
    v, e = an.jackknife(lambda a: a.mean(), s)
    assert np.isclose(v, s.mean()) and np.isclose(e, s.std(ddof=1) / np.sqrt(s.size))
    pairs = np.stack([rng.normal(2.0, 0.1, 300), rng.normal(4.0, 0.1, 300)], axis=1)
    v, e = an.jackknife(lambda a: a[:, 0].mean() / a[:, 1].mean(), pairs)
    assert abs(v - 0.5) < 4 * e and 0.001 < e < 0.01


def test_correlator_and_effective_mass():
    t = np.arange(32)
    tm, m = 16, 0.37
    mean = 0.2 * np.ones(32)
    xx0 = np.exp(-m * np.abs(t - tm)) + mean * mean[tm]
    c = an.connected_correlator(mean, xx0, tm)
    assert np.allclose(c, np.exp(-m * np.abs(t - tm)))
    assert np.allclose(an.effective_mass(c, tm), m)
