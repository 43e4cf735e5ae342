"""The statistical yardstick of the GPU posterior tests, checked on series with known answers (CPU)."""
import numpy as np

from stat_helpers import autocorr_time, compare_marginal, n_eff


def ar1(rng, n, rho, mu=0.0):
    e = rng.standard_normal(n) * np.sqrt(1 - rho * rho)
    x = np.empty(n)
    x[0] = rng.standard_normal()
    for i in range(1, n):
        x[i] = rho * x[i - 1] + e[i]
    return x + mu


def test_autocorrelation_time_of_ar1():
    rng = np.random.default_rng(1)
    for rho in (0.0, 0.5, 0.9):
        tau = autocorr_time(ar1(rng, 200000, rho))
        assert abs(tau - (1 + rho) / (1 - rho)) < 0.15 * (1 + rho) / (1 - rho)
    assert abs(n_eff([ar1(rng, 50000, 0.8) for _ in range(4)]) - 4 * 50000 / 9.0) < 0.2 * 4 * 50000 / 9.0


def test_same_distribution_passes_and_a_shift_fails():
    rng = np.random.default_rng(2)
    a = [ar1(rng, 40000, 0.7) for _ in range(4)]
    b = [ar1(rng, 30000, 0.3) for _ in range(3)]
    assert compare_marginal(a, b)["ok"]
    c = [ar1(rng, 30000, 0.3, mu=0.12) for _ in range(3)]       # a shift of 0.12 sigma at n_eff ~ 28 000
    assert not compare_marginal(a, c)["ok"]
    d = [1.15 * ar1(rng, 30000, 0.3) for _ in range(3)]          # 15 % wider
    assert not compare_marginal(a, d)["ok"]
