"""Direct ulp test of include/jxpow.h (the x^y shared by the oracle and the device functors in pow_mode = 1) against a
50-digit mpmath reference.  Every bit-exact GPU test compares jx_pow with itself on this one function, so its accuracy
is established here independently: < 1 ulp on the equation-of-state domain  P = C0 (rho*theta)^gamma
(src/kernel/physics/constitutiveLaw.jl:22-24; Julia's own `^` is < 1 ulp), and on a wide sweep of bases / exponents."""
import math

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")


def _ulp_err(x, y, got):
    mp.mp.dps = 50
    exact = mp.power(mp.mpf(float(x)), mp.mpf(float(y)))
    ulp = math.ulp(float(exact))
    return float(abs(mp.mpf(float(got)) - exact) / ulp)


def _worst(oracle_lib, xs, ys):
    worst = 0.0
    for x, y in zip(xs, ys):
        worst = max(worst, _ulp_err(x, y, oracle_lib.jxo_pow(float(x), float(y), 1)))
    return worst


def test_jx_pow_eos_domain_below_one_ulp(oracle_lib):
    rng = np.random.default_rng(3)
    gamma = 1004.0 / 718.0
    # rho*theta of the rising-bubble states (60 .. 400), plus a wider atmospheric range
    xs = np.concatenate([rng.uniform(60.0, 400.0, 6000), np.exp(rng.uniform(np.log(1e-2), np.log(1e4), 3000))])
    assert _worst(oracle_lib, xs, np.full(xs.size, gamma)) < 0.75


def test_jx_pow_wide_sweep_below_one_ulp(oracle_lib):
    rng = np.random.default_rng(4)
    xs = np.exp(rng.uniform(np.log(1e-30), np.log(1e30), 4000))
    ys = rng.uniform(-3.0, 3.0, xs.size)
    assert _worst(oracle_lib, xs, ys) < 1.0
    # exact cases
    for x in (0.5, 1.0, 2.0, 1024.0):
        assert oracle_lib.jxo_pow(x, 1.0, 1) == x
        assert oracle_lib.jxo_pow(x, 0.0, 1) == 1.0
    assert oracle_lib.jxo_pow(4.0, 0.5, 1) == 2.0


def test_jx_pow_agrees_with_libm_within_two_ulp(oracle_lib):
    rng = np.random.default_rng(5)
    xs = rng.uniform(60.0, 400.0, 20000)
    gamma = 1004.0 / 718.0
    a = np.array([oracle_lib.jxo_pow(float(x), gamma, 1) for x in xs])
    b = np.array([oracle_lib.jxo_pow(float(x), gamma, 0) for x in xs])
    assert np.max(np.abs(a - b) / np.spacing(np.abs(b))) <= 2.0
