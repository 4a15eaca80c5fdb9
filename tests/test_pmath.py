"""pmath.h (the deterministic math layer shared by device and oracle) against numpy/libm."""
import numpy as np

import oracle


def ulp_err(a, b):
    return np.max(np.abs(a - b) / np.spacing(np.abs(b)))


RNG = np.random.default_rng(20240611)


def test_exp_log_within_1_ulp():
    x = RNG.uniform(-708, 709, 300000)
    assert ulp_err(oracle.pm("exp", x), np.exp(x)) <= 1.0
    x = np.exp(RNG.uniform(-700, 700, 300000))
    assert ulp_err(oracle.pm("log", x), np.log(x)) <= 1.0
    x = RNG.uniform(0.5, 2.0, 300000)
    assert ulp_err(oracle.pm("log", x), np.log(x)) <= 1.0


def test_tanh_sech_pow_accuracy():
    for lo, hi in ((-25, 25), (-1, 1), (-1e-3, 1e-3)):
        x = RNG.uniform(lo, hi, 200000)
        assert ulp_err(oracle.pm("tanh", x), np.tanh(x)) <= 4.0
    x = RNG.uniform(-50, 50, 200000)
    assert ulp_err(oracle.pm("sech", x), 1 / np.cosh(x)) <= 4.0
    x = np.exp(RNG.uniform(-23, 7, 200000))
    y = RNG.uniform(-2, 2, 200000)
    assert ulp_err(oracle.pm("pow", x, y), np.power(x, y)) <= 48.0   # naive exp(y log x): ~|y log x| ulp


def test_special_values():
    with np.errstate(all="ignore"):
        x = np.array([-800.0, 800.0, np.nan, 0.0, -745.2, 709.8, 709.7, -745.1, np.inf, -np.inf])
        assert np.array_equal(oracle.pm("exp", x), np.exp(x), equal_nan=True)
        x = np.array([0.0, -1.0, np.inf, np.nan, 5e-324, 1e-310, 1.0])
        assert np.array_equal(oracle.pm("log", x), np.log(x), equal_nan=True)
        x = np.array([0.0, 30.0, -30.0, np.nan, np.inf, -np.inf, 1e-320])
        assert np.array_equal(oracle.pm("tanh", x), np.tanh(x), equal_nan=True)
        r = oracle.pm("pow", np.array([0.0, 0.0, -1.0, 2.0, np.nan, 2.0]), np.array([1.0, -1.0, 0.5, 0.0, 1.0, np.nan]))
        assert r[0] == 0.0 and np.isinf(r[1]) and np.isnan(r[2]) and r[3] == 1.0 and np.isnan(r[4]) and np.isnan(r[5])
        s = oracle.pm("sech", np.array([0.0, 400.0, np.nan]))
        assert s[0] == 1.0 and s[1] < 1e-150 and np.isnan(s[2])


def test_eps_matches_spacing():
    x = np.concatenate([[0.0, 1.0, 600.0, 86400.0, 1e-320, 1e300, 2.2250738585072014e-308],
                        RNG.uniform(0, 1e6, 2000), np.exp(RNG.uniform(-740, 700, 2000))])
    assert np.array_equal(oracle.pm("eps", x), np.spacing(x))


def test_pinned_log_qoldinit_constant():
    """physics.h hard-codes PH_LOG_QOLDINIT = pm_log(1e-4)."""
    assert oracle.pm("log", np.array([1e-4]))[0].hex() == "-0x1.26bb1bbb55515p+3"


def test_libm_build_of_the_oracle_agrees_to_1e_9():
    """Cross-check of pmath itself: the same oracle built against the system libm ends the
    example_00_minimal run within 1e-9 relative of the pmath build (no accept/reject flips)."""
    from scenarios import SCENARIOS
    from common import make_oracle
    g, P, wind, DT, n = SCENARIOS["minimal"]()
    res = []
    for variant in ("default", "libm"):
        o = make_oracle(g, P, variant=variant)
        o.seed(*wind(0.0))
        t = 0.0
        for _ in range(n):
            o.step(t, DT, *wind(t), *wind(t + DT))
            t += DT
        res.append((o.state(), o.counters()))
    assert res[0][1]["n_substeps"] == res[1][1]["n_substeps"]
    a, b = res[0][0], res[1][0]
    nz = np.abs(a) > 0
    assert np.max(np.abs(a[nz] - b[nz]) / np.abs(a[nz])) < 1e-9
