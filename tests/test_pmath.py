"""pmath.h (the deterministic math layer shared by device and oracle) against numpy/libm."""
import os

import numpy as np
import pytest

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


def test_trig_for_the_grid_metric_within_1_ulp():
    """pmath_trig.h: sin/cos in radians (Cody-Waite + fdlibm kernels) against libm, sind/cosd/tand
    with exact reduction in degrees against the radian functions at a quarter of the argument
    error; exact values at multiples of 30 / 45 / 90 degrees."""
    for lo, hi in ((-7, 7), (-1e-3, 1e-3), (-400, 400), (-1e5, 1e5)):
        x = RNG.uniform(lo, hi, 200000)
        assert ulp_err(oracle.pm("sin", x), np.sin(x)) <= 1.0
        assert ulp_err(oracle.pm("cos", x), np.cos(x)) <= 1.0
    d = np.array([0.0, 30.0, 90.0, 180.0, 270.0, 360.0, -90.0, -30.0])
    assert np.array_equal(oracle.pm("sind", d), [0.0, 0.5, 1.0, 0.0, -1.0, 0.0, -1.0, -0.5])
    assert np.array_equal(oracle.pm("cosd", np.array([0.0, 60.0, 90.0, 180.0, 270.0, -60.0])), [1.0, 0.5, 0.0, -1.0, 0.0, 0.5])
    assert np.array_equal(oracle.pm("tand", np.array([0.0, 45.0, -45.0, 90.0])), [0.0, 1.0, -1.0, np.inf])
    # tan(radians(lat)) carries the rounding of the degree->radian conversion, amplified near
    # the poles, so the reference here is 200-bit arithmetic
    mpmath = pytest.importorskip("mpmath")
    mpmath.mp.prec = 200
    lat = RNG.uniform(-89.99, 89.99, 2000)
    ref = np.array([float(mpmath.tan(mpmath.mpf(float(v)) * mpmath.pi / 180)) for v in lat])
    assert ulp_err(oracle.pm("tand", lat), ref) <= 2.0
    ref = np.array([float(mpmath.sin(mpmath.mpf(float(v)) * mpmath.pi / 180)) for v in 4 * lat])
    assert ulp_err(oracle.pm("sind", 4 * lat), ref) <= 1.0
    small = RNG.uniform(-1.0, 1.0, 100000)
    assert ulp_err(oracle.pm("tand", small), np.tan(np.deg2rad(small))) <= 4.0


def test_grid_metric_matches_the_reference_formulas():
    """oracle.grid_metric (the arithmetic k_grid_metric runs on the device) against a numpy
    restatement of TripolarGridMOM6.jl:448-459 and spherical_grid_corrections.jl:13."""
    n = 20000
    dx = RNG.uniform(500.0, 120e3, n)
    dy = RNG.uniform(500.0, 120e3, n)
    ang = RNG.uniform(-180.0, 180.0, n)
    lat = RNG.uniform(-89.0, 89.9, n)
    M, pc = oracle.grid_metric(dx, dy, ang, lat)
    ca, sa = np.cos(ang * np.pi / 180), np.sin(ang * np.pi / 180)
    ref = np.stack([ca / dx, sa / dy, -sa / dx, ca / dy])
    assert np.max(np.abs(M - ref) / np.maximum(np.abs(ref), 1e-300)) < 1e-12
    sg = np.sign(lat)
    ref_pc = sg * np.minimum(sg * np.tan(np.deg2rad(lat)), 60.0) / 6.3710e6
    assert np.max(np.abs(pc - ref_pc) / np.abs(ref_pc)) < 1e-12
    # the cap at 60 binds above atan(60) = 89.045 degrees, on both hemispheres
    _, pcs = oracle.grid_metric(np.ones(3), np.ones(3), np.zeros(3), np.array([89.5, -89.5, 0.0]))
    assert pcs[0] == 60.0 / 6.3710e6 and pcs[1] == -60.0 / 6.3710e6 and pcs[2] == 0.0


def test_table_driven_exp_against_exact_arithmetic():
    """pmath's exp reduces by ln2/128 with a 128-entry (hi, lo) table of 2^(j/128) (pmath_exptab.h, generated in 80-digit
    arithmetic): measured against mpmath, not against another libm — 0.503 ulp at the time of writing; tanh and sech, built
    on the same reduction, 2.4 and 1.8 ulp.  Also: the table itself, entry by entry."""
    import mpmath as mp
    import re
    mp.mp.dps = 40
    rng = np.random.default_rng(7)

    def worst(name, xs, f):
        got = oracle.pm(name, xs)
        return max(abs((mp.mpf(float(a)) - f(mp.mpf(float(b)))) / mp.mpf(float(np.spacing(abs(a))))) for a, b in zip(got, xs) if a != 0)

    xs = np.concatenate([rng.uniform(-708, 709, 1500), rng.uniform(-1, 1, 1000), rng.uniform(-1e-3, 1e-3, 500)])
    assert worst("exp", xs, mp.exp) < 0.52
    assert worst("tanh", np.concatenate([rng.uniform(-25, 25, 1000), rng.uniform(-1e-3, 1e-3, 500)]), mp.tanh) < 3.0
    assert worst("sech", rng.uniform(-50, 50, 1000), mp.sech) < 2.5
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "picles_b200", "csrc", "pmath_exptab.h")).read()
    vals = [float.fromhex(v) for v in re.findall(r"-?0x[0-9a-f.]+p[+-]\d+", src)]
    assert len(vals) == 256
    for j in range(128):
        exact = mp.power(2, mp.mpf(j) / 128)
        assert vals[2 * j] == float(exact)
        assert vals[2 * j + 1] == float(exact - mp.mpf(vals[2 * j]))
