"""Wind ingestion on the CPU: intermediate wind levels and the wind mesh sampler.

The reference calls the wind closures at every Runge-Kutta stage time
(particle_waves_v5.jl:489-495); the product integrates against levels staged per model step.
Checked here:
  * the device code (host build of physics.h) and the oracle agree bit for bit with 1-3
    intermediate levels, single domain and strips;
  * the staged-level runs converge to the closure-exact run (oracle_set_wind_closure: the
    reference's semantics) as levels are added, on the time-varying wind of
    tests/T03_PIC_tripolar_aqua.jl:67-68;
  * the wind mesh sampler (wind_mesh.h) equals the oracle's restatement of
    Interpolations.LinearInterpolation(..., extrapolation_bc=Periodic()) bit for bit and an
    independent scipy interpolator to rounding.
"""
import math

import numpy as np
import pytest

import oracle
from common import HostShim, bits_equal, cartesian_grid, compare_models, default_params, make_oracle, shim_lib, _p


def aqua_wind(x, y, t):
    """tests/T03_PIC_tripolar_aqua.jl:67-68: u = 15, v = -10 cos(5 t / (3600 * 2 pi)), plus a slow
    zonal modulation so both components change in time"""
    w = 5.0 / (3600.0 * 2.0 * math.pi)
    return 15.0 + 2.0 * math.sin(0.7 * w * t), -10.0 * math.cos(w * t)


def wind_arrays(g, t):
    u, v = aqua_wind(0.0, 0.0, t)
    return np.full((g["Ny"], g["Nx"]), u), np.full((g["Ny"], g["Nx"]), v)


def mid_times(t, DT, n_mid):
    # the formula picles_step_wind_mesh uses: t + DT*k/(n_mid+1)
    return [t + DT * float(k) / float(n_mid + 1) for k in range(1, n_mid + 1)]


def run_levels(model, g, DT, nsteps, n_mid):
    u0, v0 = wind_arrays(g, 0.0)
    model.seed(u0, v0)
    t = 0.0
    for _ in range(nsteps):
        if n_mid:
            lv = [wind_arrays(g, tm) for tm in mid_times(t, DT, n_mid)]
            model.set_wind_midlevels([a for a, _ in lv], [b for _, b in lv])
        ut, vt = wind_arrays(g, t)
        ut1, vt1 = wind_arrays(g, t + DT)
        model.step(t, DT, ut, vt, ut1, vt1)
        t += DT


@pytest.mark.parametrize("n_mid", [1, 2, 3])
@pytest.mark.parametrize("nstrips", [1, 2])
def test_midlevels_device_code_matches_oracle(n_mid, nstrips):
    g = cartesian_grid(14, 12)
    P = default_params(DT=1200.0)
    ref = make_oracle(g, P)
    dut = HostShim(g, P, nstrips=nstrips, halo=2)
    run_levels(ref, g, 1200.0, 4, n_mid)
    run_levels(dut, g, 1200.0, 4, n_mid)
    compare_models(ref, dut)


def test_midlevels_are_consumed_by_one_step():
    """levels set before step k must not leak into step k+1"""
    g = cartesian_grid(9, 9)
    P = default_params(DT=1200.0)
    a, b = make_oracle(g, P), HostShim(g, P)
    for m in (a, b):
        u0, v0 = wind_arrays(g, 0.0)
        m.seed(u0, v0)
        lv = [wind_arrays(g, tm) for tm in mid_times(0.0, 1200.0, 2)]
        m.set_wind_midlevels([x for x, _ in lv], [y for _, y in lv])
        m.step(0.0, 1200.0, *wind_arrays(g, 0.0), *wind_arrays(g, 1200.0))
        m.step(1200.0, 1200.0, *wind_arrays(g, 1200.0), *wind_arrays(g, 2400.0))
    compare_models(a, b)
    # the second step alone, from the same state, with no levels set: identical
    c = make_oracle(g, P)
    u0, v0 = wind_arrays(g, 0.0)
    c.seed(u0, v0)
    lv = [wind_arrays(g, tm) for tm in mid_times(0.0, 1200.0, 2)]
    c.set_wind_midlevels([x for x, _ in lv], [y for _, y in lv])
    c.step(0.0, 1200.0, *wind_arrays(g, 0.0), *wind_arrays(g, 1200.0))
    c.set_wind_midlevels([], [])
    c.step(1200.0, 1200.0, *wind_arrays(g, 1200.0), *wind_arrays(g, 2400.0))
    assert bits_equal(a.state(), c.state())


def test_polynomial_winds_are_reproduced_exactly():
    """a wind that is a cubic in time is the interpolant itself with 2 intermediate levels:
    the staged run equals the closure-exact run to rounding"""
    g = cartesian_grid(7, 7)
    P = default_params(DT=900.0)
    DT = 900.0

    def cubic(x, y, t):
        s = t / 3600.0
        return 9.0 + 1.5 * s - 0.8 * s * s + 0.3 * s ** 3, 7.0 - 2.0 * s + 0.5 * s * s

    def arrays(t):
        u, v = cubic(0, 0, t)
        return np.full((7, 7), u), np.full((7, 7), v)

    exact = make_oracle(g, P)
    exact.set_wind_closure(cubic, g["x"], g["y"])
    staged = make_oracle(g, P)
    for m in (exact, staged):
        m.seed(*arrays(0.0))
    t = 0.0
    for _ in range(3):
        lv = [arrays(tm) for tm in mid_times(t, DT, 2)]
        staged.set_wind_midlevels([a for a, _ in lv], [b for _, b in lv])
        for m in (exact, staged):
            m.step(t, DT, *arrays(t), *arrays(t + DT))
        t += DT
    Se, Ss = exact.state(), staged.state()
    assert np.max(np.abs(Se - Ss) / np.maximum(np.abs(Se), 1e-300)) < 1e-9


def test_staged_levels_converge_to_the_closure():
    """relative error of lne and c̄ (the north-star's tolerance: 1e-6) against the closure-exact
    run after 6 steps of DT = 20 min: falls with every level added, and is below 1e-6 with 3"""
    g = cartesian_grid(9, 9)
    DT, nsteps = 1200.0, 6
    P = default_params(DT=DT)
    exact = make_oracle(g, P)
    exact.set_wind_closure(aqua_wind, g["x"], g["y"])
    run_levels(exact, g, DT, nsteps, 0)
    pe = exact.particles()
    act = (pe["flags"] & 8) != 0
    errs = []
    for n_mid in (0, 1, 2, 3):
        m = make_oracle(g, P)
        run_levels(m, g, DT, nsteps, n_mid)
        pm = m.particles()
        e_lne = np.max(np.abs(pm["z"][0][act] - pe["z"][0][act]) / np.abs(pe["z"][0][act]))
        cg_e = np.hypot(pe["z"][1][act], pe["z"][2][act])
        e_cg = np.max(np.hypot(pm["z"][1][act] - pe["z"][1][act], pm["z"][2][act] - pe["z"][2][act]) / cg_e)
        errs.append(max(e_lne, e_cg))
    print("relative error vs closure-exact, n_mid = 0..3:", errs)
    assert errs[0] > 1e-5            # two levels: the documented fork is visible on this wind
    assert errs[1] < errs[0] / 20 and errs[2] < errs[1] and errs[3] <= errs[2] * 1.5
    assert errs[3] < 1e-6


# ---- wind mesh ------------------------------------------------------------------------

def synthetic_mesh(seed=3):
    rng = np.random.default_rng(seed)
    xw = np.cumsum(rng.uniform(0.5, 1.5, 13)) * 3000.0 - 2000.0     # non-uniform knots
    yw = np.linspace(-1000.0, 30000.0, 9)
    tw = np.array([0.0, 21600.0, 43200.0, 64800.0, 86400.0])         # 6-hourly, as ERA5 in T03_..._realistic
    U = rng.normal(8.0, 4.0, (tw.size, yw.size, xw.size))
    V = rng.normal(-3.0, 5.0, (tw.size, yw.size, xw.size))
    return xw, yw, tw, U, V


def shim_sample(xw, yw, tw, U, V, x, y, t):
    lib = shim_lib()
    x = np.ascontiguousarray(x, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    u, v = np.empty(x.shape), np.empty(x.shape)
    lib.shim_wind_mesh_sample(xw.size, yw.size, tw.size, _p(xw), _p(yw), _p(tw), _p(np.ascontiguousarray(U)),
                              _p(np.ascontiguousarray(V)), x.size, _p(x), _p(y), float(t), _p(u), _p(v))
    return u, v


def test_wind_mesh_sampler_matches_oracle_bitwise():
    xw, yw, tw, U, V = synthetic_mesh()
    rng = np.random.default_rng(0)
    # points inside, on knots, outside on both sides (periodic wrap), far outside
    x = np.concatenate([rng.uniform(xw[0] - 3 * (xw[-1] - xw[0]), xw[-1] + 3 * (xw[-1] - xw[0]), 4000), xw, [xw[0], xw[-1]]])
    y = np.concatenate([rng.uniform(yw[0] - 2 * (yw[-1] - yw[0]), yw[-1] + 2 * (yw[-1] - yw[0]), 4000),
                        np.resize(yw, xw.size), [yw[-1], yw[0]]])
    for t in (0.0, 100.0, 21600.0, 50000.5, 86400.0, 90000.0, -500.0, 3 * 86400.0 + 17.0):
        uo, vo = oracle.wind_mesh_sample(xw, yw, tw, U, V, x, y, t)
        us, vs = shim_sample(xw, yw, tw, U, V, x, y, t)
        assert bits_equal(uo, us) and bits_equal(vo, vs)
    # a regular node grid: the nodes of a row share y (what the four-nodes-at-once form of the device
    # sampler reuses); rows of 37 nodes, so groups of four also straddle row ends
    gx, gy = np.meshgrid(np.linspace(xw[0] - 5000.0, xw[-1] + 9000.0, 37), np.linspace(yw[0] - 700.0, yw[-1] + 700.0, 11))
    for t in (0.0, 50000.5):
        uo, vo = oracle.wind_mesh_sample(xw, yw, tw, U, V, gx.ravel(), gy.ravel(), t)
        us, vs = shim_sample(xw, yw, tw, U, V, gx.ravel(), gy.ravel(), t)
        assert bits_equal(uo, us) and bits_equal(vo, vs)


def test_wind_mesh_sampler_against_scipy():
    from scipy.interpolate import RegularGridInterpolator
    xw, yw, tw, U, V = synthetic_mesh()
    rng = np.random.default_rng(1)
    x = rng.uniform(xw[0], xw[-1], 2000)
    y = rng.uniform(yw[0], yw[-1], 2000)
    for t in (10.0, 30000.0, 86000.0):
        it = RegularGridInterpolator((tw, yw, xw), U)
        ref = it(np.stack([np.full_like(x, t), y, x], axis=1))
        uo, _ = oracle.wind_mesh_sample(xw, yw, tw, U, V, x, y, t)
        assert np.max(np.abs(uo - ref)) < 1e-11
    # Periodic(): one period away gives the same value
    Lx, Ly, Lt = xw[-1] - xw[0], yw[-1] - yw[0], tw[-1] - tw[0]
    a, _ = oracle.wind_mesh_sample(xw, yw, tw, U, V, x, y, 1234.0)
    b, _ = oracle.wind_mesh_sample(xw, yw, tw, U, V, x + Lx, y - Ly, 1234.0 + Lt)
    assert np.max(np.abs(a - b)) < 1e-9
