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


# ---- B-1 as run: the wind a switched-off particle is tested against ---------------------------

@pytest.mark.parametrize("shape", ["strengthening", "pulse"])
def test_off_particles_test_the_wind_at_their_own_clock(shape):
    """mapping_2D.jl:132,172-176: an off particle is tested against winds(x, y, integ.t + DT).  As the
    reference runs (`on` frozen at seed, SURVEY B-1) a particle seeded off never integrates, so its
    clock stays at 0 and that time is DT on every step.  The closure-mode oracle calls the closure
    there literally; the staged-level paths (oracle and the device code) must take the same
    decisions from the level they keep from the first step — and would not with the model-clock
    level t + DT, which both winds make visible: one that is calm at DT and strengthens later (never
    reseeded, the clock level would), and a pulse that peaks at DT (reseeded on every step, the clock
    level would stop after the first)."""
    g = cartesian_grid(12, 9)
    DT, nsteps = 900.0, 7
    P = default_params(DT=DT, wind_min_squared=2.0)

    def amp(t):
        if shape == "strengthening":
            return 0.9 * (0.5 + t / DT)          # |u| = 1.35 at DT: speed^2 = 1.9125 < 2
        return 0.3 + 1.5 * math.exp(-((t - DT) / DT) ** 2)  # 0.85 at 0, 1.8 at DT, back to 0.85 at 2 DT

    def wind(x, y, t):
        return (amp(t) if x < 12000.0 else 0.2), 0.3

    def arrays(t):
        u = np.where(g["x"] < 12000.0, amp(t), 0.2)
        return u, np.full_like(u, 0.3)

    exact = make_oracle(g, P)
    exact.set_wind_closure(wind, g["x"], g["y"])
    staged = make_oracle(g, P)
    dev = HostShim(g, P)
    for m in (exact, staged, dev):
        m.seed(*arrays(0.0))
    off = (staged.particles()["flags"] & 9) == 8  # iterated and seeded off: all of them
    assert off.sum() == 10 * 7
    windy = int((off & (g["x"] < 12000.0)).sum())
    lag_rule, clock_rule = [], []
    t = 0.0
    for k in range(nsteps):
        for m in (exact, staged, dev):
            m.step(t, DT, *arrays(t), *arrays(t + DT))
        u1, v1 = arrays(t + DT)
        clock_rule.append(int(((u1 * u1 + v1 * v1 >= 2.0) & off).sum()))
        t += DT
        ce, cs, cd = exact.counters(), staged.counters(), dev.counters()
        assert ce["n_reseed_advance"] == cs["n_reseed_advance"] == cd["n_reseed_advance"]
        assert ce["n_deposited"] == cs["n_deposited"] == cd["n_deposited"]
        for nm in ("n_remesh_A", "n_remesh_B", "n_remesh_D"):
            assert ce[nm] == cs[nm] == cd[nm]
        assert bits_equal(staged.state(), dev.state())
        # nobody integrates (`on` is frozen off), so State holds re-deposited wind seas only and the closure
        # run has nothing to differ by
        assert ce["n_integrated"] == 0 and bits_equal(exact.state(), staged.state())
        lag_rule.append(cs["n_reseed_advance"])
    if shape == "strengthening":
        assert lag_rule == [0] * nsteps and clock_rule[0] == 0 and clock_rule[-1] == windy
    else:
        assert lag_rule == [windy] * nsteps and clock_rule == [windy] + [0] * (nsteps - 1)
    compare_models(staged, dev)
    # as intended (on_persist): `on` follows the remesh, a particle that is switched on integrates from then on
    P2 = default_params(DT=DT, wind_min_squared=2.0, on_persist=True)
    a, b = make_oracle(g, P2), HostShim(g, P2)
    for m in (a, b):
        m.seed(*arrays(0.0))
    t = 0.0
    for k in range(nsteps):
        for m in (a, b):
            m.step(t, DT, *arrays(t), *arrays(t + DT))
        t += DT
    compare_models(a, b)


@pytest.mark.parametrize("seed", range(12))
def test_wind_mesh_sampler_on_random_meshes(seed):
    """random knot vectors (uniform or not, 2 to 19 knots per axis), random fields, query points inside, on the knots and
    up to three periods outside, times inside and outside: device header = oracle bit for bit, and the oracle within 1e-11
    of scipy's RegularGridInterpolator inside the mesh (300 + 200 seeds were run once)"""
    from scipy.interpolate import RegularGridInterpolator
    rng = np.random.default_rng(1000 + seed)
    nx, ny, nt = int(rng.integers(2, 20)), int(rng.integers(2, 15)), int(rng.integers(2, 8))

    def knots(n):
        if rng.random() < 0.5:
            return np.linspace(rng.uniform(-1e5, 1e5), rng.uniform(2e5, 9e5), n)
        return np.cumsum(rng.uniform(0.1, 3.0, n)) * rng.uniform(10, 1e4) + rng.uniform(-1e5, 1e5)

    xw, yw, tw = knots(nx), knots(ny), knots(nt)
    U, V = rng.normal(5, 6, (nt, ny, nx)), rng.normal(-2, 5, (nt, ny, nx))
    n = int(rng.integers(1, 700))
    Lx, Ly, Lt = xw[-1] - xw[0], yw[-1] - yw[0], tw[-1] - tw[0]
    x = np.concatenate([rng.uniform(xw[0] - 3 * Lx, xw[-1] + 3 * Lx, n), xw, [xw[0], xw[-1]]])
    y = np.concatenate([rng.uniform(yw[0] - 2 * Ly, yw[-1] + 2 * Ly, n), np.resize(yw, xw.size), [yw[-1], yw[0]]])
    for t in list(rng.uniform(tw[0] - 2 * Lt, tw[-1] + 2 * Lt, 4)) + [tw[0], tw[-1], tw[len(tw) // 2]]:
        uo, vo = oracle.wind_mesh_sample(xw, yw, tw, U, V, x, y, float(t))
        us, vs = shim_sample(xw, yw, tw, U, V, x, y, float(t))
        assert bits_equal(uo, us) and bits_equal(vo, vs)
    xi, yi = rng.uniform(xw[0], xw[-1], 300), rng.uniform(yw[0], yw[-1], 300)
    for t in rng.uniform(tw[0], tw[-1], 3):
        ref = RegularGridInterpolator((tw, yw, xw), U)(np.stack([np.full_like(xi, t), yi, xi], axis=1))
        got, _ = oracle.wind_mesh_sample(xw, yw, tw, U, V, xi, yi, float(t))
        assert np.max(np.abs(got - ref)) < 1e-11
