"""Scenarios of the one-dimensional model (WaveGrowth1D), shapes taken from the reference's 1-D scripts
(tests/B01_1D_regtest_wave_growth.jl:60-75, tests/T03_PIC_propagation_1d.jl:38-52): name -> (grid, params, wind, DT, steps)."""
import math

import numpy as np

from picles_b200 import FetchRelations as FR
from picles_b200.ParticleMesh import OneDGrid, OneDGridNotes
from picles_b200.ParticleSystems import particle_waves_v5 as PW
from picles_b200.params import make_params


def params_1d(DT=600.0, solver="Tsit5", periodic=False, dt=1e-3, dtmin=1e-4, force_dtmin=True, wind_min_squared=4.0,
              log_energy_maximum=math.log(17), nan_eest_rejects=False, **switches):
    pars, cid, scg = PW.ODEParameters(r_g=0.85)
    ps = PW.particle_equations(None, γ=cid.γ, q=cid.q, **switches)   # one forcing field: the 1-D system
    pars1 = dict(r_g=pars["r_g"], C_α=pars["C_α"], C_e=pars["C_e"])  # default_ODE_parameters of the 1-D scripts
    sets = PW.ODESettings(Parameters=pars1, log_energy_minimum=FR.MinimalWindsea(10, 0, DT)["lne"], saving_step=DT, timestep=DT,
                          total_time=6 * 86400.0, dt=dt, dtmin=dtmin, force_dtmin=force_dtmin, solver=solver,
                          wind_min_squared=wind_min_squared, log_energy_maximum=log_energy_maximum)
    return make_params(sets, ps, FR.MinimalState(2, 0, DT), defaults=None, periodic_boundary=periodic,
                       nan_eest_rejects=nan_eest_rejects)


def grid_1d(xmin, xmax, Nx):
    g = OneDGrid(xmin, xmax, Nx)
    return dict(Nx=g.Nx, xmin=g.xmin, dx=g.dx, x=OneDGridNotes(g).x)


def _steady(U):
    return lambda x, t: np.full(np.shape(x), float(U)) if np.ndim(x) else float(U)


def _ramp(L, U=12.0):
    """calm foot, linear ramp in x, modulated in time: on/off thresholds, reseeds, both remesh branches"""
    x0 = 0.2 * L

    def u(x, t):
        r = np.where(np.asarray(x) < x0, 0.02, (np.asarray(x) - x0) / (L - x0))
        out = U * r * (0.6 + 0.4 * np.sin(2 * np.pi * t / 7200.0))
        return out if np.ndim(x) else float(out)
    return u


SCENARIOS_1D = {
    # B01 case (u10 = 15, DT = 10 min, Nx = 51), grid offset as in T03_PIC_propagation_1d (xmin = 1 km: the weights' frame
    # and the particles' frame differ by design of the reference)
    "steady_nonperiodic": lambda: (grid_1d(1e3, 1500e3, 51), params_1d(600.0), _steady(15.0), 600.0, 8),
    "steady_periodic": lambda: (grid_1d(0.0, 400e3, 41), params_1d(600.0, periodic=True), _steady(10.0), 600.0, 8),
    "negative_periodic": lambda: (grid_1d(0.0, 200e3, 33), params_1d(1200.0, periodic=True), _steady(-12.0), 1200.0, 6),
    "ramp_winds": lambda: (grid_1d(0.0, 600e3, 61), params_1d(1200.0, wind_min_squared=2.0), _ramp(600e3), 1200.0, 8),
    # small cells: a particle crosses many of them per step (reach >> 1: the gather's window and its wrap)
    "fast_periodic": lambda: (grid_1d(0.0, 20e3, 101), params_1d(1200.0, periodic=True), _steady(14.0), 1200.0, 5),
    "fast_nonperiodic": lambda: (grid_1d(0.0, 30e3, 76), params_1d(1200.0), _steady(14.0), 1200.0, 5),
    "dp5": lambda: (grid_1d(1e3, 800e3, 41), params_1d(600.0, solver="DP5", dt=10.0, dtmin=1.0, force_dtmin=False,
                                                       log_energy_maximum=math.log(27)), _steady(15.0), 600.0, 6),
    "propagation_only": lambda: (grid_1d(0.0, 300e3, 31), params_1d(1200.0, input=False, dissipation=False, peak_shift=False),
                                 _steady(10.0), 1200.0, 5),
    "emax_reset": lambda: (grid_1d(0.0, 300e3, 21), params_1d(600.0, log_energy_maximum=-6.0), _steady(15.0), 600.0, 4),
}


def run_pair_1d(a, b, g, wind, DT, steps, compare):
    """seed and step two implementations side by side on the same staged winds, comparing after every step"""
    x = g["x"]
    u0 = np.asarray(wind(x, 0.0), np.float64)
    a.seed(u0)
    b.seed(u0)
    compare(a, b, 0)
    t = 0.0
    for k in range(steps):
        u_t = np.asarray(wind(x, t), np.float64)
        u_t1 = np.asarray(wind(x, t + DT), np.float64)
        a.step(t, DT, u_t, u_t1)
        b.step(t, DT, u_t, u_t1)
        t += DT
        compare(a, b, k + 1)


def bits_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        return bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))
    return bool(np.array_equal(a, b))


def compare_models_1d(a, b, step):
    assert bits_equal(a.state(), b.state()), f"State differs after step {step}"
    pa, pb = a.particles(), b.particles()
    for k in ("z", "t", "dt", "flags", "status"):
        assert bits_equal(pa[k], pb[k]), f"particle {k} differs after step {step}"
    if step > 0:
        ca, cb = a.counters(), b.counters()
        for k in ("n_integrated", "n_substeps", "n_rejects", "n_rhs", "n_reseed_advance", "n_fixups", "n_failed", "n_deposited",
                  "n_remesh_A", "n_remesh_B", "n_remesh_D", "max_attempts"):
            assert ca[k] == cb[k], f"counter {k}: {ca[k]} != {cb[k]} after step {step}"


def _nan_at(i, U):
    def u(x, t):
        a = np.full(np.shape(x), float(U)) if np.ndim(x) else float(U)
        if np.ndim(x):
            a[i] = np.nan
        return a
    return u


# edge cases, checked on the CPU only (host build of the device header against the oracle): not part of the GPU
# parametrisation, whose scenarios were all run on a B200 before they were committed
EDGE_1D = {
    # five nodes, particles crossing several cells: the gather's window covers the whole chain (full scan) and wraps
    "tiny_periodic_fast": lambda: (grid_1d(0.0, 2e3, 5), params_1d(1200.0, periodic=True), _steady(14.0), 1200.0, 4),
    "two_nodes": lambda: (grid_1d(0.0, 50e3, 2), params_1d(600.0, periodic=True), _steady(11.0), 600.0, 4),
    # no wind to speak of: every particle seeded off, nothing ever integrates, remesh keeps them off
    "calm": lambda: (grid_1d(0.0, 300e3, 31), params_1d(600.0), _steady(0.5), 600.0, 3),
    "zero_wind": lambda: (grid_1d(0.0, 300e3, 11), params_1d(600.0, periodic=True), _steady(0.0), 600.0, 2),
    # a NaN in the wind field at one node: its particle is seeded off with a NaN state and stays off; its neighbours
    # read a NaN wind when they come near it
    "nan_wind_node": lambda: (grid_1d(0.0, 300e3, 21), params_1d(600.0), _nan_at(10, 13.0), 600.0, 4),
    # maxiters: the integrators stop, as status codes
    "maxiters": lambda: (grid_1d(0.0, 300e3, 15), _with(params_1d(600.0), maxiters=12), _steady(15.0), 600.0, 3),
    "dtmin_no_force": lambda: (grid_1d(0.0, 300e3, 15), params_1d(600.0, dt=1e-3, dtmin=5.0, force_dtmin=False), _steady(15.0), 600.0, 3),
}


def _with(P, **kw):
    for k, v in kw.items():
        setattr(P, k, v)
    return P
