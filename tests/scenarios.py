"""Named parity scenarios shared by the CPU (host build of the device code) and GPU tests.

Each scenario returns (grid, params, wind_fn, DT, nsteps) where wind_fn(t) -> (u, v) arrays
of shape (Ny, Nx) (or scalars).  Sizes are chosen so the oracle finishes in seconds.
"""
from __future__ import annotations

import numpy as np

from common import (BND_NONPERIODIC, BND_PERIODIC, cartesian_grid, default_params, tripolar_grid)


def _const(u, v):
    return lambda t: (u, v)


def sc_minimal():
    """C1 example_00_minimal.jl: 51x51, 2 km, u=v=10, DT=10 min, 13 steps."""
    g = cartesian_grid(51, 51)
    return g, default_params(), _const(10.0, 10.0), 600.0, 13


def sc_minimal_dp5():
    """bench06 solver settings: DP5, dt=10, dtmin=1, log_e_max=log(27), seed timescale 30 min."""
    g = cartesian_grid(40, 33)
    P = default_params(solver="DP5", dt=10.0, dtmin=1.0, force_dtmin=False, log_energy_maximum=float(np.log(27)),
                       timestep=1800.0)
    return g, P, _const(10.0, 10.0), 600.0, 8


def sc_periodic_model_flag():
    """T04_2D_reg_test.jl:49,145: non-periodic grid, model periodic_boundary=true: grid-boundary
    nodes join ocean_points as a second class (canonical order: mask 1 first, then mask 3)."""
    g = cartesian_grid(30, 24)
    return g, default_params(periodic_boundary=True), _const(-8.0, 5.0), 600.0, 6


def sc_periodic_grid():
    """fully periodic box, strong wind and fast particles (reach 2 cells): wrap path of the gather."""
    g = cartesian_grid(24, 20, dx=600.0, dy=500.0, bx=BND_PERIODIC, by=BND_PERIODIC)
    return g, default_params(periodic_boundary=True), _const(12.0, -9.0), 900.0, 6


def sc_periodic_x_only():
    g = cartesian_grid(26, 18, dx=1000.0, dy=1500.0, bx=BND_PERIODIC, by=BND_NONPERIODIC)
    return g, default_params(periodic_boundary=False), _const(-11.0, 4.0), 600.0, 6


def sc_land_block():
    """S02_2D_box_mesh_grid_single_steps.jl:82-86 style land rectangle inside the box."""
    ocean = np.ones((40, 48), np.uint8)
    ocean[12:22, 20:34] = 0
    g = cartesian_grid(48, 40, ocean=ocean)
    return g, default_params(), _const(9.0, 6.0), 600.0, 8


def _growing(g, U10, V10):
    """T04_2D_growing_decaying_winds.jl:126-132: 0.1 m/s left of x0, linear ramp to the right,
    modulated in time so both staged wind levels differ."""
    x = g["x"]
    Lx = x.max()
    x0 = 50.0 / 260.0 * Lx

    def wind(t):
        ramp = np.where(x < x0, 0.1 / max(abs(U10), 1e-9), (x - x0) / (Lx - x0))
        f = 0.6 + 0.4 * np.sin(2 * np.pi * t / 7200.0)
        u = U10 * ramp * f
        v = V10 * ramp * f + 0.05
        return u, v

    return wind


def sc_growing_winds(on_persist=False):
    """C3-style growing/decaying winds with on/off particles (wind_min_squared=2)."""
    g = cartesian_grid(66, 21, dx=4000.0, dy=4000.0)
    P = default_params(DT=1200.0, wind_min_squared=2.0, on_persist=on_persist)
    return g, P, _growing(g, 10.0, 3.0), 1200.0, 8


def sc_growing_winds_persist():
    return sc_growing_winds(on_persist=True)


def sc_pulse_winds():
    """B-1 as run: the left half is calm at t = 0 (particles seeded off, `on` frozen), the wind there peaks
    at t = DT and decays again.  advance! tests an off particle against winds(x, y, integ.t + DT) with
    integ.t stuck at 0 (mapping_2D.jl:132,172-176): the DT level on every step, so these particles
    re-deposit the same wind sea for ever; the right half integrates normally."""
    g = cartesian_grid(40, 18)
    DT = 900.0
    P = default_params(DT=DT, wind_min_squared=2.0)
    left = g["x"] < 40000.0

    def wind(t):
        a = 0.3 + 1.5 * np.exp(-((t - DT) / DT) ** 2)
        return np.where(left, a, 9.0 + 2.0 * np.sin(t / 2000.0)), np.where(left, 0.3, 4.0)

    return g, P, wind, DT, 6


def sc_tripolar():
    """C4-style synthetic tripolar grid: periodic x, tripolar-north fold, per-node rotated
    kernel and great-circle term, land blobs and masked poles; winds as
    T03_PIC_tripolar_aqua.jl:67-68."""
    Nx, Ny = 48, 36
    ocean = np.ones((Ny, Nx), np.uint8)
    ocean[:2, :] = 0
    ocean[10:16, 8:15] = 0
    ocean[Ny - 3:, 20:27] = 0
    g = tripolar_grid(Nx, Ny, ocean=ocean)
    # coarse cells: make the particles fast enough to cross the seam by shrinking the metric
    g["M"] = g["M"] * 60.0
    P = default_params(DT=1200.0, periodic_boundary=True)

    def wind(t):
        return 15.0, -10.0 * np.cos(5 * t / (3600 * 2 * np.pi))

    return g, P, wind, 1200.0, 6


def sc_tripolar_tall():
    """the tripolar scenario on 72 rows: two strips of 36 rows are tall enough (> 2 x 15) for picles_step_strip's
    overlapped path — boundary zones advanced and exchanged, their reach all-reduced, while the interior integrates"""
    Nx, Ny = 48, 72
    ocean = np.ones((Ny, Nx), np.uint8)
    ocean[:2, :] = 0
    ocean[20:30, 8:15] = 0
    ocean[Ny - 3:, 20:27] = 0
    g = tripolar_grid(Nx, Ny, ocean=ocean)
    g["M"] = g["M"] * 60.0
    P = default_params(DT=1200.0, periodic_boundary=True)

    def wind(t):
        return 15.0, -10.0 * np.cos(5 * t / (3600 * 2 * np.pi))

    return g, P, wind, 1200.0, 5


def sc_tripolar_propagation_only():
    """T03_PIC_tripolar_aqua.jl:149-155: source terms off, default particle, pure advection
    across the fold."""
    Nx, Ny = 40, 30
    g = tripolar_grid(Nx, Ny)
    g["M"] = g["M"] * 80.0
    P = default_params(DT=1800.0, periodic_boundary=True, defaults=[-3.0, 2.0, 6.0, 0.0, 0.0], input=False,
                       dissipation=False, peak_shift=False, direction=False)
    return g, P, _const(5.0, 5.0), 1800.0, 5


def sc_emax_clamp():
    """advance!: lne > log_energy_maximum -> clamp + pending dt reset (mapping_2D.jl:222-233)."""
    g = cartesian_grid(20, 16)
    return g, default_params(log_energy_maximum=float(np.log(2e-3))), _const(14.0, 9.0), 900.0, 6


def sc_maxiters():
    """integrator stops with retcode MaxIters for every particle (status code, not a call
    failure) and stays dead on the following steps; remesh keeps resetting u from the nodes."""
    g = cartesian_grid(18, 14)
    P = default_params()
    P.maxiters = 4
    return g, P, _const(10.0, 10.0), 600.0, 4


def sc_dtmin_no_force():
    """force_dtmin=false with a large dtmin: DtLessThanMin retcode."""
    g = cartesian_grid(16, 12)
    return g, default_params(force_dtmin=False, dt=1e-3, dtmin=10.0), _const(10.0, 10.0), 600.0, 3


def sc_dp5_blowup():
    """DP5 at the reference's tolerances with wind speeds across the 14 m/s band where its second substep (~85 s,
    proposed by the PI controller after the tiny first one) overflows in the stages: EEst = NaN, the rejected
    step leaves dt = NaN and the integrator ends (DtNaN, coded UNSTABLE) — for the columns inside the band only;
    the others integrate normally beside them (tests/test_independent_integrator.py has the single-particle
    view and the wind scan)."""
    g = cartesian_grid(24, 10)
    v = np.broadcast_to(np.linspace(13.6, 14.4, 24), (10, 24)).copy()
    u = np.zeros((10, 24))
    return g, default_params(solver="DP5"), (lambda t: (u, v)), 600.0, 3


def sc_nan_wind():
    """a patch of NaN wind: integration goes NaN, advance! reseeds from the (NaN) wind, the NaN
    charge poisons the four nodes it is deposited on and spreads by one cell per step."""
    g = cartesian_grid(24, 20)

    def wind(t):
        u = np.full((20, 24), 10.0)
        v = np.full((20, 24), 6.0)
        if t > 0:
            u[8:10, 10:13] = np.nan
        return u, v

    return g, default_params(), wind, 600.0, 4


def sc_nan_defaults():
    """ParticleDefaults with a NaN energy: the integrator gives up (NaN dt), advance! takes the
    NaN fix-up (mapping_2D.jl:196-209) and the NaN charge poisons the nodes it lands on."""
    g = cartesian_grid(14, 10)
    return g, default_params(defaults=[float("nan"), 0.5, 0.4, 0.0, 0.0]), _const(10.0, 6.0), 600.0, 3


def sc_inf_defaults():
    """ParticleDefaults with an infinite energy: the Inf fix-up (mapping_2D.jl:211-220)."""
    g = cartesian_grid(14, 10)
    return g, default_params(defaults=[float("inf"), 0.5, 0.4, 0.0, 0.0]), _const(10.0, 6.0), 600.0, 3


def sc_all_land():
    """no ocean at all: nothing is iterated, State stays zero."""
    g = cartesian_grid(12, 9, ocean=np.zeros((9, 12), np.uint8))
    return g, default_params(), _const(10.0, 10.0), 600.0, 2


def sc_calm():
    """wind below sqrt(2) everywhere: minimal particles, seeded off, never switched on
    (remesh branch D only), nothing deposited."""
    g = cartesian_grid(15, 11)
    return g, default_params(), _const(0.6, -0.5), 600.0, 3


def sc_tiny():
    """3 x 3: a single interior particle whose deposits fall on grid-boundary nodes."""
    g = cartesian_grid(3, 3)
    return g, default_params(), _const(10.0, -10.0), 600.0, 4


def sc_odd_periodic_strip():
    """periodic in both axes with odd sizes and Nx not a multiple of 4 (pitched record rows),
    fast particles: reach 2-3, wrap in x and y, two deposit classes absent."""
    g = cartesian_grid(17, 13, dx=500.0, dy=400.0, bx=BND_PERIODIC, by=BND_PERIODIC)
    return g, default_params(periodic_boundary=True), _const(-13.0, 11.0), 1200.0, 5


def sc_fast_box():
    """fine Cartesian box (150 x 90, larger than a gather tile) with particles crossing 2-4 cells
    per step: reach 2, then 3, then 4 — the reach-2 window and the wide (12 + 4 rows) tile
    geometry of the gather, on interior nodes."""
    g = cartesian_grid(150, 90, dx=700.0, dy=800.0)
    return g, default_params(DT=900.0), _const(13.0, -9.0), 900.0, 6


SCENARIOS = {
    "minimal": sc_minimal,
    "minimal_dp5": sc_minimal_dp5,
    "periodic_model_flag": sc_periodic_model_flag,
    "periodic_grid": sc_periodic_grid,
    "periodic_x_only": sc_periodic_x_only,
    "land_block": sc_land_block,
    "growing_winds": sc_growing_winds,
    "growing_winds_persist": sc_growing_winds_persist,
    "pulse_winds": sc_pulse_winds,
    "tripolar": sc_tripolar,
    "tripolar_tall": sc_tripolar_tall,
    "tripolar_propagation_only": sc_tripolar_propagation_only,
    "emax_clamp": sc_emax_clamp,
    "maxiters": sc_maxiters,
    "dtmin_no_force": sc_dtmin_no_force,
    "dp5_blowup": sc_dp5_blowup,
    "nan_wind": sc_nan_wind,
    "nan_defaults": sc_nan_defaults,
    "inf_defaults": sc_inf_defaults,
    "all_land": sc_all_land,
    "calm": sc_calm,
    "tiny": sc_tiny,
    "odd_periodic_strip": sc_odd_periodic_strip,
    "fast_box": sc_fast_box,
}


def run_pair(ref, dut, wind, DT, nsteps, compare, every=1):
    """seed + nsteps on both models, comparing after the seed and every `every` steps."""
    u0, v0 = wind(0.0)
    ref.seed(u0, v0)
    dut.seed(u0, v0)
    compare(ref, dut)
    t = 0.0
    for k in range(nsteps):
        ut, vt = wind(t)
        ut1, vt1 = wind(t + DT)
        ref.step(t, DT, ut, vt, ut1, vt1)
        dut.step(t, DT, ut, vt, ut1, vt1)
        t += DT
        if (k + 1) % every == 0 or k == nsteps - 1:
            compare(ref, dut)
