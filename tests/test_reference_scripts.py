"""The reference's own scripts — its only "tests" are interactive scripts a human looks at (tests/runtests.jl is empty,
SURVEY §4) — restated statement for statement on the Python mirror of its API, so that they read like the originals:
    tests/T04_2D_reg_test.jl:40-151              the sign / periodicity sweep of the box (movie_time_step!)
    tests/S02_2D_box_mesh_grid_single_steps.jl   x-periodic box with a land block, ParticleDefaults, source terms off,
                                                 bare time_step! calls
    tests/T04_2D_on_off_particle_tests.jl:40-152 calm foot + linear ramp, wind_min_squared = 2, Δt = DT/2 with the seed
                                                 time scale DT (the legacy TwoDGrid replaced by the mesh grid it became)
Instead of a GIF each script is checked three ways:
  * the mirror API drives the host build of the DEVICE code (ShimEngine: the CUDA path's arithmetic on the CPU) — what
    `architecture = B200()` runs on a GPU (tests/test_gpu_host_api.py);
  * the oracle, driven by hand with the same staged winds, must equal it bit for bit;
  * the third reading of the model step (tests/test_independent_model.py, written from the Julia sources, closures
    called as the reference calls them) must agree with the oracle in closure mode to rounding level.
Sizes are the scripts'; the number of steps is cut to keep the suite short."""
import math

import numpy as np
import pytest

pytest.importorskip("mpmath")

from common import ShimEngine, bits_equal, compare_models, grid_dict_from_mesh, make_oracle  # noqa: E402
from test_independent_model import RefModel  # noqa: E402

from picles_b200 import FetchRelations  # noqa: E402
from picles_b200.Grids.CartesianGrid import TwoDCartesianGridMesh  # noqa: E402
from picles_b200.Models.WaveGrowthModels2D import WaveGrowth2D  # noqa: E402
from picles_b200.Operators.core_2D import ParticleDefaults  # noqa: E402
from picles_b200.Operators.TimeSteppers import movie_time_step, time_step  # noqa: E402
from picles_b200.ParticleSystems import particle_waves_v5 as PW  # noqa: E402
from picles_b200.Simulations import Simulation, initialize_simulation  # noqa: E402

minutes, hours, days = 60.0, 3600.0, 86400.0


def three_ways(model, u, v, Δt, nsteps, mode, rtol=2e-9, third_steps=None):
    """mode: 'run' (State .= 0 before every time_step!, run.jl:75-82), 'bare' (time_step! adds to State) or 'movie'
    (movie_time_step!: adds, State .= 0 after the remesh; MovieState is the field a frame shows)"""
    grid = model.grid
    g = grid_dict_from_mesh(grid)
    g["mask_py"] = np.asarray(g["mask"], np.int64)
    model._engine = ShimEngine(g, model.params)
    model._rows = slice(0, model.Ny)
    sim = Simulation(model, Δt=Δt, stop_time=nsteps * Δt)
    initialize_simulation(sim)
    P = model.params
    d = model.ODEdefaults
    defaults = None if d is None else [d.lne, d.c̄_x, d.c̄_y, 0.0, 0.0]
    wind = lambda x, y, t: (float(u(x, y, t)), float(v(x, y, t)))
    sample = lambda t: tuple(np.array([[wind(g["x"][j, i], g["y"][j, i], t)[k] for i in range(g["Nx"])] for j in range(g["Ny"])])
                             for k in (0, 1))
    # The scripts leave `solver` at the ODESettings default, AutoTsit5(Rosenbrock23()), and with their C_φ = c_β = 4e-2
    # (2 000 x ODEParameters' 1.81e-5) the direction relaxation IS stiff: every particle of the reg-test sweep goes over
    # to Rosenbrock23 in its first model step (961 switches of 961).  The mirror + device code + staged oracle run that
    # default.  The third reading has no Rosenbrock23 branch, and plain Tsit5 on this parameter set sits on its stability
    # boundary (279 substeps per step, c̄_x != c̄_y under a symmetric wind, results that move by 1e-4 with the last bit
    # of the seed) — nothing to compare to rounding level.  So the third way runs only where the default solver never
    # switched in the script (calm cases); tests/test_independent_model.py covers the model step on non-stiff sets.
    probe = make_oracle(g, P)
    probe.seed(*sample(0.0))
    probe.set_accumulate(0 if mode == "run" else 1)
    switches = 0
    for k in range(nsteps):
        probe.step(k * Δt, Δt, *sample(k * Δt), *sample((k + 1) * Δt))
        switches += probe.counters()["n_stiff_switches"] + probe.counters()["n_stiff_attempts"]
        if mode == "movie":
            probe.set_state(np.zeros((3, g["Ny"], g["Nx"])))
    third = switches == 0
    P3 = type(P).from_buffer_copy(P)
    if int(P3.solver) == 2:
        P3.solver = 0
    staged, exact = make_oracle(g, P), make_oracle(g, P3)
    exact.set_wind_closure(wind, g["x"], g["y"])
    for o in (staged, exact):
        o.seed(*sample(0.0))
        o.set_accumulate(0 if mode == "run" else 1)
    solver = {0: "Tsit5", 1: "DP5"}[int(P3.solver)]
    ref = RefModel(g, P3, wind, solver, float(P.seed_timescale), defaults)
    ref.minimal_state = [float(P.minimal_state[0]), float(P.minimal_state[1])]
    ref.seed()
    assert bits_equal(np.asarray(model.State).transpose(2, 1, 0), staged.state())          # init_z0_to_State!
    t = 0.0
    for k in range(nsteps):
        if mode == "movie":
            movie_time_step(model, Δt)
            shown = np.asarray(model.MovieState).transpose(2, 1, 0)
        else:
            time_step(model, Δt, zero_state_first=(mode == "run"))
            shown = np.asarray(model.State).transpose(2, 1, 0)
        staged.step(t, Δt, *sample(t), *sample(t + Δt))
        assert bits_equal(shown, staged.state()), (k, "mirror API + device code against the oracle")
        if third and (third_steps is None or k < third_steps):
            exact.step(t, Δt, *sample(t), *sample(t + Δt))
            ref.step(Δt, zero_first=(mode == "run"), zero_after=False)
            So = exact.state()
            mn = np.hypot(ref.S[1], ref.S[2])
            scale = np.maximum(np.stack([np.abs(ref.S[0]), mn, mn]), 1e-300)
            with np.errstate(invalid="ignore"):
                bad = np.abs(So - ref.S) > rtol * scale + 1e-300
            assert not bad.any(), (k, "third reading against the closure-mode oracle", np.argwhere(bad)[:3])
        t += Δt
        if mode == "movie":
            for o in (staged, exact):
                o.set_state(np.zeros((3, g["Ny"], g["Nx"])))
            ref.S[:] = 0.0
        assert model.clock.time == pytest.approx(t)
    compare_models(staged, model.engine)
    staged.stiff_work_seen, staged.third_way_ran = switches, third
    return model, staged, ref


def settings(U10, V10, DT, **kw):
    """the ODESettings block the three scripts share (T04_2D_reg_test.jl:62-84), with their parameter tuple
    `(r_g, C_α = Const_Scg.C_alpha, C_φ = Const_ID.c_β, C_e = Const_ID.C_e, g)` — C_φ is NOT ODEParameters' here"""
    ODEpars, Const_ID, Const_Scg = PW.ODEParameters(r_g=0.85)
    default_ODE_parameters = dict(r_g=0.85, C_α=Const_Scg.C_alpha, C_φ=Const_ID.c_β, C_e=Const_ID.C_e, g=9.81)
    WindSeamin = FetchRelations.MinimalWindsea(U10, V10, DT)
    args = dict(Parameters=default_ODE_parameters, log_energy_minimum=WindSeamin["lne"], log_energy_maximum=math.log(17),
                saving_step=DT, timestep=DT, total_time=6 * days, adaptive=True, dt=1e-3, dtmin=1e-4, force_dtmin=True,
                callbacks=None, save_everystep=False)
    args.update(kw)
    return PW.ODESettings(**args), Const_ID


@pytest.mark.parametrize("periodic", [True, False])
@pytest.mark.parametrize("U10,V10", [(i, j) for i in (-10, 0, 10) for j in (-10, 0, 10)])
def test_T04_2D_reg_test_sweep(U10, V10, periodic):
    """tests/T04_2D_reg_test.jl:122-151: `for (U10, V10, per) in gridmesh` … `movie_time_step!` per frame (4 frames
    here, 36 there).  (0, 0) seeds every particle off (MinimalParticle with rand_sign() = +1, B-9)."""
    DT = 10 * minutes
    grid = TwoDCartesianGridMesh(120e3, 31, 120e3, 31)
    u = lambda x, y, t: U10 + x * 0 + y * 0 + t * 0
    v = lambda x, y, t: V10 + x * 0 + y * 0 + t * 0
    ODE_settings, Const_ID = settings(5.0, 5.0, DT)          # the settings block is built once, with U10, V10 = 5, 5 (:43)
    particle_system = PW.particle_equations(u, v, γ=Const_ID.γ, q=Const_ID.q)
    wave_model = WaveGrowth2D(grid=grid, winds=dict(u=u, v=v), ODEsys=particle_system, ODEsets=ODE_settings,
                              ODEinit_type="wind_sea", periodic_boundary=periodic, boundary_type="same",
                              minimal_particle=FetchRelations.MinimalParticle(U10, V10, DT), movie=True)
    model, o, ref = three_ways(wave_model, u, v, DT, 4, "movie", third_steps=2)   # the pure-Python third way: two frames
    c = o.counters()
    if (U10, V10) == (0, 0):
        assert c["n_integrated"] == 0 and c["n_remesh_D"] == c["n_active"]
    else:
        assert c["n_integrated"] == c["n_active"] == (31 * 31 if periodic else 29 * 29) and c["n_failed"] == 0


def test_S02_2D_box_mesh_grid_single_steps():
    """tests/S02_2D_box_mesh_grid_single_steps.jl:46-172: wind on the left half only, x periodic / y open, a land block,
    every particle seeded from ParticleDefaults(log 5, 5, 5), propagation and direction only, five bare time_step!
    calls (State is never zeroed: it accumulates)."""
    U10, V10 = 15.0, 10.0
    DT = 20 * minutes
    u = lambda x, y, t: (U10 if x < 250e3 else 0.00) + y * 0.0 + t * 0.0
    v = lambda x, y, t: (V10 if x < 250e3 else 0.00) + y * 0.0 + t * 0.0
    mask = np.ones((51, 41), dtype=bool)
    mask[19:35, 19:35] = False                                # mask[20:35, 20:35] .= 0
    grid = TwoDCartesianGridMesh(500e3, 51, 400e3, 41, periodic_boundary=(True, False), mask=mask)
    ODE_settings, Const_ID = settings(U10, V10, DT, log_energy_maximum=math.log(27))
    particle_system = PW.particle_equations(u, v, γ=Const_ID.γ, q=Const_ID.q, propagation=True, input=False,
                                            dissipation=False, peak_shift=False, direction=True)
    wave_model = WaveGrowth2D(grid=grid, winds=dict(u=u, v=v), ODEsys=particle_system, ODEsets=ODE_settings,
                              ODEinit_type=ParticleDefaults(math.log(5), 5.0, 5.0, 0.0, 0.0), periodic_boundary=False,
                              boundary_type="same", movie=True)
    model, o, ref = three_ways(wave_model, u, v, 20 * minutes, 5, "bare", third_steps=2)
    S = np.asarray(model.State)
    assert (grid.data.mask == 2).sum() > 0 and (grid.data.mask == 0).sum() > 0
    assert S[:, :, 0].max() > 5.0 * 2                      # accumulated over the steps: more than one particle's energy


@pytest.mark.parametrize("U10,V10", [(-8.0, 0.0), (8.0, 0.0)])
def test_T04_2D_on_off_particle_tests(U10, V10):
    """tests/T04_2D_on_off_particle_tests.jl:84-152: no wind at all left of x0 (exactly 0: the particles there are
    seeded off from MinimalParticle), a linear ramp to the right, wind_min_squared = 2, the model stepped with
    Δt = DT/2 while the seed and reseed time scale stays DT (SURVEY B-5); run!'s loop (State .= 0 every step)."""
    DT = 30 * minutes
    grid = TwoDCartesianGridMesh(100e3, 21, 50e3, 11)
    x0 = 50e3
    Lx = (21 - 1) * grid.stats.dx
    u = lambda x, y, t: (x * 0 + 0 if x < x0 else U10 * (x - x0) / (Lx - x0)) + y * 0 + t * 0
    v = lambda x, y, t: (x * 0 + 0 if x < x0 else V10 * (x - x0) / (Lx - x0)) + y * 0 + t * 0
    ODE_settings, Const_ID = settings(U10, V10, DT, log_energy_maximum=math.log(27), wind_min_squared=2.0)
    particle_system = PW.particle_equations(u, v, γ=Const_ID.γ, q=Const_ID.q)
    wave_model = WaveGrowth2D(grid=grid, winds=dict(u=u, v=v), ODEsys=particle_system, ODEsets=ODE_settings,
                              ODEinit_type="wind_sea", periodic_boundary=False, boundary_type="same",
                              minimal_particle=FetchRelations.MinimalParticle(U10, V10, DT),
                              minimal_state=FetchRelations.MinimalState(2, 2, DT) * 1, movie=True)
    model, o, ref = three_ways(wave_model, u, v, DT / 2, 6, "run")
    seen = ref.seen if hasattr(ref, "seen") else {}
    c = o.counters()
    assert c["n_integrated"] > 0 and c["n_remesh_D"] > 0 and c["n_integrated"] < c["n_active"]
    del seen
