"""Host-side mirror of the reference interface (WaveGrowth2D / Simulation / run! / time_step! /
movie_time_step! / grids / masks), exercised on CPU with the host build of the device code
standing in for the GPU engine, and compared with the oracle driven by hand."""
import math

import numpy as np
import pytest

from common import (ShimEngine, bits_equal, compare_models, default_params, grid_dict_from_mesh, make_oracle)

from picles_b200 import FetchRelations
from picles_b200.Architectures import B200, CPU, N_NonPeriodic, N_Periodic, N_TripolarNorth
from picles_b200.Grids.CartesianGrid import ProjetionKernel, TwoDCartesianGridMesh
from picles_b200.Grids.mask_utils import make_boundaries, make_boundary_lists
from picles_b200.Grids.TripolarGridMOM6 import synthetic_MOM6GridMesh
from picles_b200.Models.WaveGrowthModels2D import WaveGrowth2D
from picles_b200.Operators.core_2D import GetGroupVelocity, ParticleDefaults
from picles_b200.Operators.TimeSteppers import movie_time_step, time_step
from picles_b200.ParticleSystems import particle_waves_v5 as PW
from picles_b200.Simulations import Simulation, initialize_simulation, reset_simulation, run

minutes, hours, days = 60.0, 3600.0, 86400.0


def example_00_minimal(winds_uv=(10.0, 10.0), grid=None, **model_kw):
    """examples/example_00_minimal.jl:17-67, line by line."""
    U10, V10 = winds_uv
    DT = 10 * minutes
    u = lambda x, y, t: U10
    v = lambda x, y, t: V10
    grid = grid or TwoDCartesianGridMesh(100e3, 51, 100e3, 51)
    ODEpars, Const_ID, Const_Scg = PW.ODEParameters(r_g=0.85)
    particle_system = PW.particle_equations(u, v, γ=Const_ID.γ, q=Const_ID.q)
    WindSeamin = FetchRelations.MinimalWindsea(U10, V10, DT)
    ODE_settings = PW.ODESettings(Parameters=ODEpars, log_energy_minimum=WindSeamin["lne"], saving_step=DT, timestep=DT,
                                  total_time=6 * days, dt=1e-3, dtmin=1e-4, force_dtmin=True)
    kw = dict(periodic_boundary=False, minimal_particle=FetchRelations.MinimalParticle(U10, V10, DT), movie=True)
    kw.update(model_kw)
    model = WaveGrowth2D(grid=grid, winds=dict(u=u, v=v), ODEsys=particle_system, ODEsets=ODE_settings, **kw)
    return model, DT


def attach_shim(model):
    model._engine = ShimEngine(grid_dict_from_mesh(model.grid), model.params)
    model._rows = slice(0, model.Ny)
    return model


def test_example_00_minimal_run_matches_oracle():
    model, DT = example_00_minimal()
    attach_shim(model)
    sim = Simulation(model, Δt=DT, stop_time=2 * hours)
    run(sim, cash_store=True)
    # '>=' in run! => floor(stop/Δt)+1 = 13 steps, 14 cached states (initial + one per step)
    assert model.clock.iteration == 13 and model.clock.time == 13 * DT
    assert len(sim.store.store) == 14
    g = grid_dict_from_mesh(model.grid)
    o = make_oracle(g, default_params())
    o.seed(10.0, 10.0)
    assert bits_equal(sim.store.store[0].transpose(2, 1, 0), o.state())
    t = 0.0
    for k in range(13):
        o.step(t, DT, 10.0, 10.0, 10.0, 10.0)
        t += DT
        assert bits_equal(sim.store.store[k + 1].transpose(2, 1, 0), o.state()), k
    compare_models(o, model.engine)
    S = model.State
    assert S.shape == (51, 51, 3)
    hs = 4 * np.sqrt(S[:, :, 0])
    assert 1.5 < hs.max() < 2.5          # ~1.9 m after 2 h of 14 m/s wind
    c = GetGroupVelocity(S)
    assert np.nanmax(c["c_x"]) > 1.0


def test_params_flattening_matches_reference_defaults():
    model, DT = example_00_minimal()
    P = model.params
    assert P.r_g == 0.85 and P.C_alpha == -1.41 and P.C_varphi == 1.81e-5
    assert P.C_e == pytest.approx(2.2117647058823533e-4, rel=1e-15)
    assert (P.p, P.q, P.n) == (0.75, -0.25, 2.0)
    assert P.e_T == pytest.approx(0.5040608763647848, rel=1e-15)
    assert P.abstol == 1e-4 and P.reltol == 1e-3 and P.maxiters == 10000
    assert P.log_energy_maximum == math.log(17) and P.wind_min_squared == 4.0
    assert list(P.minimal_state) == FetchRelations.MinimalState(2, 2, DT)
    assert P.minimal_state[0] == pytest.approx(1.253106339976604e-6, rel=1e-14)
    assert P.has_defaults == 0 and P.periodic_boundary == 0
    assert P.solver == 2                                      # ODESettings default: AutoTsit5(Rosenbrock23())
    assert P.dtmax == 6 * days


def test_ocean_points_lists_and_order():
    g = TwoDCartesianGridMesh(100e3, 11, 100e3, 7)
    m, _ = example_00_minimal(grid=g, periodic_boundary=False)
    assert len(m.ocean_points) == 9 * 5 and len(m.boundary_points) == 2 * 11 + 2 * 5
    assert m.ocean_points[0].tolist() == [2, 2] and m.ocean_points[1].tolist() == [3, 2]   # column-major, i fastest
    m2, _ = example_00_minimal(grid=g, periodic_boundary=True)
    # model periodic flag: grid-boundary nodes join ocean_points AFTER the ocean nodes
    assert len(m2.ocean_points) == 11 * 7 and m2.ocean_points[45].tolist() == [1, 1]
    o = make_oracle(grid_dict_from_mesh(g), m2.params)
    op = o.ocean_points()
    ref = [(int(l % 11) + 1, int(l // 11) + 1) for l in op]
    assert ref == [tuple(x) for x in m2.ocean_points.tolist()]


def test_bare_time_step_accumulates_and_movie_time_step_zeroes():
    model, DT = example_00_minimal(winds_uv=(8.0, -6.0))
    attach_shim(model)
    sim = Simulation(model, Δt=DT, stop_time=1 * hours)
    initialize_simulation(sim)
    o = make_oracle(grid_dict_from_mesh(model.grid), model.params)
    o.set_accumulate(True)
    o.seed(8.0, -6.0)
    assert bits_equal(model.State.transpose(2, 1, 0), o.state())
    time_step(model, DT)                 # adds to the seeded State, like the reference
    o.step(0.0, DT, 8.0, -6.0, 8.0, -6.0)
    assert bits_equal(model.State.transpose(2, 1, 0), o.state())
    movie_time_step(model, DT)
    o.step(DT, DT, 8.0, -6.0, 8.0, -6.0)
    assert bits_equal(model.MovieState.transpose(2, 1, 0), o.state())
    assert not model.State.any()
    assert model.clock.iteration == 2
    reset_simulation(sim)
    assert model.clock.time == 0.0 and not model.State.any()


def test_time_varying_wind_closures_are_staged_at_t_and_t_plus_dt():
    U = lambda x, y, t: 6.0 + 4.0 * np.sin(2 * np.pi * t / 7200.0) + x * 1e-5
    V = lambda x, y, t: 3.0 - 2e-5 * y + 1e-3 * t
    g = TwoDCartesianGridMesh(60e3, 31, 40e3, 21)
    ODEpars, cid, _ = PW.ODEParameters(r_g=0.85)
    ps = PW.particle_equations(U, V, γ=cid.γ, q=cid.q)
    sets = PW.ODESettings(Parameters=ODEpars, log_energy_minimum=-13.0, saving_step=600.0, timestep=600.0,
                          total_time=6 * days, dt=1e-3, dtmin=1e-4, force_dtmin=True)
    model = attach_shim(WaveGrowth2D(grid=g, winds=(U, V), ODEsys=ps, ODEsets=sets, periodic_boundary=False))
    sim = Simulation(model, Δt=600.0, stop_time=3000.0)
    run(sim)
    gd = grid_dict_from_mesh(g)
    o = make_oracle(gd, model.params)
    w = lambda t: (U(gd["x"], gd["y"], t), V(gd["x"], gd["y"], t))
    o.seed(*w(0.0))
    t = 0.0
    for _ in range(6):
        o.step(t, 600.0, *w(t), *w(t + 600.0))
        t += 600.0
    compare_models(o, model.engine)


def test_scalar_only_closures_fall_back_to_pointwise_evaluation():
    import math as m
    U = lambda x, y, t: 5.0 + m.sin(x / 1e4)       # math.sin rejects arrays
    V = lambda x, y, t: 4.0
    g = TwoDCartesianGridMesh(20e3, 9, 20e3, 8)
    model, _ = example_00_minimal(grid=g)
    model.winds.u, model.winds.v = U, V
    model._engine = None
    attach_shim(model)
    u, v = model._wind_planes(0.0)
    assert u.shape == (8, 9) and u[3, 4] == U(g.data.x[4, 3], 0, 0) and (v == 4.0).all()


def test_grid_types_and_kernels():
    g = TwoDCartesianGridMesh(100e3, 51, 100e3, 51)
    assert isinstance(g.stats.Nx, N_NonPeriodic) and g.stats.dx == 2000.0
    assert ProjetionKernel(g.stats).tolist() == [1 / 2000.0, 0.0, 0.0, 1 / 2000.0]
    gp = TwoDCartesianGridMesh(0.0, 10.0, 11, 0.0, 5.0, 6, periodic_boundary=(True, False), angle=30.0)
    assert isinstance(gp.stats.Nx, N_Periodic) and isinstance(gp.stats.Ny, N_NonPeriodic)
    assert (gp.data.mask[:, 0] == 3).all() and (gp.data.mask[0, 1:-1] == 1).all()
    M = ProjetionKernel(gp.stats)                   # rotated kernel has NO minus sign (SURVEY B-8)
    assert M[1] > 0 and M[2] > 0
    t = synthetic_MOM6GridMesh(36, 30, k=2)
    assert isinstance(t.stats.Nx, N_Periodic) and isinstance(t.stats.Ny, N_TripolarNorth)
    met = t.device_metric()
    assert met["M"].shape == (4, 36, 30) and (met["M"][2] == -np.sin(t.data.angle_dx * np.pi / 180) / t.data.dx).all()
    assert not (t.data.mask[:, 0] == 1).any()       # south cap masked by TripolarGrid_mask_pols!


def test_unsupported_requests_fail_loudly():
    with pytest.raises(NotImplementedError):
        example_00_minimal(architecture=CPU())
    with pytest.raises(NotImplementedError):
        example_00_minimal(layers=2)
    model, DT = example_00_minimal()
    attach_shim(model)
    with pytest.raises(RuntimeError):
        time_step(model, DT)            # not seeded yet
    with pytest.raises(ValueError):
        PW.ODESettings(Parameters={}, log_energy_minimum=0, saving_step=1, timestep=1, total_time=1,
                       solver="Rosenbrock23").solver_id()
    assert B200().devices == (0,)


# ---- wind ingestion through the host API ---------------------------------------------------

def test_wind_levels_option_stages_intermediate_levels():
    """B200(wind_levels=4): the closures are evaluated at 4 equally spaced times of every step
    and the run equals the oracle fed with the same levels by hand"""
    w = 5.0 / (3600.0 * 2.0 * math.pi)
    u = lambda x, y, t: 15.0 + 0.0 * x
    v = lambda x, y, t: -10.0 * math.cos(w * t) + 0.0 * x
    model, DT = example_00_minimal(grid=TwoDCartesianGridMesh(40e3, 11, 30e3, 9), architecture=B200(wind_levels=4))
    model.winds.u, model.winds.v = u, v
    attach_shim(model)
    sim = Simulation(model, Δt=DT, stop_time=30 * minutes)
    run(sim)
    g = grid_dict_from_mesh(model.grid)
    o = make_oracle(g, default_params())
    full = lambda val: np.full((g["Ny"], g["Nx"]), val)
    o.seed(full(15.0), full(v(0, 0, 0.0)))
    t = 0.0
    for _ in range(model.clock.iteration):
        tm = [t + DT * float(k) / 3.0 for k in (1, 2)]
        o.set_wind_midlevels([full(15.0) for _ in tm], [full(v(0, 0, x)) for x in tm])
        o.step(t, DT, full(15.0), full(v(0, 0, t)), full(15.0), full(v(0, 0, t + DT)))
        t += DT
    compare_models(o, model.engine)
    with pytest.raises(ValueError):
        B200(wind_levels=7)


def test_gridded_winds_are_sampled_by_the_engine():
    """winds=wind_interpolator(wind_grid): the mesh is handed to the engine once and every level
    of every step comes from the sampler; the run equals the oracle driven with the oracle's own
    restatement of LinearInterpolation(..., extrapolation_bc=Periodic())"""
    import oracle
    from picles_b200.Utils.WindEmulator import wind_interpolator
    rng = np.random.default_rng(5)
    xi = np.linspace(-5e3, 45e3, 6)
    yi = np.linspace(-5e3, 35e3, 5)
    ti = np.array([0.0, 900.0, 1800.0, 3600.0])
    ug = 9.0 + 3.0 * rng.random((xi.size, yi.size, ti.size))
    vg = 6.0 + 3.0 * rng.random((xi.size, yi.size, ti.size))
    winds = wind_interpolator(dict(u=ug, v=vg, x=xi, y=yi, t=ti))
    with pytest.raises(RuntimeError):
        winds.u(0.0, 0.0, 0.0)           # no CPU evaluation path
    model, DT = example_00_minimal(grid=TwoDCartesianGridMesh(40e3, 11, 30e3, 9), architecture=B200(wind_levels=3))
    model.winds = winds
    model._gridded_winds = winds
    attach_shim(model)
    g = grid_dict_from_mesh(model.grid)
    winds.bind(model.engine, g["x"], g["y"])
    sim = Simulation(model, Δt=DT, stop_time=40 * minutes)
    run(sim)
    assert winds.u(None, None, 450.0).shape == (11, 9)
    o = make_oracle(g, default_params())
    U, V = ug.transpose(2, 1, 0), vg.transpose(2, 1, 0)
    samp = lambda t: oracle.wind_mesh_sample(xi, yi, ti, U, V, g["x"], g["y"], t)
    o.seed(*samp(0.0))
    t = 0.0
    for _ in range(model.clock.iteration):
        um, vm = samp(t + DT * 1.0 / 2.0)
        o.set_wind_midlevels([um], [vm])
        o.step(t, DT, *samp(t), *samp(t + DT))
        t += DT
    compare_models(o, model.engine)


def test_spherical_grid_mesh_through_the_model():
    """TwoDSphericalGridMesh (src/Grids/SphericalGrid.jl): spacing in metres from the lon/lat mesh,
    the per-node kernel exactly as the reference writes it (cos of dy*pi/180, dy in metres) and the
    great-circle coefficient; a model on it steps like the oracle given the same planes"""
    from picles_b200.Grids.SphericalGrid import TwoDSphericalGridMesh, cal_dx_meters
    grid = TwoDSphericalGridMesh(-20.0, 20.0, 21, 10.0, 40.0, 16, periodic_boundary=(True, False))
    st, d = grid.stats, grid.data
    assert isinstance(st.Nx, N_Periodic) and isinstance(st.Ny, N_NonPeriodic)
    assert st.dx_deg == 2.0 and st.dy_deg == 2.0
    assert d.x[0, 0] == -20.0 and d.x[-1, 0] == 20.0 and d.y[0, -1] == 40.0
    R = 6371.0e3
    assert d.dy[3, 5] == pytest.approx(2.0 * np.pi / 180 * R, rel=1e-15)
    assert d.dx[3, 5] == pytest.approx(2.0 * np.pi / 180 * R * np.cos(np.deg2rad(d.y[3, 5])), rel=1e-14)
    assert d.dx[0, 5] == pytest.approx(d.dx[3, 5], rel=1e-14)            # one-sided end difference, same spacing
    met = grid.device_metric()
    i, j = 4, 7
    assert met["M"][0][i, j] == 1.0 / (np.cos(d.dy[i, j] * np.pi / 180) * d.dx[i, j])      # the reference's formula, as is
    assert met["M"][3][i, j] == 1.0 / d.dy[i, j] and np.all(met["M"][1] == 0) and np.all(met["M"][2] == 0)
    assert met["pc"][i, j] == pytest.approx(np.tan(np.deg2rad(d.y[i, j])) / 6.3710e6, rel=1e-14)
    assert np.all(d.mask[:, 0] == 3) and np.all(d.mask[:, -1] == 3) and np.all(d.mask[:, 1:-1] == 1)   # x periodic
    model, DT = example_00_minimal(grid=grid, periodic_boundary=True)
    attach_shim(model)
    sim = Simulation(model, Δt=DT, stop_time=30 * minutes)
    run(sim)
    g = grid_dict_from_mesh(grid)
    o = make_oracle(g, default_params(periodic_boundary=True))
    o.seed(10.0, 10.0)
    t = 0.0
    for _ in range(model.clock.iteration):
        o.step(t, DT, 10.0, 10.0, 10.0, 10.0)
        t += DT
    compare_models(o, model.engine)
    assert np.nanmax(model.State[:, :, 0]) > 0


def test_nan_eest_rejects_reaches_the_parameter_struct():
    """B200(nan_eest_rejects=True) (or the model keyword) sets picles_params_t::nan_eest_rejects; the default is 0"""
    model, _ = example_00_minimal()
    assert model.params.nan_eest_rejects == 0
    model, _ = example_00_minimal(architecture=B200(nan_eest_rejects=True))
    assert model.params.nan_eest_rejects == 1
    model, _ = example_00_minimal(nan_eest_rejects=True)
    assert model.params.nan_eest_rejects == 1


@pytest.mark.parametrize("seed", range(6))
def test_random_scripts_through_the_mirror_api(seed):
    """a random user script — grid size and spacing, periodic axes, land, the model's periodic flag, solver (the default
    AutoTsit5 included), thresholds, B200(wind_levels = 2, 3 or 5), wind closures varying in space and time — driven through
    a mix of run!-style steps, bare time_step!, movie_time_step! with a CHANGING Δt, then reset_simulation! and run!(...,
    cash_store=true): at every point the fields equal the oracle driven by hand with the winds staged the way the library
    documents, bit for bit (60 seeds were run once)."""
    rng = np.random.default_rng(seed)
    Nx, Ny = int(rng.integers(5, 12)), int(rng.integers(5, 10))
    d = float(rng.choice([500.0, 1000.0, 2000.0]))
    per = (bool(rng.random() < 0.5), bool(rng.random() < 0.5))
    mask = (rng.random((Nx, Ny)) > 0.12) if rng.random() < 0.5 else None
    grid = TwoDCartesianGridMesh(d * (Nx - 1), Nx, d * (Ny - 1), Ny, periodic_boundary=per, mask=mask)
    DT = float(rng.choice([600.0, 900.0, 1200.0]))
    a, b, c = rng.uniform(-9, 9, 3)
    d_, e, f = rng.uniform(-9, 9, 3)
    om = 2 * math.pi / float(rng.choice([3600.0, 7200.0, 1e9]))
    calm = rng.uniform(0, 0.6)
    x1, y1 = max(d * (Nx - 1), 1.0), max(d * (Ny - 1), 1.0)

    def w(x, y, t):
        sx, sy = x / x1, y / y1
        sw = 1.0 + 0.9 * math.sin(om * t + 2.0 * sy)
        if sx < calm:
            return 1.1 * sw, -0.6 * sw
        return (a + b * sx) * (1.0 + 0.3 * math.sin(om * t)) + c * sy, d_ + e * sy + f * math.cos(om * t)

    u = lambda x, y, t: w(x, y, t)[0]
    v = lambda x, y, t: w(x, y, t)[1]
    ODEpars, CID, _ = PW.ODEParameters(r_g=0.85)
    solver = str(rng.choice(["Tsit5", "DP5", "AutoTsit5"]))
    S = PW.ODESettings(Parameters=ODEpars, log_energy_minimum=FetchRelations.MinimalWindsea(10, 10, DT)["lne"],
                       log_energy_maximum=float(rng.choice([math.log(17), math.log(3e-3)])),
                       wind_min_squared=float(rng.choice([2.0, 4.0])), saving_step=DT, timestep=DT, total_time=6 * days,
                       dt=1e-3, dtmin=1e-4, force_dtmin=True, solver=solver)
    levels = int(rng.choice([2, 2, 3, 5]))
    m = WaveGrowth2D(grid=grid, winds=dict(u=u, v=v), ODEsys=PW.particle_equations(u, v, γ=CID.γ, q=CID.q), ODEsets=S,
                     periodic_boundary=bool(rng.random() < 0.5), architecture=B200(wind_levels=levels), movie=True)
    attach_shim(m)
    g = grid_dict_from_mesh(m.grid)
    wa = lambda t: tuple(np.array([[w(g["x"][j, i], g["y"][j, i], t)[k] for i in range(g["Nx"])] for j in range(g["Ny"])])
                         for k in (0, 1))
    clock = [0.0]

    def ostep(o, accumulate, dt):
        t = clock[0]
        o.set_accumulate(accumulate)
        nm = levels - 2
        if nm:
            lv = [wa(t + dt * float(k) / float(nm + 1)) for k in range(1, nm + 1)]
            o.set_wind_midlevels([x for x, _ in lv], [y for _, y in lv])
        o.step(t, dt, *wa(t), *wa(t + dt))
        clock[0] = t + dt

    o = make_oracle(g, m.params)
    o.seed(*wa(0.0))
    sim = Simulation(m, Δt=DT, stop_time=3 * DT)
    initialize_simulation(sim)
    assert bits_equal(np.asarray(m.State).transpose(2, 1, 0), o.state())
    for kind, dt in [("run", DT), ("bare", DT), ("movie", DT / 2), ("run", DT), ("bare", DT / 2)]:
        if kind == "run":
            time_step(m, dt, zero_state_first=True)
            ostep(o, 0, dt)
            shown = m.State
        elif kind == "bare":
            time_step(m, dt)
            ostep(o, 1, dt)
            shown = m.State
        else:
            movie_time_step(m, dt)
            ostep(o, 1, dt)
            shown = m.MovieState
        assert bits_equal(np.asarray(shown).transpose(2, 1, 0), o.state()), (kind, dt)
        if kind == "movie":
            o.set_state(np.zeros((3, g["Ny"], g["Nx"])))
    compare_models(o, m.engine)
    reset_simulation(sim)
    o = make_oracle(g, m.params)
    o.seed(*wa(0.0))
    clock[0] = 0.0
    sim.stop_time = 2 * DT
    run(sim, cash_store=True)
    for _ in range(3):
        ostep(o, 0, DT)
    assert len(sim.store.store) == 4 and bits_equal(np.asarray(sim.store.store[-1]).transpose(2, 1, 0), o.state())
