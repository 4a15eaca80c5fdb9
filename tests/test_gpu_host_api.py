"""The reference-facing API (WaveGrowth2D / Simulation / run!) on the real B200 architecture:
the scripts of the reference, line by line, against the oracle driven by hand."""
import numpy as np
import pytest

from common import bits_equal, compare_models, default_params, grid_dict_from_mesh, make_oracle
from test_host_api import example_00_minimal, hours

pytestmark = pytest.mark.gpu


def test_gpu_example_00_minimal_through_run(gpu_lib):
    """examples/example_00_minimal.jl on architecture=B200(): 13 steps, CashStore filled by
    asynchronous snapshots; every cached State equals the oracle's bit for bit."""
    from picles_b200.Architectures import B200
    from picles_b200.Simulations import Simulation, run
    model, DT = example_00_minimal(architecture=B200())
    sim = Simulation(model, Δt=DT, stop_time=2 * hours)
    run(sim, cash_store=True)
    assert model.clock.iteration == 13 and len(sim.store.store) == 14
    o = make_oracle(grid_dict_from_mesh(model.grid), default_params())
    o.seed(10.0, 10.0)
    assert bits_equal(sim.store.store[0].transpose(2, 1, 0), o.state())
    t = 0.0
    for k in range(13):
        o.step(t, DT, 10.0, 10.0, 10.0, 10.0)
        t += DT
        assert bits_equal(sim.store.store[k + 1].transpose(2, 1, 0), o.state()), k
    compare_models(o, model.engine)
    f = model.engine.fields()
    assert 1.5 < np.nanmax(f["Hs"]) < 2.5


def test_gpu_tripolar_model_with_device_formed_metric(gpu_lib):
    """MOM6GridMesh (synthetic super-grid) through WaveGrowth2D on B200: the per-node kernel
    and great-circle coefficient come from k_grid_metric; the oracle gets oracle.grid_metric
    of the same mesh planes."""
    import oracle
    from picles_b200 import FetchRelations
    from picles_b200.Architectures import B200
    from picles_b200.Grids.TripolarGridMOM6 import synthetic_MOM6GridMesh
    from picles_b200.Models.WaveGrowthModels2D import WaveGrowth2D
    from picles_b200.Operators.TimeSteppers import time_step
    from picles_b200.ParticleSystems import particle_waves_v5 as PW
    from picles_b200.Simulations import Simulation, initialize_simulation
    grid = synthetic_MOM6GridMesh(72, 60, k=2)
    DT = 1200.0
    u = lambda x, y, t: 15.0
    v = lambda x, y, t: -10.0 * np.cos(5 * t / (3600 * 2 * np.pi))
    ODEpars, Const_ID, _ = PW.ODEParameters(r_g=0.85)
    ps = PW.particle_equations(u, v, γ=Const_ID.γ, q=Const_ID.q)
    sets = PW.ODESettings(Parameters=ODEpars, log_energy_minimum=FetchRelations.MinimalWindsea(10, 10, DT)["lne"],
                          saving_step=DT, timestep=DT, total_time=6 * 86400.0, dt=1e-3, dtmin=1e-4, force_dtmin=True)
    model = WaveGrowth2D(grid=grid, winds=dict(u=u, v=v), ODEsys=ps, ODEsets=sets, periodic_boundary=True,
                         architecture=B200())
    sim = Simulation(model, Δt=DT, stop_time=3 * DT)
    initialize_simulation(sim)
    g = grid_dict_from_mesh(grid)
    T = lambda a: np.ascontiguousarray(np.asarray(a).T)
    M, pc = oracle.grid_metric(T(grid.data.dx), T(grid.data.dy), T(grid.data.angle_dx), T(grid.data.y))
    Me, pce = model.engine.metric()
    assert np.array_equal(M.view(np.uint64), Me.view(np.uint64)) and np.array_equal(pc.view(np.uint64), pce.view(np.uint64))
    g = dict(g, M=M, pc=pc)
    o = make_oracle(g, model.params)
    X, Y = T(grid.data.x), T(grid.data.y)
    W = lambda t: (np.full(X.shape, 15.0), np.full(X.shape, v(0, 0, t)))
    o.seed(*W(0.0))
    compare_models(o, model.engine)
    t = 0.0
    for _ in range(3):
        time_step(model, DT, zero_state_first=True)
        o.step(t, DT, *W(t), *W(t + DT))
        t += DT
        compare_models(o, model.engine)


def test_gpu_gridded_winds_through_run(gpu_lib):
    """winds = wind_interpolator(wind_grid) (src/Utils/WindEmulator.jl:18-43) on B200: the wind
    mesh is uploaded once and sampled on the device for every level of every step (wind_levels=3);
    `winds.u(x, y, t)` is served by the device too.  Oracle: its own restatement of the
    interpolation, levels fed by hand."""
    import oracle
    from picles_b200.Architectures import B200
    from picles_b200.Simulations import Simulation, run
    from picles_b200.Utils.WindEmulator import wind_interpolator
    rng = np.random.default_rng(2)
    xi, yi = np.linspace(-5e3, 110e3, 9), np.linspace(-5e3, 110e3, 8)
    ti = np.array([0.0, 1500.0, 3100.0, 7300.0])
    ug = 9.0 + 3.0 * rng.random((xi.size, yi.size, ti.size))
    vg = 6.0 + 3.0 * rng.random((xi.size, yi.size, ti.size))
    winds = wind_interpolator(dict(u=ug, v=vg, x=xi, y=yi, t=ti))
    model, DT = example_00_minimal(architecture=B200(wind_levels=3))
    model.winds = winds
    model._gridded_winds = winds
    sim = Simulation(model, Δt=DT, stop_time=1 * hours)
    run(sim)
    g = grid_dict_from_mesh(model.grid)
    U, V = ug.transpose(2, 1, 0), vg.transpose(2, 1, 0)
    samp = lambda t: oracle.wind_mesh_sample(xi, yi, ti, U, V, g["x"], g["y"], t)
    assert bits_equal(winds.u(None, None, 450.0).T, samp(450.0)[0])
    o = make_oracle(g, default_params())
    o.seed(*samp(0.0))
    t = 0.0
    for _ in range(model.clock.iteration):
        um, vm = samp(t + DT * 1.0 / 2.0)
        o.set_wind_midlevels([um], [vm])
        o.step(t, DT, *samp(t), *samp(t + DT))
        t += DT
    compare_models(o, model.engine)


def test_gpu_spherical_grid_model(gpu_lib):
    """TwoDSphericalGridMesh (src/Grids/SphericalGrid.jl) on B200: per-node kernel planes and the
    great-circle coefficient through picles_set_grid; x periodic, y closed"""
    from picles_b200.Architectures import B200
    from picles_b200.Grids.SphericalGrid import TwoDSphericalGridMesh
    from picles_b200.Simulations import Simulation, run
    grid = TwoDSphericalGridMesh(-20.0, 20.0, 41, 10.0, 40.0, 31, periodic_boundary=(True, False))
    model, DT = example_00_minimal(grid=grid, periodic_boundary=True, architecture=B200())
    sim = Simulation(model, Δt=DT, stop_time=1 * hours)
    run(sim)
    o = make_oracle(grid_dict_from_mesh(grid), default_params(periodic_boundary=True))
    o.seed(10.0, 10.0)
    t = 0.0
    for _ in range(model.clock.iteration):
        o.step(t, DT, 10.0, 10.0, 10.0, 10.0)
        t += DT
    compare_models(o, model.engine)
