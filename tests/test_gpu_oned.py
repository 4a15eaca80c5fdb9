"""The one-dimensional model on the B200 (picles1d_* through the C ABI) against the CPU oracle, bit for bit:
State, particle state, clocks, flags, status and counters after every step of every scenario."""
import numpy as np
import pytest

from scenarios_1d import SCENARIOS_1D, compare_models_1d, run_pair_1d

pytestmark = pytest.mark.gpu


def _engine(g, P):
    from picles_b200.engine1d import B200Engine1D
    return B200Engine1D(g["Nx"], g["xmin"], g["dx"], g["x"], P, device=0)


def _oracle(g, P):
    from oracle import oned
    return oned.Oracle1D(g["Nx"], g["xmin"], g["dx"], g["x"], P)


@pytest.mark.parametrize("name", sorted(SCENARIOS_1D))
def test_gpu_1d_matches_the_oracle(name):
    g, P, wind, DT, steps = SCENARIOS_1D[name]()
    e = _engine(g, P)
    run_pair_1d(_oracle(g, P), e, g, wind, DT, steps, compare_models_1d)
    c = e.counters()
    assert c["n_active"] == g["Nx"]
    if name.startswith("fast"):
        assert c["reach"] > 3   # the window of the gather, and its wrap on the periodic grid, were exercised
    e.close()


def test_gpu_1d_large_grid_window_gather():
    """20 001 nodes: the gather scans a window of 2*reach + 1 particles per node, not the whole grid"""
    from scenarios_1d import grid_1d, params_1d
    g = grid_1d(0.0, 2000e3, 20001)
    P = params_1d(600.0, periodic=True)
    wind = lambda x, t: 8.0 + 6.0 * np.sin(2 * np.pi * np.asarray(x) / 2000e3)
    e = _engine(g, P)
    run_pair_1d(_oracle(g, P), e, g, wind, 600.0, 4, compare_models_1d)
    e.close()


def test_gpu_1d_host_api():
    """WaveGrowth1D(; grid, winds, ODEsys, ODEsets, ...) / init_particles! / time_step! as the reference's 1-D scripts
    call them (tests/T03_PIC_propagation_1d.jl:100-125), against the oracle driven with the same staged winds"""
    from picles_b200 import FetchRelations as FR
    from picles_b200.Models.WaveGrowthModels1D import WaveGrowth1D, init_particles_1D, time_step_1D
    from picles_b200.ParticleMesh import OneDGrid
    from picles_b200.ParticleSystems import particle_waves_v5 as PW
    DT = 1200.0
    pars, cid, _ = PW.ODEParameters(r_g=0.85)
    u = lambda x, t: 10.0 + x * 0 + t * 0
    system = PW.particle_equations(u, γ=cid.γ, q=cid.q)
    sets = PW.ODESettings(Parameters=dict(r_g=0.85, C_α=pars["C_α"], C_e=pars["C_e"]), log_energy_minimum=FR.MinimalWindsea(10, 0, DT)["lne"],
                          log_energy_maximum=np.log(17), saving_step=10, timestep=DT, total_time=2 * 86400.0, adaptive=True,
                          dt=1e-3, dtmin=1e-12, force_dtmin=True, solver="Tsit5")
    grid = OneDGrid(1e3, 3e6 - 1e3, 40)
    model = WaveGrowth1D(grid=grid, winds=u, ODEsys=system, ODEsets=sets, ODEinit_type="wind_sea", periodic_boundary=False,
                         boundary_type="same")
    init_particles_1D(model)
    from oracle import oned
    from picles_b200.params import make_params
    P = make_params(sets, system, model.minimal_state, defaults=None, periodic_boundary=False)
    o = oned.Oracle1D(grid.Nx, grid.xmin, grid.dx, model.gridnotes.x, P)
    o.seed(np.full(grid.Nx, 10.0))
    assert np.array_equal(model.State, o.state().T)
    for k in range(5):
        t = model.clock.time
        time_step_1D(model, DT)
        o.step(t, DT, np.full(grid.Nx, 10.0), np.full(grid.Nx, 10.0))
        assert np.array_equal(model.State, o.state().T), f"step {k + 1}"
    assert model.clock.iteration == 5 and model.State.shape == (40, 3)
    assert model.boundary == [1, 40]


def test_gpu_1d_simulation_run_with_cash_store():
    """Simulation(model, Δt, stop_time) / run!(sim, cash_store=true) on the 1-D model (tests/T03_PIC_propagation_1d.jl:
    176-182): floor(stop/Δt) + 1 steps, one State copy per step plus the initial one"""
    from picles_b200 import FetchRelations as FR
    from picles_b200.Models.WaveGrowthModels1D import WaveGrowth1D
    from picles_b200.ParticleMesh import OneDGrid
    from picles_b200.ParticleSystems import particle_waves_v5 as PW
    from picles_b200.Simulations.run import Simulation, run
    DT = 600.0
    pars, cid, _ = PW.ODEParameters(r_g=0.85)
    u = lambda x, t: 12.0
    system = PW.particle_equations(u, γ=cid.γ, q=cid.q)
    sets = PW.ODESettings(Parameters=dict(r_g=0.85, C_α=pars["C_α"], C_e=pars["C_e"]), log_energy_minimum=FR.MinimalWindsea(10, 0, DT)["lne"],
                          saving_step=DT, timestep=DT, total_time=86400.0, dt=1e-3, dtmin=1e-4, force_dtmin=True, solver="Tsit5")
    model = WaveGrowth1D(grid=OneDGrid(0.0, 500e3, 51), winds=u, ODEsys=system, ODEsets=sets, periodic_boundary=True)
    sim = Simulation(model, Δt=DT, stop_time=3600.0, verbose=False)
    run(sim, cash_store=True)
    assert model.clock.iteration == 7 and len(sim.store.store) == 8
    assert all(s.shape == (51, 3) for s in sim.store.store)
    assert np.array_equal(sim.store.store[-1], model.State) and np.all(np.isfinite(model.State))
    assert model.counters()["n_integrated"] == 51 and model.counters()["n_failed"] == 0
