"""GPU tests (-m gpu): AutoTsit5(Rosenbrock23()) through the C ABI — the monitor-carrying
instantiation of k_advance and the out-of-line Rosenbrock23 branch — bit-exact against the oracle."""
import numpy as np
import pytest

from common import bits_equal, compare_models
from scenarios import SCENARIOS, run_pair
from test_autotsit5 import AUTOTSIT5, with_solver
from test_gpu_parity import StripSet, engine_for
from common import make_oracle

pytestmark = pytest.mark.gpu


def cmp_state(a, b):
    compare_models(a, b, check_aux=False)
    act = (a.particles()["flags"] & 8) != 0
    assert np.array_equal(a.solver_state()[act], b.solver_state()[act])


@pytest.mark.parametrize("name", ["growing_winds", "growing_winds_persist"])
def test_gpu_autotsit5_stiff_branch_matches_oracle(gpu_lib, name):
    g, P, wind, DT, nsteps = SCENARIOS[name]()
    P = with_solver(P, AUTOTSIT5)
    ref, dut = make_oracle(g, P), engine_for(g, P)
    tot = [0, 0]

    def cmp(a, b):
        cmp_state(a, b)
        c = b.counters()
        tot[0] += c["n_stiff_switches"]
        tot[1] += c["n_stiff_attempts"]

    run_pair(ref, dut, wind, DT, nsteps, cmp)
    assert tot[0] > 500 and tot[1] > 5000


@pytest.mark.parametrize("name", ["minimal", "tripolar", "periodic_grid"])
def test_gpu_autotsit5_equals_tsit5_without_switches(gpu_lib, name):
    g, P, wind, DT, nsteps = SCENARIOS[name]()
    ref = make_oracle(g, with_solver(P, 0))
    dut = engine_for(g, with_solver(P, AUTOTSIT5))
    run_pair(ref, dut, wind, DT, nsteps, lambda a, b: compare_models(a, b, check_aux=False))
    c = dut.counters()
    assert c["n_stiff_switches"] == 0 and c["n_stiff_attempts"] == 0


def test_gpu_autotsit5_checkpoint_carries_the_switch_state(gpu_lib):
    """save after the first step of the growing-wind run (hundreds of particles under
    Rosenbrock23), restore into a fresh handle, continue: bit-identical to the uninterrupted run"""
    g, P, wind, DT, nsteps = SCENARIOS["growing_winds"]()
    P = with_solver(P, AUTOTSIT5)
    a, b = engine_for(g, P), engine_for(g, P)
    a.seed(*wind(0.0))
    t = 0.0
    for _ in range(1):
        a.step(t, DT, *wind(t), *wind(t + DT))
        t += DT
    assert (a.solver_state() > 60).sum() > 100
    b.restore(a.checkpoint())
    for _ in range(4):
        for e in (a, b):
            e.step(t, DT, *wind(t), *wind(t + DT))
        t += DT
    assert bits_equal(a.state(), b.state())
    assert np.array_equal(a.solver_state(), b.solver_state())


@pytest.mark.parametrize("name", ["nan_wind", "maxiters", "dtmin_no_force", "emax_clamp", "calm", "land_block", "fast_box"])
def test_gpu_edge_scenarios_under_autotsit5(gpu_lib, name):
    g, P, wind, DT, nsteps = SCENARIOS[name]()
    P = with_solver(P, AUTOTSIT5)
    run_pair(make_oracle(g, P), engine_for(g, P), wind, DT, nsteps, cmp_state)


def test_gpu_autotsit5_in_strips(gpu_lib):
    """three strip handles on one GPU (host-driven exchange): parked particles are resumed per
    strip, fields equal the single-domain oracle"""
    g, P, wind, DT, nsteps = SCENARIOS["growing_winds"]()
    P = with_solver(P, AUTOTSIT5)
    ref, dut = make_oracle(g, P), StripSet(g, P, 3, 2)
    run_pair(ref, dut, wind, DT, nsteps, lambda a, b: compare_models(a, b, check_aux=False))
