"""AutoTsit5(Rosenbrock23()) — the ODESettings default solver (particle_waves_v5.jl:47) — on the CPU:
the oracle's restatement of OrdinaryDiffEq's AutoSwitch + Rosenbrock23 and the device code (host
build of physics.h + stiff.h), bit for bit, plus independent checks of the pieces the reference
does not pin: the dual-number Jacobian against finite differences and the stiff integrator
against scipy's Radau."""
import copy

import numpy as np
import pytest

import oracle
from common import HostShim, bits_equal, compare_models, default_params, make_oracle
from scenarios import SCENARIOS, run_pair

AUTOTSIT5 = 2
HALO = {"tripolar": 6, "periodic_grid": 5, "fast_box": 5}


def with_solver(P, solver):
    P2 = copy.copy(P)
    P2.solver = solver
    return P2


def compare_with_solver_state(a, b):
    compare_models(a, b)
    act = (a.particles()["flags"] & 8) != 0
    assert np.array_equal(a.solver_state()[act], b.solver_state()[act])


@pytest.mark.parametrize("name", ["growing_winds", "growing_winds_persist"])
@pytest.mark.parametrize("nstrips", [1, 2])
def test_device_code_matches_oracle_through_the_stiff_branch(name, nstrips):
    """the growing/decaying-wind scenario (C3): ~10^3 particles at the calm foot of the ramp are
    handed to Rosenbrock23 and back; State, particles, AutoSwitch state and counters agree"""
    g, P, wind, DT, nsteps = SCENARIOS[name]()
    P = with_solver(P, AUTOTSIT5)
    ref, dut = make_oracle(g, P), HostShim(g, P, nstrips=nstrips, halo=2)
    tot = [0, 0]

    def cmp(a, b):
        compare_with_solver_state(a, b)
        c = a.counters()
        tot[0] += c["n_stiff_switches"]
        tot[1] += c["n_stiff_attempts"]

    run_pair(ref, dut, wind, DT, nsteps, cmp)
    assert tot[0] > 500 and tot[1] > 5000          # the branch really ran
    assert (ref.solver_state() > 60).any()         # some particles end the run under Rosenbrock23


@pytest.mark.parametrize("name", ["minimal", "periodic_grid", "tripolar", "land_block", "fast_box"])
def test_autotsit5_is_tsit5_where_the_monitor_never_fires(name):
    """homogeneous box, periodic, land and tripolar scenarios: no switch, and the AutoTsit5 run is
    the Tsit5 run bit for bit (oracle and device code)"""
    g, P, wind, DT, nsteps = SCENARIOS[name]()
    ref = make_oracle(g, with_solver(P, 0))
    dut = HostShim(g, with_solver(P, AUTOTSIT5), nstrips=2, halo=HALO.get(name, 2))
    switches = [0]

    def cmp(a, b):
        compare_models(a, b)
        switches[0] += b.counters()["n_stiff_switches"] + b.counters()["n_stiff_attempts"]

    run_pair(ref, dut, wind, DT, nsteps, cmp)
    assert switches[0] == 0


def test_stiff_branch_stays_within_tolerance_of_tsit5():
    g, P, wind, DT, nsteps = SCENARIOS["growing_winds"]()
    S = {}
    for solver in (0, AUTOTSIT5):
        o = make_oracle(g, with_solver(P, solver))
        o.seed(*wind(0.0))
        t = 0.0
        for _ in range(nsteps):
            o.step(t, DT, *wind(t), *wind(t + DT))
            t += DT
        S[solver] = o.state()
    rel = np.nanmax(np.abs(S[0] - S[AUTOTSIT5]) / np.maximum(np.abs(S[0]), 1e-30))
    assert 0 < rel < 5e-3          # different integrators, reltol 1e-3


def test_dual_number_jacobian_against_finite_differences():
    P = default_params(DT=1200.0)
    M = (1 / 4000.0, 1e-5, -2e-5, 1 / 3000.0)
    rng = np.random.default_rng(0)
    for _ in range(20):
        z = np.array([rng.uniform(-12, 2), rng.uniform(-8, 8), rng.uniform(-8, 8), 0.0, 0.0])
        u, v = rng.uniform(-15, 15, 2)
        J, dT = oracle.rhs_jacobian(P, z, u, v, 0.01, -0.02, M=M, pc=1e-7)
        Jfd = np.zeros((5, 5))
        for j in range(5):
            h = 1e-6 * max(1.0, abs(z[j]))
            zp, zm = z.copy(), z.copy()
            zp[j] += h
            zm[j] -= h
            Jfd[:, j] = (oracle.rhs(P, zp, u, v, M=M, pc=1e-7) - oracle.rhs(P, zm, u, v, M=M, pc=1e-7)) / (2 * h)
        scale = np.maximum(np.abs(Jfd), 1e-6 * np.abs(Jfd).max())   # central differences: ~1e-9 relative to the largest entry
        assert np.max(np.abs(J - Jfd) / scale) < 1e-4
        h = 1e-5
        dTfd = (oracle.rhs(P, z, u + 0.01 * h, v - 0.02 * h, M=M, pc=1e-7)
                - oracle.rhs(P, z, u - 0.01 * h, v + 0.02 * h, M=M, pc=1e-7)) / (2 * h)
        assert np.max(np.abs(dT - dTfd)) < 1e-6 * max(1.0, np.abs(dTfd).max())
    assert np.all(J[:, 3:] == 0.0)      # the system does not depend on the particle position


def test_rosenbrock23_converges_to_an_independent_stiff_solver():
    """a particle forced to start under Rosenbrock23 (as_state = (0, 1)): as the tolerances tighten
    the result converges to scipy's Radau at rtol 1e-11"""
    from scipy.integrate import solve_ivp
    P = default_params(DT=1200.0, wind_min_squared=2.0)
    M = (1 / 4000.0, 0.0, 0.0, 1 / 4000.0)
    z0 = np.array([-2.5, 3.0, 1.0, 0.0, 0.0])
    wind = (9.0, 4.0)
    ref = solve_ivp(lambda t, y: oracle.rhs(P, y, wind[0], wind[1], M=M), (0, 1200.0), z0, method="Radau", rtol=1e-11,
                    atol=1e-13).y[:, -1]
    errs = []
    for tol in (1e-3, 1e-5, 1e-7):
        P2 = with_solver(P, AUTOTSIT5)
        P2.reltol, P2.abstol = tol, tol * 0.1
        r = oracle.integrate_one(P2, z0, dt_reset=True, wind0=wind, DT=1200.0, M=M, as_state=(0, 1))
        assert r["counters"]["n_stiff_attempts"] >= 4
        errs.append(np.max(np.abs(r["u"] - ref) / np.maximum(np.abs(ref), 1e-3)))
    assert errs[0] < 5e-3 and errs[2] < errs[0] / 50 and errs[2] < 2e-5


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_every_scenario_under_autotsit5(name):
    """all parity scenarios (edge cases included: NaN winds, maxiters, dtmin, all land, ...) with the
    solver id switched to AutoTsit5: device code == oracle"""
    g, P, wind, DT, nsteps = SCENARIOS[name]()
    P = with_solver(P, AUTOTSIT5)
    run_pair(make_oracle(g, P), HostShim(g, P), wind, DT, nsteps, compare_with_solver_state)
