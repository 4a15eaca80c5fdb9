"""CPU tests: the product's device code (picles_b200/csrc/physics.h, compiled for the host
by tests/host_shim.cpp) must reproduce the oracle bit-for-bit — particles, State, counters —
including the gather's wrap/fold paths and the y-strip + halo decomposition."""
import numpy as np
import pytest

from common import HostShim, compare_models, make_oracle
from scenarios import SCENARIOS, run_pair


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_single_strip_bit_exact(name):
    g, P, wind, DT, n = SCENARIOS[name]()
    run_pair(make_oracle(g, P), HostShim(g, P), wind, DT, n, compare_models)


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_specialised_copies_bit_exact(name):
    """The copies the kernels pick per launch — the right-hand side without its term switches
    (every term on, n = 2), the Tsit5 instantiation with its compile-time tableau and no
    zero-coefficient tests — give the oracle's bits too (host build with the same choices)."""
    from common import shim_lib
    lib = shim_lib()
    lib.shim_set_specialised(1)
    try:
        assert lib.shim_get_specialised() == 1
        g, P, wind, DT, n = SCENARIOS[name]()
        run_pair(make_oracle(g, P), HostShim(g, P), wind, DT, n, compare_models)
    finally:
        lib.shim_set_specialised(0)


@pytest.mark.parametrize("name,nstrips,halo", [
    ("minimal", 2, 2), ("minimal", 3, 1), ("periodic_grid", 2, 5), ("periodic_grid", 4, 5),
    ("land_block", 3, 2), ("tripolar", 2, 8), ("tripolar", 3, 6), ("growing_winds", 2, 2),
    ("periodic_model_flag", 2, 2), ("dp5_blowup", 2, 2),
])
def test_strips_bit_exact(name, nstrips, halo):
    """N y-strips with a halo of particle records give the same bits as one strip."""
    g, P, wind, DT, n = SCENARIOS[name]()
    run_pair(make_oracle(g, P), HostShim(g, P, nstrips=nstrips, halo=halo), wind, DT, n, compare_models)


def test_scenarios_exercise_their_paths():
    """Guard against vacuous parity: the scenarios must actually hit wrap, fold, off
    particles, both remesh branches and reach >= 2."""
    seen = dict(reach2=False, D=False, B=False, reseed=False, rejects=False, fixups=False, failed=False)
    for name in ("periodic_grid", "growing_winds", "growing_winds_persist", "tripolar", "emax_clamp", "maxiters",
                 "dtmin_no_force", "nan_wind", "nan_defaults", "inf_defaults", "odd_periodic_strip"):
        g, P, wind, DT, n = SCENARIOS[name]()
        o = make_oracle(g, P)
        u0, v0 = wind(0.0)
        o.seed(u0, v0)
        t = 0.0
        for _ in range(n):
            o.step(t, DT, *wind(t), *wind(t + DT))
            t += DT
            c = o.counters()
            seen["reach2"] |= c["reach"] >= 2
            seen["D"] |= c["n_remesh_D"] > 0
            seen["B"] |= c["n_remesh_B"] > 0
            seen["reseed"] |= c["n_reseed_advance"] > 0
            seen["rejects"] |= c["n_rejects"] > 0
            seen["fixups"] |= c["n_fixups"] > 0
            seen["failed"] |= c["n_failed"] > 0
    assert all(seen.values()), seen


@pytest.mark.parametrize("name", ["minimal", "tripolar", "periodic_grid"])
def test_bare_time_step_accumulates_onto_state(name):
    """time_step! without the State .= 0 of run! adds deposits to the current node values
    (first step: on top of the seeded State), in the reference's order."""
    g, P, wind, DT, n = SCENARIOS[name]()
    o, s = make_oracle(g, P), HostShim(g, P)
    o.set_accumulate(True)
    s.set_accumulate(True)
    run_pair(o, s, wind, DT, 3, compare_models)
    # and it differs from the zero-first run after the first step
    o2 = make_oracle(g, P)
    u0, v0 = wind(0.0)
    o2.seed(u0, v0)
    o2.step(0.0, DT, *wind(0.0), *wind(DT))
    o3 = make_oracle(g, P)
    o3.set_accumulate(True)
    o3.seed(u0, v0)
    o3.step(0.0, DT, *wind(0.0), *wind(DT))
    assert not np.array_equal(o2.state(), o3.state())


@pytest.mark.parametrize("seed", range(10))
def test_random_configurations_bit_exact(seed):
    """the randomized configurations of tests/test_independent_model.py (grid, boundaries, land, model flag, solver,
    thresholds, winds varying in space and time, staged here as the two levels of each step) through the host build
    of the device header: one strip, the specialised copies, and two strips with a halo as deep as the reach — bit for
    bit against the oracle.  (150 seeds x the three modes were run once: no difference.)"""
    from common import shim_lib
    from test_independent_model import fuzz_case
    g, P, winds, DT, _ = fuzz_case(seed)

    def wind(t):
        return tuple(np.array([[winds(g["x"][j, i], g["y"][j, i], t)[k] for i in range(g["Nx"])] for j in range(g["Ny"])])
                     for k in (0, 1))

    ref = make_oracle(g, P)
    run_pair(ref, HostShim(g, P), wind, DT, 4, compare_models)
    reach = 1
    probe = make_oracle(g, P)
    probe.seed(*wind(0.0))
    for k in range(4):
        probe.step(k * DT, DT, *wind(k * DT), *wind((k + 1) * DT))
        reach = max(reach, probe.counters()["reach"])
    run_pair(make_oracle(g, P), HostShim(g, P, nstrips=2, halo=reach), wind, DT, 4, compare_models)
    lib = shim_lib()
    lib.shim_set_specialised(1)
    try:
        run_pair(make_oracle(g, P), HostShim(g, P), wind, DT, 4, compare_models)
    finally:
        lib.shim_set_specialised(0)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("mode", ["AutoTsit5", "on_persist", "AutoTsit5+on_persist"])
def test_random_configurations_default_solver_and_persisting_flags(seed, mode):
    """the same generator under the reference's default solver (45 of 120 seeds reach the Rosenbrock23 branch) and
    with `on` following the remesh (on_persist = 1): bit for bit (120 seeds x 3 modes were run once)"""
    from common import default_params
    from test_independent_model import fuzz_case
    g, P, winds, DT, solver = fuzz_case(seed)
    P2 = default_params(DT=DT, solver="AutoTsit5" if "AutoTsit5" in mode else solver, periodic_boundary=bool(P.periodic_boundary),
                        wind_min_squared=P.wind_min_squared, log_energy_maximum=P.log_energy_maximum,
                        on_persist="on_persist" in mode)

    def wind(t):
        return tuple(np.array([[winds(g["x"][j, i], g["y"][j, i], t)[k] for i in range(g["Nx"])] for j in range(g["Ny"])])
                     for k in (0, 1))

    run_pair(make_oracle(g, P2), HostShim(g, P2), wind, DT, 5, compare_models)


def _blowup_case(nan_eest_rejects):
    """scenario dp5_blowup (DP5 trial steps that overflow in a band of columns) with the reject switch"""
    from common import cartesian_grid, default_params
    g = cartesian_grid(24, 10)
    v = np.broadcast_to(np.linspace(13.6, 14.4, 24), (10, 24)).copy()
    u = np.zeros((10, 24))
    return g, default_params(solver="DP5", nan_eest_rejects=nan_eest_rejects), (lambda t: (u, v)), 600.0, 3


@pytest.mark.parametrize("mode", ["one strip", "specialised", "two strips"])
def test_nan_error_estimate_rejected_instead_of_ending_the_integrator(mode):
    """picles_params_t::nan_eest_rejects = 1 (the fastpow / fastpower reading, include/picles_b200.h): the particles that
    scenario dp5_blowup loses integrate on — nobody fails — bit for bit between oracle and device code"""
    from common import shim_lib
    g, P, wind, DT, n = _blowup_case(True)
    lib = shim_lib()
    lib.shim_set_specialised(1 if mode == "specialised" else 0)
    try:
        ref = make_oracle(g, P)
        kw = dict(nstrips=2, halo=2) if mode == "two strips" else {}
        run_pair(ref, HostShim(g, P, **kw), wind, DT, n, compare_models)
    finally:
        lib.shim_set_specialised(0)
    o = make_oracle(g, P)
    o.seed(*wind(0.0))
    failed = rejects = 0
    for k in range(n):
        o.step(k * DT, DT, *wind(k * DT), *wind((k + 1) * DT))
        failed += o.counters()["n_failed"]
        rejects += o.counters()["n_rejects"]
    assert failed == 0 and rejects >= 176 and not (o.particles()["status"] & 4).any()
    g0, P0, wind0, _, _ = _blowup_case(False)
    o0 = make_oracle(g0, P0)
    o0.seed(*wind0(0.0))
    o0.step(0.0, DT, *wind0(0.0), *wind0(DT))
    assert o0.counters()["n_failed"] == 32


@pytest.mark.parametrize("seed", [2, 5, 9, 13, 17, 23])
def test_random_configurations_with_the_reject_switch(seed):
    """random configurations (several of these seeds lose particles to overflowing trial steps with the switch off) with
    nan_eest_rejects = 1, under the configuration's solver: device code = oracle bit for bit, nobody ends with DtNaN"""
    from common import default_params
    from test_independent_model import fuzz_case
    g, P, winds, DT, solver = fuzz_case(seed)
    P2 = default_params(DT=DT, solver=solver, periodic_boundary=bool(P.periodic_boundary), wind_min_squared=P.wind_min_squared,
                        log_energy_maximum=P.log_energy_maximum, nan_eest_rejects=True)

    def wind(t):
        return tuple(np.array([[winds(g["x"][j, i], g["y"][j, i], t)[k] for i in range(g["Nx"])] for j in range(g["Ny"])])
                     for k in (0, 1))

    ref = make_oracle(g, P2)
    run_pair(ref, HostShim(g, P2), wind, DT, 4, compare_models)


@pytest.mark.parametrize("seed", range(6))
def test_random_tall_configurations_in_three_and_four_strips(seed):
    """the random generator on grids four times as tall, cut in 3 and 4 strips whose height covers the halo (a reach beyond
    a strip's own height is an error of the library, not a case): bit for bit (80 seeds x {3, 4} strips x two halo
    widths were run once)"""
    from test_independent_model import fuzz_case
    g, P, winds, DT, _ = fuzz_case(seed, ny_scale=4)

    def wind(t):
        return tuple(np.array([[winds(g["x"][j, i], g["y"][j, i], t)[k] for i in range(g["Nx"])] for j in range(g["Ny"])])
                     for k in (0, 1))

    probe = make_oracle(g, P)
    probe.seed(*wind(0.0))
    reach = 1
    for k in range(4):
        probe.step(k * DT, DT, *wind(k * DT), *wind((k + 1) * DT))
        reach = max(reach, probe.counters()["reach"])
    ran = 0
    for ns in (3, 4):
        if g["Ny"] // ns < reach + 2:
            continue
        run_pair(make_oracle(g, P), HostShim(g, P, nstrips=ns, halo=reach), wind, DT, 4, compare_models)
        ran += 1
    if not ran:
        pytest.skip("deposits reach further than a strip is tall")
