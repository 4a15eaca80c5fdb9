"""GPU tests (-m gpu) of the wind-ingestion row, through the C ABI:
intermediate wind levels (picles_set_wind_midlevels) and the device-resident wind mesh
(picles_set_wind_mesh / picles_sample_wind_mesh / picles_seed_wind_mesh / picles_step_wind_mesh),
bit-exact against the CPU oracle."""
import math

import numpy as np
import pytest

import oracle
from common import bits_equal, cartesian_grid, compare_models, default_params, make_oracle, tripolar_grid
from test_gpu_parity import StripSet, engine_for
from test_wind_levels import aqua_wind, mid_times, run_levels, synthetic_mesh, wind_arrays

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_mid", [1, 2, 3])
def test_gpu_midlevels_match_oracle(gpu_lib, n_mid):
    g = cartesian_grid(70, 45)
    P = default_params(DT=1200.0)
    ref, dut = make_oracle(g, P), engine_for(g, P)
    run_levels(ref, g, 1200.0, 4, n_mid)
    run_levels(dut, g, 1200.0, 4, n_mid)
    compare_models(ref, dut, check_aux=False)


def test_gpu_midlevels_on_a_per_node_metric_grid(gpu_lib):
    """tripolar kernel (k_advance<true>) with spatially varying, time-varying winds and 2
    intermediate levels"""
    g = tripolar_grid(48, 36)
    P = default_params(DT=1200.0)
    w = 5.0 / (3600.0 * 2.0 * math.pi)

    def wind(t):
        lat = g["y"]
        return 12.0 * np.cos(np.deg2rad(lat)) + 2.0 * np.sin(0.7 * w * t), -8.0 * np.cos(w * t) * np.ones_like(lat)

    ref, dut = make_oracle(g, P), engine_for(g, P)
    for m in (ref, dut):
        m.seed(*wind(0.0))
        t = 0.0
        for _ in range(4):
            lv = [wind(tm) for tm in mid_times(t, 1200.0, 2)]
            m.set_wind_midlevels([a for a, _ in lv], [b for _, b in lv])
            m.step(t, 1200.0, *wind(t), *wind(t + 1200.0))
            t += 1200.0
    compare_models(ref, dut, check_aux=False)


def test_gpu_midlevels_are_consumed_and_can_be_cleared(gpu_lib):
    g = cartesian_grid(33, 20)
    P = default_params(DT=1200.0)
    a, b = engine_for(g, P), engine_for(g, P)
    for e in (a, b):
        e.seed(*wind_arrays(g, 0.0))
    lv = [wind_arrays(g, tm) for tm in mid_times(0.0, 1200.0, 1)]
    a.set_wind_midlevels([x for x, _ in lv], [y for _, y in lv])
    a.set_wind_midlevels()                      # cleared again: a == b
    a.step(0.0, 1200.0, *wind_arrays(g, 0.0), *wind_arrays(g, 1200.0))
    b.step(0.0, 1200.0, *wind_arrays(g, 0.0), *wind_arrays(g, 1200.0))
    assert bits_equal(a.state(), b.state())
    b.set_wind_midlevels([x for x, _ in lv], [y for _, y in lv])
    b.step(1200.0, 1200.0, *wind_arrays(g, 1200.0), *wind_arrays(g, 2400.0))
    a.step(1200.0, 1200.0, *wind_arrays(g, 1200.0), *wind_arrays(g, 2400.0))
    assert not bits_equal(a.state(), b.state())  # the levels were used ...
    a.step(2400.0, 1200.0, *wind_arrays(g, 2400.0), *wind_arrays(g, 3600.0))
    ref = make_oracle(g, P)
    ref.seed(*wind_arrays(g, 0.0))
    t = 0.0
    for _ in range(3):                           # ... and a, never given any, equals the plain oracle
        ref.step(t, 1200.0, *wind_arrays(g, t), *wind_arrays(g, t + 1200.0))
        t += 1200.0
    compare_models(ref, a, check_aux=False)


def test_gpu_wind_mesh_sampler_matches_oracle(gpu_lib):
    xw, yw, tw, U, V = synthetic_mesh()
    g = cartesian_grid(257, 131, dx=137.0, dy=311.0)   # nodes spill over the mesh on both axes: periodic wrap
    x, y = g["x"] - 7000.0, g["y"] - 9000.0
    e = engine_for(g, default_params())
    e.set_wind_mesh(xw, yw, tw, U, V, x, y)
    for t in (0.0, 100.0, 21600.0, 50000.5, 86400.0, 90000.0, -500.0, 3 * 86400.0 + 17.0):
        uo, vo = oracle.wind_mesh_sample(xw, yw, tw, U, V, x, y, t)
        ud, vd = e.sample_wind_mesh(t)
        assert bits_equal(uo, ud) and bits_equal(vo, vd), t


@pytest.mark.parametrize("n_mid", [0, 2])
def test_gpu_step_wind_mesh_matches_oracle(gpu_lib, n_mid):
    """every wind level sampled on the device; the oracle is driven with its own restatement of
    the interpolation.  The time knots are not multiples of DT, so kinks fall inside steps."""
    rng = np.random.default_rng(11)
    g = cartesian_grid(60, 44)
    xw = np.linspace(-4000.0, 125000.0, 9)
    yw = np.linspace(-4000.0, 95000.0, 7)
    tw = np.array([0.0, 1000.0, 2500.0, 4000.0, 7000.0])
    U = 9.0 + 4.0 * rng.random((tw.size, yw.size, xw.size))
    V = 5.0 + 4.0 * rng.random((tw.size, yw.size, xw.size))
    P = default_params(DT=600.0)
    ref, dut = make_oracle(g, P), engine_for(g, P)
    dut.set_wind_mesh(xw, yw, tw, U, V, g["x"], g["y"])
    samp = lambda t: oracle.wind_mesh_sample(xw, yw, tw, U, V, g["x"], g["y"], t)
    ref.seed(*samp(0.0))
    dut.seed_wind_mesh(0.0)
    compare_models(ref, dut, check_aux=False)
    t = 0.0
    for _ in range(6):
        if n_mid:
            lv = [samp(tm) for tm in mid_times(t, 600.0, n_mid)]
            ref.set_wind_midlevels([a for a, _ in lv], [b for _, b in lv])
        ref.step(t, 600.0, *samp(t), *samp(t + 600.0))
        dut.step_wind_mesh(t, 600.0, n_mid)
        t += 600.0
    compare_models(ref, dut, check_aux=False)
    # a step at a time that does not continue the previous one re-samples the t level
    ref.step(5000.0, 600.0, *samp(5000.0), *samp(5600.0))
    dut.step_wind_mesh(5000.0, 600.0, 0)
    assert bits_equal(ref.state(), dut.state())


def test_gpu_wind_mesh_errors(gpu_lib):
    from picles_b200._abi import PiclesError
    g = cartesian_grid(16, 12)
    e = engine_for(g, default_params())
    e.seed(10.0, 10.0)
    with pytest.raises(PiclesError, match="set_wind_mesh"):
        e.step_wind_mesh(0.0, 600.0)
    xw, yw, tw, U, V = synthetic_mesh()
    with pytest.raises(PiclesError, match="strictly increasing"):
        e.set_wind_mesh(xw[::-1].copy(), yw, tw, U, V, g["x"], g["y"])
    with pytest.raises(PiclesError, match="n_mid"):
        e.set_wind_midlevels([g["x"]] * 4, [g["x"]] * 4)


def test_gpu_stage_wind_mesh_with_phase_split_strips(gpu_lib):
    """three strip handles on one GPU, host-driven exchange: every strip stages its own wind levels
    from its resident mesh (picles_stage_wind_mesh), then the phase-split calls run as usual"""
    from dist_worker import mesh_of
    from scenarios import SCENARIOS
    g, P, _, DT, nsteps = SCENARIOS["minimal"]()
    mesh = mesh_of(g, DT, nsteps)
    samp = lambda t: oracle.wind_mesh_sample(*mesh, g["x"], g["y"], t)
    ref = make_oracle(g, P)
    dut = StripSet(g, P, 3, 2)
    for (a, b), e in zip(dut.bounds, dut.e):
        e.set_wind_mesh(*mesh, g["x"][a:b], g["y"][a:b])
        e.seed_wind_mesh(0.0)
    ref.seed(*samp(0.0))
    t = 0.0
    for _ in range(nsteps):
        mids = [samp(t + DT * float(k) / 3.0) for k in (1, 2)]
        ref.set_wind_midlevels([m[0] for m in mids], [m[1] for m in mids])
        ref.step(t, DT, *samp(t), *samp(t + DT))
        for e in dut.e:                      # StripSet.step with staged instead of uploaded winds
            e.stage_wind_mesh(t, DT, 2)
            e.step_advance(t, DT)
            e.halo_pack()
        for e in dut.e:
            e.synchronize()
        bufs = [e.halo_buffers() for e in dut.e]
        for r, e in enumerate(dut.e):
            (slo, shi, rlo, rhi), nb = bufs[r]
            if r - 1 >= 0:
                e.copy_dev(rlo, bufs[r - 1][0][1], nb)
            if r + 1 < dut.ns:
                e.copy_dev(rhi, bufs[r + 1][0][0], nb)
            e.halo_unpack()
        for e in dut.e:
            e.step_project_remesh(t, DT)
        t += DT
        compare_models(ref, dut, check_aux=False)
