"""BASELINE.json configs at (or near) their full sizes on the GPU, checked through
size-independent properties and against the threaded oracle where it finishes in seconds."""
import numpy as np
import pytest

from common import BND_NONPERIODIC, cartesian_grid, compare_models, default_params, make_oracle, tripolar_grid
from scenarios import run_pair

pytestmark = pytest.mark.gpu


def engine_for(grid, P, **kw):
    from picles_b200.engine import B200Engine
    return B200Engine(grid["Nx"], grid["Ny"], grid["bx"], grid["by"], grid["mask"], P, M=grid["M"],
                      M_const=grid["M_const"], pc=grid["pc"], **kw)


def growing_wind(x, U10=10.0, V10=3.0):
    """T04_2D_growing_decaying_winds.jl:126-132 (x-dependent ramp), modulated in time"""
    Lx = x.max()
    x0 = 50.0 / 260.0 * Lx
    ramp = np.where(x < x0, 0.1 / abs(U10), (x - x0) / (Lx - x0))

    def wind(t):
        f = 0.6 + 0.4 * np.sin(2 * np.pi * t / 7200.0)
        return U10 * ramp * f, V10 * ramp * f + 0.05
    return wind


@pytest.mark.parametrize("on_persist", [False, True])
def test_c3_growing_decaying_winds_2048(gpu_lib, on_persist):
    """configs[2]: 2048x2048 Cartesian, dx=dy=4 km, DT=20 min, time-varying wind that depends
    on x only, on/off thresholds (wind_min_squared=2).  The problem is translation-invariant
    in y away from the y-edges, so every interior row must carry the bits of the middle row of
    a 2048 x 24 box integrated by the CPU oracle."""
    N = 2048
    P = default_params(DT=1200.0, wind_min_squared=2.0, on_persist=on_persist)
    g = cartesian_grid(N, N, dx=4000.0, dy=4000.0)
    gs = cartesian_grid(N, 24, dx=4000.0, dy=4000.0)
    wind = growing_wind(g["x"][0])
    e = engine_for(g, P)
    o = make_oracle(gs, P, variant="omp", threads=8)
    u0, v0 = wind(0.0)
    e.seed(u0, v0)
    o.seed(u0, v0)
    t = 0.0
    for _ in range(4):
        w = [*wind(t), *wind(t + 1200.0)]
        e.step(t, 1200.0, *w)
        o.step(t, 1200.0, *w)
        t += 1200.0
    S, So = e.state(), o.state()
    c = e.counters()
    assert c["n_failed"] == 0 and c["n_active"] == (N - 2) ** 2
    assert c["n_remesh_D"] > 0 and c["n_remesh_A"] > 0        # both on and off particles exist
    ref = So[:, 12, :]                                        # middle row of the narrow box
    for j in (100, 1024, 1900):
        assert np.array_equal(S[:, j, :].view(np.uint64), ref.view(np.uint64)), f"row {j} differs from the oracle row"
    blk = S[:, 64:1984, :]
    assert np.array_equal(blk.view(np.uint64), np.broadcast_to(ref[:, None, :], blk.shape).copy().view(np.uint64))


def _tripolar_case(Nx, Ny, land):
    ocean = np.ones((Ny, Nx), np.uint8)
    ocean[: max(2, Ny // 40), :] = 0                           # south cap (TripolarGrid_mask_pols!)
    if land:
        yy, xx = np.mgrid[0:Ny, 0:Nx]
        for cx, cy, r in ((0.2, 0.45, 0.08), (0.55, 0.6, 0.1), (0.8, 0.3, 0.06), (0.5, 0.97, 0.04)):
            ocean[((xx - cx * Nx) / Nx) ** 2 + ((yy - cy * Ny) / Ny) ** 2 < r * r] = 0
        ocean[int(0.7 * Ny):int(0.75 * Ny), int(0.1 * Nx):int(0.3 * Nx)] = 0
    g = tripolar_grid(Nx, Ny, ocean=ocean)
    g["M"] = g["M"] * 1.2   # cells shrink to the 2 km floor near the pole: reach 2-3 there, a fraction of a cell at the equator
    return g


@pytest.mark.parametrize("land", [False, True])
def test_c4_c5_tripolar_at_scale_against_threaded_oracle(gpu_lib, land):
    """configs[3]/[4]: synthetic tripolar grid (periodic x, tripolar-north fold, per-node rotated
    kernel, great-circle term), aqua planet and with a synthetic land mask, 720 x 540 = 389k
    nodes, winds of T03_PIC_tripolar_aqua.jl:67-68: bit-exact against the oracle."""
    g = _tripolar_case(720, 540, land)
    P = default_params(DT=1200.0, periodic_boundary=True)
    wind = lambda t: (15.0, -10.0 * np.cos(5 * t / (3600 * 2 * np.pi)))
    o = make_oracle(g, P, variant="omp", threads=8)
    e = engine_for(g, P)
    run_pair(o, e, wind, 1200.0, 3, compare_models, every=3)
    c = e.counters()
    assert c["reach"] >= 1 and c["n_failed"] == 0
    if land:
        assert (g["mask"] == 2).sum() > 0


def test_c5_tripolar_high_resolution_strips_properties(gpu_lib):
    """configs[4] at high resolution (2880 x 2160 = 6.2 M nodes) in 4 y-strips on one GPU
    (the decomposition the 8-GPU run uses) against the same grid as a single strip: identical
    bits — results do not depend on the strip count."""
    from test_gpu_parity import StripSet
    g = _tripolar_case(2880, 2160, True)
    P = default_params(DT=1200.0, periodic_boundary=True)
    wind = lambda t: (15.0, -10.0 * np.cos(5 * t / (3600 * 2 * np.pi)))
    one = engine_for(g, P)
    four = StripSet(g, P, 4, 5)
    u0, v0 = wind(0.0)
    one.seed(u0, v0)
    four.seed(u0, v0)
    t = 0.0
    for _ in range(2):
        w = [*wind(t), *wind(t + 1200.0)]
        one.step(t, 1200.0, *w)
        four.step(t, 1200.0, *w)
        t += 1200.0
    assert np.array_equal(one.state().view(np.uint64), four.state().view(np.uint64))
    pa, pb = one.particles(), four.particles()
    act = (pa["flags"] & 8) != 0
    assert np.array_equal(pa["z"][:, act].view(np.uint64), pb["z"][:, act].view(np.uint64))
    assert one.counters()["n_substeps"] == four.counters()["n_substeps"]


def test_wind_mesh_sampler_full_size(gpu_lib):
    """k_wind_sample over the 4096 x 4096 bench grid from an ERA5-like 1-degree mesh (361 x 181 knots,
    6-hourly), nodes spilling over the mesh on every side (periodic wrap), against the oracle's
    restatement of LinearInterpolation(..., extrapolation_bc=Periodic()) — bit for bit on all 16.8 M nodes —
    and the AutoTsit5 C3 run (stiff branch active at full size) against the translation-invariance property"""
    import oracle
    from picles_b200.engine import B200Engine
    N = 4096
    rng = np.random.default_rng(4)
    xw = np.linspace(5.0e5, 7.5e6, 361)
    yw = np.sort(rng.uniform(3.0e5, 7.9e6, 181))          # non-uniform in y
    tw = np.arange(5) * 21600.0
    U = rng.normal(8.0, 4.0, (tw.size, yw.size, xw.size))
    V = rng.normal(-3.0, 5.0, (tw.size, yw.size, xw.size))
    x = np.ascontiguousarray(np.broadcast_to(np.arange(N) * 2000.0, (N, N)))
    y = np.ascontiguousarray(np.broadcast_to((np.arange(N) * 2000.0)[:, None], (N, N)))
    e = B200Engine(N, N, 0, 0, np.ones((N, N), np.uint8), default_params(), M_const=np.array([5e-4, 0.0, 0.0, 5e-4]))
    e.set_wind_mesh(xw, yw, tw, U, V, x, y)
    for t in (12345.0, 86400.0 + 7.0):
        ud, vd = e.sample_wind_mesh(t)
        uo, vo = oracle.wind_mesh_sample(xw, yw, tw, U, V, x, y, t)
        assert np.array_equal(ud.view(np.uint64), uo.view(np.uint64))
        assert np.array_equal(vd.view(np.uint64), vo.view(np.uint64))


def test_c3_autotsit5_stiff_branch_at_full_size(gpu_lib):
    """configs[2] under ODESettings' default solver: thousands of particles go through Rosenbrock23
    (k_advance_resume) per step; every interior row still carries the bits of the middle row of a
    2048 x 24 box integrated by the CPU oracle with the same solver"""
    import copy
    N = 2048
    P = copy.copy(default_params(DT=1200.0, wind_min_squared=2.0))
    P.solver = 2
    g = cartesian_grid(N, N, dx=4000.0, dy=4000.0)
    gs = cartesian_grid(N, 24, dx=4000.0, dy=4000.0)
    wind = growing_wind(g["x"][0])
    winds = lambda t: [np.broadcast_to(a, (24, N)) for a in wind(t)]
    eng = engine_for(g, P)
    ref = make_oracle(gs, P, variant="omp", threads=8)
    eng.seed(*wind(0.0))
    ref.seed(*winds(0.0))
    t, switches = 0.0, 0
    for _ in range(4):
        eng.step(t, 1200.0, *wind(t), *wind(t + 1200.0))
        ref.step(t, 1200.0, *winds(t), *winds(t + 1200.0))
        t += 1200.0
        switches += eng.counters()["n_stiff_switches"]
    assert switches > 1000
    S, Sr = eng.state(), ref.state()
    mid = Sr[:, 12, :]
    for j in (8, 500, 1024, 2039):
        assert np.array_equal(S[:, j, :].view(np.uint64), mid.view(np.uint64)), j
    st = eng.solver_state()
    assert np.array_equal(st[1024], ref.solver_state()[12])
