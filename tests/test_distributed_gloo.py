"""world_size>1 on CPU (gloo): the strip decomposition, neighbour logic and halo exchange of
picles_b200.distributed, with the host build of the device code as each rank's strip,
bit-exact against the single-domain oracle."""
import os
import socket

import pytest

from picles_b200.distributed import neighbours, row_cost_model, strip_bounds, strip_bounds_weighted


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("name,world,halo", [("minimal", 2, 2), ("periodic_grid", 2, 5), ("tripolar", 3, 6),
                                             ("growing_winds", 2, 2)])
def test_strips_over_gloo_match_oracle(tmp_path, name, world, halo):
    import torch.multiprocessing as mp

    import dist_worker
    out = tmp_path / "result"
    mp.spawn(dist_worker.run, args=(world, _free_port(), name, halo, str(out)), nprocs=world, join=True)
    assert out.read_text() == "ok"


@pytest.mark.parametrize("name,world,halo,wind_mode,n_mid", [("minimal", 2, 2, "mesh", 0), ("tripolar", 2, 6, "mesh", 2),
                                                             ("growing_winds", 2, 2, "host", 3)])
def test_strips_with_ingested_winds_over_gloo(tmp_path, name, world, halo, wind_mode, n_mid):
    """wind ingestion on strips: each rank samples its own rows from the wind mesh (mesh) or
    stages intermediate levels of its rows (host); the single-domain oracle gets the oracle's
    restatement of the sampler"""
    import torch.multiprocessing as mp

    import dist_worker
    out = tmp_path / "result"
    mp.spawn(dist_worker.run, args=(world, _free_port(), name, halo, str(out), "gloo", "torch-p2p", wind_mode, n_mid),
             nprocs=world, join=True)
    assert out.read_text() == "ok"


def test_strip_bounds_and_neighbours():
    assert strip_bounds(10, 3) == [(0, 3), (3, 6), (6, 10)]
    assert [b - a for a, b in strip_bounds(4096 * 8, 8)] == [4096] * 8
    assert neighbours(0, 1, True) == (-1, -1)
    assert neighbours(0, 4, False) == (-1, 1) and neighbours(3, 4, False) == (2, -1)
    assert neighbours(0, 4, True) == (3, 1) and neighbours(3, 4, True) == (2, 0)
    assert neighbours(0, 2, True) == (1, 1)  # two-strip ring: both neighbours are the other rank


def test_weighted_strip_bounds():
    import numpy as np
    # equal costs: the plain split
    assert strip_bounds_weighted(np.ones(12), 3) == [(0, 4), (4, 8), (8, 12)]
    # a land slab in the first third moves the cuts towards the ocean rows
    cost = np.r_[np.zeros(40), np.ones(80)]
    b = strip_bounds_weighted(cost, 4, min_rows=2)
    assert b[0][0] == 0 and b[-1][1] == 120 and all(x[1] == y[0] for x, y in zip(b, b[1:]))
    sums = [cost[a:c].sum() for a, c in b]
    assert max(sums) - min(sums) <= 1.0
    # min_rows is honoured even when the cost is concentrated
    b = strip_bounds_weighted(np.r_[np.zeros(30), [100.0], np.zeros(5)], 4, min_rows=3)
    assert all(c - a >= 3 for a, c in b)
    with pytest.raises(ValueError):
        strip_bounds_weighted(np.ones(5), 3, min_rows=2)
    mask = np.ones((6, 10), np.uint8)
    mask[:2] = 0
    mask[:, 0] = 3
    c = row_cost_model(mask, periodic_boundary=False, reach_rows=[1, 1, 1, 1, 2, 3])
    # land row: the gather of its 10 nodes only (0.068 each at reach 1); ocean row: + 9 active particles;
    # reach 2 and 3: the window grows as (2R+1)^2
    assert c[0] == pytest.approx(0.68) and c[2] == pytest.approx(9.68)
    assert c[4] - 9 == pytest.approx(0.68 * 25 / 9) and c[5] - 9 == pytest.approx(0.68 * 49 / 9)
    # measured costs: each strip's kernel times spread over its rows
    from picles_b200.distributed import row_cost_measured
    active = (mask == 1).sum(axis=1)
    m = row_cost_measured(active, 10, [1, 1, 1, 1, 2, 3], [(0, 3), (3, 6)], ms_advance=[1.0, 3.0], ms_gather=[0.3, 0.9])
    assert m[:3].sum() == pytest.approx(1.3) and m[3:].sum() == pytest.approx(3.9)
    assert m[0] == m[1] == pytest.approx(0.1) and m[2] == pytest.approx(1.1) and m[5] > m[4] > m[3]


def test_overlapped_step_partition():
    """picles_step_strip's overlapped step costs max_r(advance_r) + the strip's own gather, so the cut
    evens out the advance unless that piles gather on one strip.  Checked on the numbers of profiles/r02_bench_n8.json:
    the estimate reproduces every strip's measured step up to one constant."""
    import json
    import numpy as np
    from picles_b200.distributed import overlapped_step_estimate, row_cost_measured, strip_bounds_overlapped
    rng = np.random.default_rng(0)
    Ny, world = 400, 8
    adv = 1.0 + 0.3 * np.sin(np.arange(Ny) / 40.0) + 0.05 * rng.random(Ny)
    adv[120:160] = 0.1                                     # a land belt
    gat = np.full(Ny, 0.08)
    gat[-30:] = 0.6                                        # fold rows near the pole: an expensive gather
    b, lam, est = strip_bounds_overlapped(adv, gat, world, min_rows=6)
    assert b[0][0] == 0 and b[-1][1] == Ny and all(x[1] == y[0] for x, y in zip(b, b[1:])) and all(c - a >= 6 for a, c in b)
    b_sum = strip_bounds_weighted(adv + gat, world, min_rows=6)
    assert est == pytest.approx(overlapped_step_estimate(adv, gat, b))
    assert est < overlapped_step_estimate(adv, gat, b_sum)            # better than evening out the summed cost ...
    A = [adv[a:c].sum() for a, c in b]
    assert lam == 0.0 and max(A) - min(A) <= 2 * adv.max()            # ... by evening out the advance
    # a gather so heavy that evening out the advance alone is not the best cut any more
    gat2 = gat.copy()
    gat2[-30:] = 6.0
    assert strip_bounds_overlapped(adv, gat2, world, min_rows=6)[1] > 0.0
    assert strip_bounds_overlapped(adv, gat, 1)[0] == [(0, Ny)]
    # split=True returns the two parts of row_cost_measured
    act, reach = np.full(12, 10.0), np.ones(12)
    a2, g2 = row_cost_measured(act, 10, reach, [(0, 6), (6, 12)], [3.0, 6.0], [0.6, 1.2], split=True)
    assert np.allclose(a2 + g2, row_cost_measured(act, 10, reach, [(0, 6), (6, 12)], [3.0, 6.0], [0.6, 1.2]))
    assert a2[:6].sum() == pytest.approx(3.0) and g2[6:].sum() == pytest.approx(1.2)
    # the model behind it, on the measured 8-GPU and 2-GPU runs: step_r - (max advance + own gather) is one constant
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f, lo, hi in (("r02_bench_n8.json", 0.12, 0.14), ("r02_bench_n2.json", 0.10, 0.13)):
        d = json.load(open(os.path.join(root, "profiles", f)))["strong_c5"]
        res = np.array(d["ms_per_step_per_rank"]) - (max(d["ms_advance_per_rank"]) + np.array(d["ms_project_remesh_per_rank"]))
        assert lo < res.min() and res.max() < hi and res.max() - res.min() < 0.01


@pytest.mark.parametrize("seed,world,halo", [(0, 2, 5), (3, 3, 5), (6, 2, 6)])
def test_random_configurations_over_gloo(tmp_path, seed, world, halo):
    """random configurations (tall grids) in 2 and 3 strips over gloo through StripStepper's torch transport — bit for bit
    against the single-domain oracle after every step.  (The halo covers the reach: the host build behind ShimStripEngine
    has no halo_widen; the library's widening is tested on the GPU, tests/test_gpu_multi.py.)"""
    import torch.multiprocessing as mp

    import dist_worker
    out = tmp_path / "result"
    mp.spawn(dist_worker.run, args=(world, _free_port(), f"fuzz:{seed}", halo, str(out)), nprocs=world, join=True)
    res = out.read_text()
    if res == "skip":
        pytest.skip("deposits reach further than the halo or a strip is shorter than its halo")
    assert res == "ok"
