"""Multi-GPU tests (-m gpu; skipped when the box has fewer GPUs than ranks): one process
per GPU, strips exchanged over NCCL — inside the library (picles_halo_exchange) and
through torch.distributed P2P — bit-exact against the single-domain oracle."""
import socket

import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("name,world,halo,transport", [
    ("minimal", 2, 2, "nccl-lib"), ("periodic_grid", 2, 5, "nccl-lib"), ("growing_winds", 2, 2, "nccl-lib"),
    ("minimal", 2, 2, "torch-p2p"), ("tripolar", 3, 6, "nccl-lib"), ("land_block", 4, 2, "nccl-lib"),
    # one halo row, particles crossing 2-4 cells per step: the exchange widens itself (all-reduced reach, repeated
    # exchange + gather inside picles_step_strip; reach validated by the host for the torch transport)
    ("fast_box", 2, 1, "nccl-lib"), ("fast_box", 2, 1, "torch-p2p"), ("pulse_winds", 2, 2, "nccl-lib"),
    # strips taller than twice the supported reach take the overlapped path of picles_step_strip (boundary zones first,
    # their reach all-reduced while the interior integrates); fast_box above is one of them, with a widening exchange
    ("tripolar_tall", 2, 6, "nccl-lib")])
def test_strips_over_nccl_match_oracle(gpu_lib, tmp_path, name, world, halo, transport):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    import dist_worker
    out = tmp_path / "result"
    mp.spawn(dist_worker.run, args=(world, _free_port(), name, halo, str(out), "nccl", transport), nprocs=world,
             join=True)
    assert out.read_text() == "ok"


@pytest.mark.parametrize("name,world,halo,n_mid", [("minimal", 2, 2, 0), ("tripolar", 2, 6, 2)])
def test_strips_with_device_wind_mesh_over_nccl(gpu_lib, tmp_path, name, world, halo, n_mid):
    """picles_step_wind_mesh on strips: every rank samples its rows from its resident wind mesh,
    exchange inside the library, bit-exact against the single-domain oracle"""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    import dist_worker
    out = tmp_path / "result"
    mp.spawn(dist_worker.run, args=(world, _free_port(), name, halo, str(out), "nccl", "nccl-lib", "mesh", n_mid),
             nprocs=world, join=True)
    assert out.read_text() == "ok"


def test_halo_exchange_without_neighbours_is_pack_unpack(gpu_lib):
    """single strip handle with a halo and no neighbours: picles_step_strip == picles_step"""
    import numpy as np

    from common import cartesian_grid, default_params
    from picles_b200.engine import B200Engine
    g = cartesian_grid(40, 24)
    P = default_params()
    a = B200Engine(40, 24, 0, 0, g["mask"], P, M_const=g["M_const"])
    b = B200Engine(40, 24, 0, 0, g["mask"], P, M_const=g["M_const"], halo=2)
    for e in (a, b):
        e.seed(10.0, 10.0)
    t = 0.0
    for _ in range(3):
        a.step(t, 600.0)
        b.step_strip(t, 600.0, lo=-1, hi=-1)
        t += 600.0
    assert np.array_equal(a.state().view(np.uint64), b.state().view(np.uint64))
