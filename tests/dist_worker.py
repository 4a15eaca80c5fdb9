"""Worker of the world_size>1 CPU tests: one rank = one y-strip of the host build of the
device code, driven by picles_b200.distributed.StripStepper over the gloo backend."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.dirname(HERE), HERE):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def mesh_of(g, DT, nsteps, seed=7):
    """a synthetic wind mesh covering the scenario's grid and duration (knots off the step
    boundaries so kinks fall inside steps)"""
    rng = np.random.default_rng(seed)
    x0, x1, y0, y1 = g["x"].min(), g["x"].max(), g["y"].min(), g["y"].max()
    px, py = 0.07 * (x1 - x0) + 1.0, 0.07 * (y1 - y0) + 1.0
    xw, yw = np.linspace(x0 - px, x1 + px, 8), np.linspace(y0 - py, y1 + py, 7)
    tw = np.linspace(0.0, DT * nsteps * 1.13, 5)
    U = 9.0 + 4.0 * rng.random((tw.size, yw.size, xw.size))
    V = 5.0 + 4.0 * rng.random((tw.size, yw.size, xw.size))
    return xw, yw, tw, U, V


def run(rank, world, port, name, halo, result_path, backend="gloo", transport="torch-p2p", wind_mode="host", n_mid=0):
    """backend gloo: host build of the device code per rank (CPU tests);
    backend nccl: one B200Engine per rank on cuda:<rank> (GPU tests, world <= device count)."""
    import torch.distributed as dist

    from common import BND_PERIODIC, ShimStripEngine, TALLY_NAMES, bits_equal, make_oracle
    from picles_b200.distributed import StripStepper, strip_bounds
    from scenarios import SCENARIOS

    if backend == "nccl":
        import torch
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        if name.startswith("fuzz:"):
            # a random configuration of tests/test_independent_model.py on a grid four times as tall, its winds staged
            from test_independent_model import fuzz_case
            g, P, closure, DT, _ = fuzz_case(int(name[5:]), ny_scale=4)
            nsteps = 4
            wind = lambda t: tuple(np.array([[closure(g["x"][j, i], g["y"][j, i], t)[k] for i in range(g["Nx"])]
                                             for j in range(g["Ny"])]) for k in (0, 1))
            # the host build behind ShimStripEngine has no halo_widen (the library has: tests/test_gpu_multi.py): the halo
            # given must cover the reach, and a strip must be as tall as its halo
            probe = make_oracle(g, P)
            probe.seed(*wind(0.0))
            reach = 1
            for k in range(nsteps):
                probe.step(k * DT, DT, *wind(k * DT), *wind((k + 1) * DT))
                reach = max(reach, probe.counters()["reach"])
            if reach > halo or min(b - a for a, b in strip_bounds(g["Ny"], world)) < halo:
                if rank == 0:
                    with open(result_path, "w") as f:
                        f.write("skip")
                return
        else:
            g, P, wind, DT, nsteps = SCENARIOS[name]()
        j0, j1 = strip_bounds(g["Ny"], world)[rank]
        if backend == "nccl":
            from picles_b200.engine import B200Engine
            eng = B200Engine(g["Nx"], g["Ny"], g["bx"], g["by"], g["mask"][j0:j1], P,
                             M=None if g["M"] is None else g["M"][:, j0:j1], M_const=g["M_const"],
                             pc=None if g["pc"] is None else g["pc"][j0:j1], device=rank, j0=j0, ny_local=j1 - j0,
                             halo=halo)
        else:
            eng = ShimStripEngine(g, P, j0, j1, halo)
        st = StripStepper(eng, rank, world, periodic_y=(g["by"] == BND_PERIODIC), transport=transport)
        full = lambda a: np.broadcast_to(np.asarray(a, np.float64), (g["Ny"], g["Nx"]))
        if wind_mode == "mesh":
            # wind ingestion: the strip samples its own nodes from the resident wind mesh
            import oracle
            mesh = mesh_of(g, DT, nsteps)
            eng.set_wind_mesh(*mesh, g["x"][j0:j1], g["y"][j0:j1])
            wind = lambda t: oracle.wind_mesh_sample(*mesh, g["x"], g["y"], t)
        u0, v0 = wind(0.0)
        if wind_mode == "mesh":
            eng.seed_wind_mesh(0.0)
        else:
            eng.seed(full(u0)[j0:j1], full(v0)[j0:j1])
        ref = None
        if rank == 0:
            ref = make_oracle(g, P)
            ref.seed(u0, v0)
        t = 0.0
        for _ in range(nsteps):
            w = [full(x) for x in (*wind(t), *wind(t + DT))]
            mids = [wind(t + DT * float(k) / float(n_mid + 1)) for k in range(1, n_mid + 1)]
            if wind_mode == "mesh":
                st.step_wind_mesh(t, DT, n_mid)
            else:
                if n_mid:
                    eng.set_wind_midlevels([full(a)[j0:j1] for a, _ in mids], [full(b)[j0:j1] for _, b in mids])
                st.step(t, DT, winds=[x[j0:j1] for x in w])
            if ref is not None:
                if n_mid:
                    ref.set_wind_midlevels([full(a) for a, _ in mids], [full(b) for _, b in mids])
                ref.step(t, DT, *w)
            t += DT
            parts = [None] * world if rank == 0 else None
            dist.gather_object((eng.state(), eng.particles(), eng.counters()), parts, dst=0)
            if rank == 0:
                S = np.concatenate([p[0] for p in parts], axis=1)
                assert bits_equal(ref.state(), S), f"{name}: State differs after t={t}"
                pr = ref.particles()
                act = (pr["flags"] & 8) != 0
                z = np.concatenate([p[1]["z"] for p in parts], axis=1)
                for k in range(5):
                    assert bits_equal(pr["z"][k][act], z[k][act]), f"{name}: particle component {k} differs"
                fl = np.concatenate([p[1]["flags"] for p in parts], axis=0)
                assert np.array_equal(pr["flags"][act], fl[act])
                cr = ref.counters()
                for nm in TALLY_NAMES:
                    agg = max(p[2][nm] for p in parts) if nm in ("reach", "max_attempts") else sum(p[2][nm] for p in parts)
                    assert cr[nm] == agg, f"{name}: counter {nm}: oracle {cr[nm]} vs strips {agg}"
        if rank == 0:
            with open(result_path, "w") as f:
                f.write("ok")
    finally:
        dist.destroy_process_group()
