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


def run(rank, world, port, name, halo, result_path, backend="gloo", transport="torch-p2p"):
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
        u0, v0 = wind(0.0)
        eng.seed(full(u0)[j0:j1], full(v0)[j0:j1])
        ref = None
        if rank == 0:
            ref = make_oracle(g, P)
            ref.seed(u0, v0)
        t = 0.0
        for _ in range(nsteps):
            w = [full(x) for x in (*wind(t), *wind(t + DT))]
            st.step(t, DT, winds=[x[j0:j1] for x in w])
            if ref is not None:
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
