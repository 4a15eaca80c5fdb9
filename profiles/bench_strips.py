#!/usr/bin/env python
"""BASELINE.json configs[4] — "T03_PIC_tripolar_land with synthetic land mask at high resolution,
8-GPU y-strip partition" — and configs[3] on N GPUs: ONE global grid cut in N y-strips (strong
scaling: total work fixed), winds resident, halo exchange inside the library over NCCL.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P profiles/bench_strips.py [--config C5|C4] [--steps 5] [--warmup 3] [--halo 6]

One JSON line on rank 0: whole-job particle-steps/s (all ranks' active particles / max-over-ranks
CUDA-event time), per-rank kernel times (load balance) and the substep statistics.  N = 1 runs the
same code on a single strip (no exchange) for the efficiency denominator.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "profiles")]
from bench_configs import tripolar  # noqa: E402
from common import default_params  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C5", choices=["C4", "C5"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--halo", type=int, default=6)
    ap.add_argument("--balance", nargs="?", const="model", default=None, choices=["model", "measured"],
                    help="cut the strips by cost instead of by rows: 'model' = active particles + gather window "
                         "(picles_b200.distributed.row_cost_model); 'measured' = a calibration run on equal strips first "
                         "(per-strip kernel times and per-row reach, row_cost_measured), then the run proper on the re-cut strips")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)  # library banners -> stderr

    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from picles_b200.distributed import StripStepper, row_cost_measured, row_cost_model, strip_bounds, strip_bounds_weighted
    from picles_b200.engine import B200Engine

    if a.config == "C5":
        name, (Nx, Ny, land) = "C5 tripolar + land 4320x3840 (synthetic)", (4320, 3840, True)
    else:
        name, (Nx, Ny, land) = "C4 tripolar aqua 2880x2160 (synthetic)", (2880, 2160, False)
    g = tripolar(Nx, Ny, land)
    P = default_params(DT=1200.0, periodic_boundary=True)
    DT = 1200.0
    wind = lambda t: (15.0, -10.0 * np.cos(5 * t / (3600 * 2 * np.pi)))  # tests/T03_PIC_tripolar_aqua.jl:67-68
    loc_of = lambda j0, j1: (lambda x: np.ascontiguousarray(np.broadcast_to(np.asarray(x, np.float64), (j1 - j0, Nx))))
    def run(bounds, nwarm, nsteps):
        """seed + nwarm + nsteps steps on the strips `bounds`; timings and counters of the last nsteps"""
        j0, j1 = bounds[rank]
        halo = a.halo if world > 1 else 0
        eng = B200Engine(Nx, Ny, g["bx"], g["by"], g["mask"][j0:j1], P, M=g["M"][:, j0:j1], pc=g["pc"][j0:j1],
                         device=local_rank, j0=j0, ny_local=j1 - j0, halo=halo)
        st = StripStepper(eng, rank, world, periodic_y=False)
        loc = loc_of(j0, j1)
        eng.seed(*[loc(x) for x in wind(0.0)])

        def barrier():
            eng.synchronize()
            if dist is not None:
                dist.barrier()
                torch.cuda.synchronize()

        t, rows, ms_total = 0.0, [], 0.0
        for k in range(nwarm + nsteps):
            eng.upload_winds(*[loc(x) for x in (*wind(t), *wind(t + DT))])   # resident before the timed step
            barrier()
            eng.timer_start()
            st.step(t, DT)
            ms = eng.timer_stop()
            barrier()
            t += DT
            if k >= nwarm:
                rows.append(eng.counters())
                ms_total += ms
        mine = dict(ms=ms_total, active=sum(r["n_active"] for r in rows), integ=sum(r["n_integrated"] for r in rows),
                    sub=sum(r["n_substeps"] for r in rows), adv=float(np.mean([r["ms_advance"] for r in rows])),
                    prj=float(np.mean([r["ms_project"] for r in rows])), reach=max(r["reach"] for r in rows),
                    failed=sum(r["n_failed"] for r in rows), rows=j1 - j0, row_reach=eng.row_reach().tolist(),
                    halo_rows=eng.halo_rows()[0])
        eng.close()
        return mine, halo

    if a.balance == "model" and world > 1:
        # per-row reach estimate from the metric: group speed of the young sea of the first steps
        # (~3 m/s) x DT in cells of the row's smallest spacing
        cell = 1.0 / np.maximum(np.abs(g["M"][0]), np.abs(g["M"][3])).max(axis=1)
        reach_rows = np.maximum(np.ceil(3.0 * DT / cell), 1.0)
        bounds = strip_bounds_weighted(row_cost_model(g["mask"], True, reach_rows), world, min_rows=max(a.halo, 1))
    elif a.balance == "measured" and world > 1:
        eq = strip_bounds(Ny, world)
        cal, _ = run(eq, a.warmup, 2)                 # calibration on equal strips, the same step indices as the warm-up
        allc = [None] * world
        dist.all_gather_object(allc, cal)
        active_rows = ((g["mask"] == 1) | (g["mask"] == 3)).sum(axis=1)
        reach_rows = np.concatenate([np.asarray(c["row_reach"]) for c in allc])
        cost = row_cost_measured(active_rows, Nx, reach_rows, eq, [c["adv"] for c in allc], [c["prj"] for c in allc])
        bounds = strip_bounds_weighted(cost, world, min_rows=max(a.halo, 1))
    else:
        bounds = strip_bounds(Ny, world)
    mine, halo = run(bounds, a.warmup, a.steps)
    parts = [mine]
    if dist is not None:
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
    if rank == 0:
        ms_max = max(p["ms"] for p in parts)
        active = sum(p["active"] for p in parts)
        line = {"config": name, "n_gpus": world, "scaling": "strong",
                "partition": f"{world} y-strips of {Nx}x{Ny}, halo {halo} rows, cut by {('cost (' + a.balance + ')') if a.balance else 'rows'}",
                "rows_per_rank": [p["rows"] for p in parts],
                "nodes": Nx * Ny, "steps": a.steps, "warmup": a.warmup, "particle_steps_per_s": active / (ms_max * 1e-3),
                "ms_per_step": ms_max / a.steps, "ms_per_step_per_rank": [p["ms"] / a.steps for p in parts],
                "ms_advance_per_rank": [p["adv"] for p in parts], "ms_project_remesh_per_rank": [p["prj"] for p in parts],
                "active_per_rank_per_step": [p["active"] // a.steps for p in parts],
                "substeps_per_particle_step": sum(p["sub"] for p in parts) / max(sum(p["integ"] for p in parts), 1),
                "reach": max(p["reach"] for p in parts), "failed": sum(p["failed"] for p in parts)}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
