#!/usr/bin/env python
"""Throughput of every BASELINE.json config shape on one GPU (supplement to bench.py, which
times configs[1] only).  One JSON line per config: particle-steps/s over steps W+1..W+K with
inputs resident, the per-kernel CUDA-event times and the substep statistics.

  python profiles/bench_configs.py [--steps 5] [--warmup 3] > profiles/rNN_configs.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from common import cartesian_grid, default_params, tripolar_grid  # noqa: E402
from picles_b200.engine import B200Engine  # noqa: E402


def engine_for(g, P):
    return B200Engine(g["Nx"], g["Ny"], g["bx"], g["by"], g["mask"], P, M=g["M"], M_const=g["M_const"], pc=g["pc"])


def growing(x):
    Lx = x.max()
    x0 = 50.0 / 260.0 * Lx
    ramp = np.where(x < x0, 0.01, (x - x0) / (Lx - x0))
    return lambda t: (10.0 * ramp * (0.6 + 0.4 * np.sin(2 * np.pi * t / 7200.0)), 3.0 * ramp + 0.05)


def tripolar(Nx, Ny, land):
    ocean = np.ones((Ny, Nx), np.uint8)
    ocean[: max(2, Ny // 40), :] = 0
    if land:
        yy, xx = np.mgrid[0:Ny, 0:Nx]
        for cx, cy, r in ((0.2, 0.45, 0.08), (0.55, 0.6, 0.1), (0.8, 0.3, 0.06), (0.5, 0.97, 0.04)):
            ocean[((xx - cx * Nx) / Nx) ** 2 + ((yy - cy * Ny) / Ny) ** 2 < r * r] = 0
    g = tripolar_grid(Nx, Ny, ocean=ocean)
    g["M"] = g["M"] * 1.2
    return g


def configs():
    tw = lambda t: (15.0, -10.0 * np.cos(5 * t / (3600 * 2 * np.pi)))
    yield "C1 example_00_minimal 51x51", cartesian_grid(51, 51), default_params(), (lambda t: (10.0, 10.0)), 600.0
    yield "C2 homogeneous box 4096x4096 (Tsit5, dt=1e-3, dtmin=1e-4)", cartesian_grid(4096, 4096), default_params(), (lambda t: (10.0, 10.0)), 600.0
    yield ("C2 homogeneous box 4096x4096 (bench06 settings: DP5, dt=10, dtmin=1, log_e_max=log 27, seed 30 min)",
           cartesian_grid(4096, 4096),
           default_params(solver="DP5", dt=10.0, dtmin=1.0, force_dtmin=False, log_energy_maximum=float(np.log(27)), timestep=1800.0),
           (lambda t: (10.0, 10.0)), 600.0)
    g = cartesian_grid(2048, 2048, dx=4000.0, dy=4000.0)
    yield "C3 growing/decaying winds 2048x2048, on/off thresholds", g, default_params(DT=1200.0, wind_min_squared=2.0), growing(g["x"][0]), 1200.0
    yield ("C3 growing/decaying winds 2048x2048, default solver AutoTsit5(Rosenbrock23()): the stiff branch is active", g,
           default_params(DT=1200.0, wind_min_squared=2.0, solver="AutoTsit5"), growing(g["x"][0]), 1200.0)
    yield "C4 tripolar aqua 2880x2160 (synthetic)", tripolar(2880, 2160, False), default_params(DT=1200.0, periodic_boundary=True), tw, 1200.0
    yield "C5 tripolar + land 4320x3840 (synthetic)", tripolar(4320, 3840, True), default_params(DT=1200.0, periodic_boundary=True), tw, 1200.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default=None, help="substring of the config names to run")
    a = ap.parse_args()
    for name, g, P, wind, DT in configs():
        if a.only and a.only not in name:
            continue
        e = engine_for(g, P)
        u0, v0 = wind(0.0)
        e.seed(u0, v0)
        t = 0.0
        full = lambda x: np.ascontiguousarray(np.broadcast_to(np.asarray(x, np.float64), (g["Ny"], g["Nx"])))
        rows = []
        ms_total = 0.0
        for k in range(a.warmup + a.steps):
            w = [full(x) for x in (*wind(t), *wind(t + DT))]
            e.upload_winds(*w)
            e.synchronize()
            e.timer_start()
            e.step(t, DT)                       # winds resident
            ms = e.timer_stop()
            t += DT
            if k >= a.warmup:
                rows.append(e.counters())
                ms_total += ms
        n_active = sum(r["n_active"] for r in rows)
        integ = max(sum(r["n_integrated"] for r in rows), 1)
        print(json.dumps({
            "config": name, "nodes": g["Nx"] * g["Ny"], "active_per_step": rows[-1]["n_active"],
            "particle_steps_per_s": n_active / (ms_total * 1e-3), "ms_per_step": ms_total / a.steps,
            "ms_advance": float(np.mean([r["ms_advance"] for r in rows])),
            "ms_project_remesh": float(np.mean([r["ms_project"] for r in rows])),
            "substeps_per_particle_step": sum(r["n_substeps"] for r in rows) / integ,
            "rhs_per_particle_step": sum(r["n_rhs"] for r in rows) / integ,
            "max_attempts": max(r["max_attempts"] for r in rows), "rejects": sum(r["n_rejects"] for r in rows),
            "failed": sum(r["n_failed"] for r in rows), "reach": max(r["reach"] for r in rows),
            "stiff_switches": sum(r["n_stiff_switches"] for r in rows), "stiff_attempts": sum(r["n_stiff_attempts"] for r in rows),
            "remesh_ABCD": [rows[-1]["n_remesh_" + c] for c in "ABCD"]}), flush=True)
        e.close()


if __name__ == "__main__":
    main()
