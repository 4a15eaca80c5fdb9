#!/usr/bin/env python
"""The one-dimensional model on one GPU: milliseconds per model step and particle-steps/s for the reference's grid
sizes (21-201 nodes, tests/B01_1D_regtest_wave_growth.jl:60-75) and for large chains.  Launch-bound by design
(DESIGN.md §4.6): three kernels per step on a few hundred particles.

    python profiles/bench_oned.py > profiles/rNN_oned.jsonl
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from scenarios_1d import grid_1d, params_1d  # noqa: E402

from picles_b200.engine1d import B200Engine1D  # noqa: E402

for Nx in (51, 201, 2001, 20001, 200001):
    g = grid_1d(0.0, 3e4 * (Nx - 1), Nx)
    e = B200Engine1D(g["Nx"], g["xmin"], g["dx"], g["x"], params_1d(600.0, periodic=True))
    u = np.full(Nx, 15.0)
    e.seed(u)
    t = 0.0
    for _ in range(3):
        e.step(t, 600.0, u, u)
        t += 600.0
    rows, t0 = [], time.perf_counter()
    for _ in range(10):
        e.step(t, 600.0, u, u)
        t += 600.0
        rows.append(e.counters())
    wall = (time.perf_counter() - t0) / 10
    print(json.dumps({"Nx": Nx, "ms_per_step_wall": wall * 1e3, "ms_advance": float(np.mean([r["ms_advance"] for r in rows])),
                      "ms_project_remesh": float(np.mean([r["ms_project"] for r in rows])), "particle_steps_per_s": Nx / wall,
                      "substeps_per_particle_step": float(np.mean([r["n_substeps"] / max(r["n_integrated"], 1) for r in rows])),
                      "reach": int(max(r["reach"] for r in rows))}))
    e.close()
