"""Minimal driver (DP5): seed + W warm-up + K model steps of the homogeneous box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402

from bench import params, workload  # noqa: E402
from picles_b200.engine import B200Engine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
W = workload(n, n, 1, 0)
e = B200Engine(W["Nx"], W["Ny"], 0, 0, W["mask"], params("DP5"), M_const=W["M_const"])
e.seed(10.0, 10.0)
t = 0.0
for k in range(steps):
    e.step(t, 600.0)
    t += 600.0
    c = e.counters()
    print(k, {x: c[x] for x in ("n_substeps", "n_rhs", "max_attempts", "ms_advance", "ms_project", "ms_remesh")})
