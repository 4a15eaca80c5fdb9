import sys, os
sys.path[:0] = ["/root/repo", "/root/repo/tests", "/root/repo/profiles"]
import numpy as np
from bench_configs import tripolar, engine_for
from common import default_params
for f in (0.25, 1.0):
    g = tripolar(4320, 3840, True)
    g["M"] = g["M"] * f
    P = default_params(DT=1200.0, periodic_boundary=True)
    e = engine_for(g, P)
    e.seed(15.0, -10.0)
    t = 0.0
    for k in range(6):
        e.step(t, 1200.0, 15.0, -10.0, 15.0, -10.0); t += 1200.0
    c = e.counters()
    print(f, c["reach"], c["ms_advance"], c["ms_project"], c["n_remesh_A"], c["n_remesh_B"], c["n_remesh_D"])
    e.close()
