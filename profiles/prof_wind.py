"""k_wind_sample at the bench size: CUDA-event time per launch and achieved HBM bandwidth
(32 algorithmic bytes per node) for an ERA5-like 1-degree mesh (361 x 181 x 41 knots, uniform)
and a non-uniform one.  usage: python profiles/prof_wind.py [nx ny]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from bench import params  # noqa: E402
from picles_b200.engine import B200Engine  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
mask = np.ones((ny, nx), np.uint8)
eng = B200Engine(nx, ny, 0, 0, mask, params(), M_const=np.array([5e-4, 0, 0, 5e-4]))
X = np.broadcast_to(np.arange(nx) * 2000.0, (ny, nx))
Y = np.broadcast_to((np.arange(ny) * 2000.0)[:, None], (ny, nx))
rng = np.random.default_rng(0)
for name, xw, yw in (("uniform 361x181x41", np.linspace(-1e4, 2000.0 * nx + 1e4, 361), np.linspace(-1e4, 2000.0 * ny + 1e4, 181)),
                     ("non-uniform 361x181x41", np.cumsum(rng.uniform(0.3, 1.7, 361)) * (2000.0 * nx / 361), np.cumsum(rng.uniform(0.3, 1.7, 181)) * (2000.0 * ny / 181))):
    tw = np.arange(41) * 21600.0
    U = rng.normal(8, 4, (tw.size, yw.size, xw.size))
    V = rng.normal(-3, 5, (tw.size, yw.size, xw.size))
    eng.set_wind_mesh(xw, yw, tw, U, V, X, Y)
    ms = eng.measure_wind_sample(12345.0, 50)
    print(json.dumps({"kernel": "k_wind_sample", "mesh": name, "nodes": nx * ny, "ms_per_launch": ms,
                      "GBps_algorithmic_32B_per_node": nx * ny * 32 / ms / 1e6}))
