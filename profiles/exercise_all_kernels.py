"""Small end-to-end exercise of every kernel (a driver for memory checkers; compute-sanitizer is closed on this pool, so it was only run plain): seed, steps with
host winds + intermediate levels, wind mesh steps, AutoTsit5 with parked particles, strips with
pack/unpack, fields, checkpoint.  usage: python profiles/exercise_all_kernels.py"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from scenarios import SCENARIOS  # noqa: E402
from picles_b200.engine import B200Engine  # noqa: E402


def engine(g, P, **kw):
    return B200Engine(g["Nx"], g["Ny"], g["bx"], g["by"], g["mask"], P, M=g["M"], M_const=g["M_const"], pc=g["pc"], **kw)


for name, solver in (("growing_winds", 2), ("tripolar", 0), ("fast_box", 0)):
    g, P, wind, DT, n = SCENARIOS[name]()
    P = copy.copy(P)
    P.solver = solver
    e = engine(g, P)
    e.seed(*wind(0.0))
    t = 0.0
    for k in range(min(n, 3)):
        mids = [wind(t + DT * j / 3.0) for j in (1, 2)]
        e.set_wind_midlevels([np.broadcast_to(a, (g["Ny"], g["Nx"])) for a, _ in mids],
                             [np.broadcast_to(b, (g["Ny"], g["Nx"])) for _, b in mids])
        e.step(t, DT, *wind(t), *wind(t + DT))
        t += DT
    e.fields()
    blob = e.checkpoint()
    e.restore(blob)
    xw = np.linspace(g["x"].min() - 1, g["x"].max() + 1, 7)
    yw = np.linspace(g["y"].min() - 1, g["y"].max() + 1, 6)
    tw = np.array([0.0, 1000.0, 5000.0])
    U = np.full((3, 6, 7), 9.0)
    e.set_wind_mesh(xw, yw, tw, U, U * 0.5, g["x"], g["y"])
    e.step_wind_mesh(t, DT, 1)
    e.sample_wind_mesh(123.0)
    print(name, "ok", e.counters()["n_stiff_attempts"])
    e.close()
# strips on one device: pack / unpack
g, P, wind, DT, n = SCENARIOS["minimal"]()
h = 2
a = B200Engine(g["Nx"], g["Ny"], 0, 0, g["mask"][:25], P, M_const=g["M_const"], j0=0, ny_local=25, halo=h)
b = B200Engine(g["Nx"], g["Ny"], 0, 0, g["mask"][25:], P, M_const=g["M_const"], j0=25, ny_local=g["Ny"] - 25, halo=h)
for e in (a, b):
    e.seed(10.0, 10.0)
for e in (a, b):
    e.step_advance(0.0, DT)
    e.halo_pack()
    e.synchronize()
(sa, na), (sb, nb) = a.halo_buffers(), b.halo_buffers()
a.copy_dev(sa[3], sb[0], na)
b.copy_dev(sb[2], sa[1], nb)
for e in (a, b):
    e.halo_unpack()
    e.step_project_remesh(0.0, DT)
print("strips ok")
