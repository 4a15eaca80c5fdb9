"""Generate the golden fixtures of tests/golden/ from the CPU oracle.

The reference holds no golden vectors for this path (zero assertions) and Julia is
not installed here, so these fixtures freeze the ORACLE's results (parity unpinned against
the real reference).  They guard the oracle, pmath.h and the device code against silent
changes; regenerate only when the specification is changed on purpose:

    python tests/golden/make_golden.py
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

from common import make_oracle  # noqa: E402
from scenarios import SCENARIOS  # noqa: E402

CASES = {"minimal": 13, "minimal_dp5": 8, "tripolar": 6, "growing_winds": 8, "land_block": 8,
         # the default solver AutoTsit5(Rosenbrock23()) through its stiff branch (solver id 2)
         "growing_winds+autotsit5": 8,
         # two intermediate wind levels per step (cubic in time), time-varying winds
         "growing_winds+midlevels2": 6}


def run_case(name, nsteps):
    base, _, opt = name.partition("+")
    g, P, wind, DT, _ = SCENARIOS[base]()
    n_mid = 0
    if opt == "autotsit5":
        P = copy.copy(P)
        P.solver = 2
    elif opt.startswith("midlevels"):
        n_mid = int(opt[len("midlevels"):])
    o = make_oracle(g, P)
    o.seed(*wind(0.0))
    t = 0.0
    counters = []
    for _ in range(nsteps):
        if n_mid:
            lv = [wind(t + DT * float(k) / float(n_mid + 1)) for k in range(1, n_mid + 1)]
            o.set_wind_midlevels([a for a, _ in lv], [b for _, b in lv])
        o.step(t, DT, *wind(t), *wind(t + DT))
        t += DT
        c = o.counters()
        counters.append([c["n_substeps"], c["n_rejects"], c["n_rhs"], c["reach"], c["n_remesh_A"], c["n_remesh_B"],
                         c["n_remesh_D"], c["n_reseed_advance"]])
    p = o.particles()
    return dict(state=o.state(), z=p["z"], flags=p["flags"], counters=np.array(counters, dtype=np.int64))


if __name__ == "__main__":
    only = sys.argv[1:]          # regenerate only the named cases (new ones), never the frozen ones by accident
    for name, n in CASES.items():
        if only and name not in only:
            continue
        out = run_case(name, n)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
