"""Generate the golden fixtures of tests/golden/ from the CPU oracle.

The reference holds no golden vectors for this path (zero assertions) and Julia is
not installed here, so these fixtures freeze the ORACLE's results (parity unpinned against
the real reference).  They guard the oracle, pmath.h and the device code against silent
changes; regenerate only when the specification is changed on purpose:

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]

from common import make_oracle  # noqa: E402
from scenarios import SCENARIOS  # noqa: E402

CASES = {"minimal": 13, "minimal_dp5": 8, "tripolar": 6, "growing_winds": 8, "land_block": 8}


def run_case(name, nsteps):
    g, P, wind, DT, _ = SCENARIOS[name]()
    o = make_oracle(g, P)
    o.seed(*wind(0.0))
    t = 0.0
    counters = []
    for _ in range(nsteps):
        o.step(t, DT, *wind(t), *wind(t + DT))
        t += DT
        c = o.counters()
        counters.append([c["n_substeps"], c["n_rejects"], c["n_rhs"], c["reach"], c["n_remesh_A"], c["n_remesh_B"],
                         c["n_remesh_D"], c["n_reseed_advance"]])
    p = o.particles()
    return dict(state=o.state(), z=p["z"], flags=p["flags"], counters=np.array(counters, dtype=np.int64))


if __name__ == "__main__":
    for name, n in CASES.items():
        out = run_case(name, n)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
