"""GPU tests (-m gpu): the CUDA path through the C ABI against the committed golden fixtures of
tests/golden/ (frozen oracle outputs; the oracle itself is not run here)."""
import copy
import os

import numpy as np
import pytest

from common import bits_equal
from golden.make_golden import CASES
from scenarios import SCENARIOS
from test_gpu_parity import engine_for

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_golden_fixture(gpu_lib, name):
    ref = np.load(os.path.join(HERE, "golden", f"{name}.npz"))
    base, _, opt = name.partition("+")
    g, P, wind, DT, _ = SCENARIOS[base]()
    n_mid = 0
    if opt == "autotsit5":
        P = copy.copy(P)
        P.solver = 2
    elif opt.startswith("midlevels"):
        n_mid = int(opt[len("midlevels"):])
    e = engine_for(g, P)
    e.seed(*wind(0.0))
    t = 0.0
    rows = []
    for _ in range(CASES[name]):
        if n_mid:
            lv = [wind(t + DT * float(k) / float(n_mid + 1)) for k in range(1, n_mid + 1)]
            e.set_wind_midlevels([a for a, _ in lv], [b for _, b in lv])
        e.step(t, DT, *wind(t), *wind(t + DT))
        t += DT
        c = e.counters()
        rows.append([c["n_substeps"], c["n_rejects"], c["n_rhs"], c["reach"], c["n_remesh_A"], c["n_remesh_B"],
                     c["n_remesh_D"], c["n_reseed_advance"]])
    assert bits_equal(e.state(), ref["state"])
    p = e.particles()
    act = (ref["flags"] & 8) != 0
    for k in range(5):
        assert bits_equal(p["z"][k][act], ref["z"][k][act])
    assert np.array_equal(p["flags"][act], ref["flags"][act])
    assert np.array_equal(np.array(rows, dtype=np.int64), ref["counters"])
