"""I-cache hypothesis check: same kernel, longer model step (more RK attempts per particle)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from bench import params, workload
from picles_b200.engine import B200Engine
n = int(sys.argv[1]); DT = float(sys.argv[2])
W = workload(n, n, 1, 0)
e = B200Engine(W["Nx"], W["Ny"], 0, 0, W["mask"], params(), M_const=W["M_const"])
e.seed(10.0, 10.0)
t = 0.0
for k in range(4):
    e.step(t, DT); t += DT
    c = e.counters()
    print(k, DT, {x: c[x] for x in ("n_substeps", "n_rhs", "max_attempts", "ms_advance")}, "ns per rhs-warp", c["ms_advance"]*1e6/(c["n_rhs"]/32))
