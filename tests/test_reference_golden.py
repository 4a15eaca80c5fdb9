"""Reference-held golden vectors (julia/generate_golden.jl), when a machine with Julia has produced them.

The reference's own tests pin nothing on this path (tests/runtests.jl:4-6 is empty) and Julia is not in this
image, so the directory tests/golden/julia/<case>/ is normally ABSENT and these tests skip — parity with the real
PiCLES stays "unpinned" (DESIGN.md §2).  Dropping the generator's output there turns them on: the oracle (and, with
-m gpu, the CUDA path) is run on the same configuration and the largest relative error on State, lne and c̄ is
reported against the north-star's tolerance of 1e-6.  The loader itself is exercised on every run with a synthetic
directory written in the generator's layout from the oracle's own results."""
import json
import os

import numpy as np
import pytest

from common import cartesian_grid, default_params, make_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CASE = os.path.join(HERE, "golden", "julia", "example_00_minimal")
REL_TOL = 1e-6  # north_star: lne and cg_bar within 1e-6 relative after the full simulation


def load_julia_golden(d):
    """manifest.json + raw little-endian arrays in Julia memory order (i fastest) -> dict of (.., Ny, Nx) arrays"""
    with open(os.path.join(d, "manifest.json")) as f:
        man = json.load(f)
    Nx, Ny, n = man["Nx"], man["Ny"], man["nsteps"] + 1
    rd = lambda name, dt, shape: np.fromfile(os.path.join(d, name), dtype=dt).reshape(shape)
    return dict(manifest=man, state=rd("state.f64", "<f8", (n, 3, Ny, Nx)), u=rd("particles_u.f64", "<f8", (n, 5, Ny, Nx)),
                t=rd("particles_t.f64", "<f8", (n, Ny, Nx)), on=rd("particles_on.u8", "u1", (n, Ny, Nx)))


def run_minimal(make_model):
    """example_00_minimal.jl:17-67 on a model with the oracle's interface: State and particles after the seed and
    after each of the floor(stop/DT)+1 = 13 steps of run!'s '>=' loop"""
    g = cartesian_grid(51, 51)
    m = make_model(g, default_params())
    m.seed(10.0, 10.0)
    S, U, T, ON = [m.state()], [], [], []

    def grab():
        p = m.particles()
        U.append(p["z"]); T.append(p["t"]); ON.append(p["flags"] & 1)
    grab()
    t = 0.0
    for _ in range(13):
        m.step(t, 600.0, 10.0, 10.0, 10.0, 10.0)
        t += 600.0
        S.append(m.state())
        grab()
    return dict(state=np.stack(S), u=np.stack(U), t=np.stack(T), on=np.stack(ON))


def rel_errors(ref, got, active):
    """largest relative error over the iterated particles / all nodes, per quantity"""
    out = {}
    e = np.abs(got["state"] - ref["state"]) / np.maximum(np.abs(ref["state"]), 1e-300)
    out["State"] = float(np.nanmax(np.where(np.abs(ref["state"]) > 1e-30, e, 0.0)))
    lne_r, lne_g = ref["u"][:, 0][:, active], got["u"][:, 0][:, active]
    out["lne"] = float(np.nanmax(np.abs(lne_g - lne_r) / np.maximum(np.abs(lne_r), 1e-300)))
    cr = np.hypot(ref["u"][:, 1][:, active], ref["u"][:, 2][:, active])
    out["cg_bar"] = float(np.nanmax(np.hypot(got["u"][:, 1][:, active] - ref["u"][:, 1][:, active],
                                             got["u"][:, 2][:, active] - ref["u"][:, 2][:, active]) / np.maximum(cr, 1e-300)))
    return out


def test_loader_reads_the_generator_layout(tmp_path):
    """a directory in the layout julia/generate_golden.jl writes (built here from the oracle's own run) loads back
    bit for bit and compares clean — so the day real vectors arrive only the numbers are new, not the plumbing"""
    res = run_minimal(make_oracle)
    d = tmp_path / "example_00_minimal"
    d.mkdir()
    res["state"].astype("<f8").tofile(d / "state.f64")
    res["u"].astype("<f8").tofile(d / "particles_u.f64")
    res["t"].astype("<f8").tofile(d / "particles_t.f64")
    res["on"].astype("u1").tofile(d / "particles_on.u8")
    (d / "manifest.json").write_text(json.dumps({"Nx": 51, "Ny": 51, "nsteps": 13, "DT": 600.0,
                                                 "on_flag_persists_in_structarray": False}))
    gold = load_julia_golden(str(d))
    assert gold["state"].shape == (14, 3, 51, 51) and np.array_equal(gold["state"], res["state"])
    active = np.zeros((51, 51), bool)
    active[1:-1, 1:-1] = True
    errs = rel_errors(gold, res, active)
    assert errs == {"State": 0.0, "lne": 0.0, "cg_bar": 0.0}


@pytest.mark.skipif(not os.path.isdir(CASE), reason="no reference-held vectors: run julia/generate_golden.jl on a machine with Julia")
def test_oracle_against_the_julia_reference():
    gold = load_julia_golden(CASE)
    man = gold["manifest"]
    assert (man["Nx"], man["Ny"], man["nsteps"]) == (51, 51, 13), man
    if man.get("on_flag_persists_in_structarray"):
        pytest.fail("the reference's `on` flag persists: SURVEY B-1 reads the other way — run the oracle with on_persist=1 "
                    "and revisit DESIGN.md's quirk table")
    res = run_minimal(make_oracle)
    active = np.zeros((51, 51), bool)
    active[1:-1, 1:-1] = True   # ocean_points: mask == 1
    errs = rel_errors(gold, res, active)
    print("oracle vs Julia reference, largest relative errors:", errs, "OrdinaryDiffEq", man.get("OrdinaryDiffEq"))
    assert np.array_equal(gold["on"][:, active], res["on"][:, active]), "on flags differ from the reference"
    assert np.allclose(gold["t"][:, active], res["t"][:, active], rtol=0, atol=1e-9), "integrator clocks differ"
    assert errs["lne"] <= REL_TOL and errs["cg_bar"] <= REL_TOL, errs


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.isdir(CASE), reason="no reference-held vectors: run julia/generate_golden.jl on a machine with Julia")
def test_gpu_against_the_julia_reference(gpu_lib):
    from picles_b200.engine import B200Engine
    gold = load_julia_golden(CASE)

    def make(g, P):
        return B200Engine(g["Nx"], g["Ny"], g["bx"], g["by"], g["mask"], P, M_const=g["M_const"])
    res = run_minimal(make)
    active = np.zeros((51, 51), bool)
    active[1:-1, 1:-1] = True
    errs = rel_errors(gold, res, active)
    print("CUDA path vs Julia reference, largest relative errors:", errs)
    assert errs["lne"] <= REL_TOL and errs["cg_bar"] <= REL_TOL, errs


# ---- the one-dimensional model (julia/generate_golden_1d.jl) ---------------------------------------------------
CASE_1D = os.path.join(HERE, "golden", "julia", "oned_steady")


def load_julia_golden_1d(d):
    with open(os.path.join(d, "manifest.json")) as f:
        man = json.load(f)
    Nx, n = man["Nx"], man["nsteps"] + 1
    rd = lambda name, dt, shape: np.fromfile(os.path.join(d, name), dtype=dt).reshape(shape)
    return dict(manifest=man, state=rd("state.f64", "<f8", (n, 3, Nx)), u=rd("particles_u.f64", "<f8", (n, 3, Nx)),
                t=rd("particles_t.f64", "<f8", (n, Nx)), on=rd("particles_on.u8", "u1", (n, Nx)))


def run_oned_steady(make_model):
    """scenario steady_nonperiodic = the configuration julia/generate_golden_1d.jl runs (8 steps)"""
    from scenarios_1d import SCENARIOS_1D
    g, P, wind, DT, steps = SCENARIOS_1D["steady_nonperiodic"]()
    m = make_model(g, P)
    x = g["x"]
    m.seed(wind(x, 0.0))
    S, U, T, ON = [m.state()], [], [], []

    def grab():
        p = m.particles()
        U.append(p["z"]); T.append(p["t"]); ON.append(p["flags"] & 1)
    grab()
    t = 0.0
    for _ in range(steps):
        m.step(t, DT, wind(x, t), wind(x, t + DT))
        t += DT
        S.append(m.state())
        grab()
    return dict(state=np.stack(S), u=np.stack(U), t=np.stack(T), on=np.stack(ON))


def _oracle_1d(g, P):
    from oracle import oned
    return oned.Oracle1D(g["Nx"], g["xmin"], g["dx"], g["x"], P)


def test_loader_reads_the_1d_generator_layout(tmp_path):
    res = run_oned_steady(_oracle_1d)
    d = tmp_path / "oned_steady"
    d.mkdir()
    res["state"].astype("<f8").tofile(d / "state.f64")
    res["u"].astype("<f8").tofile(d / "particles_u.f64")
    res["t"].astype("<f8").tofile(d / "particles_t.f64")
    res["on"].astype("u1").tofile(d / "particles_on.u8")
    (d / "manifest.json").write_text(json.dumps({"Nx": 51, "nsteps": 8, "DT": 600.0, "on_flag_persists": True}))
    gold = load_julia_golden_1d(str(d))
    assert gold["state"].shape == (9, 3, 51) and np.array_equal(gold["state"], res["state"])
    assert np.array_equal(gold["u"], res["u"]) and np.array_equal(gold["on"], res["on"])


@pytest.mark.skipif(not os.path.isdir(CASE_1D), reason="no reference-held 1-D vectors: run julia/generate_golden_1d.jl on a machine with Julia")
def test_oracle_1d_against_the_julia_reference():
    gold = load_julia_golden_1d(CASE_1D)
    man = gold["manifest"]
    assert (man["Nx"], man["nsteps"]) == (51, 8), man
    if not man.get("on_flag_persists", True):
        pytest.fail("`on` does not persist in the reference's 1-D ParticleCollection: DESIGN.md §4.6 reads the other way")
    res = run_oned_steady(_oracle_1d)
    interior = slice(1, 50)
    assert np.array_equal(gold["on"][:, interior], res["on"][:, interior]), "on flags differ from the reference"
    e_ref, e_got = gold["state"][:, 0, interior], res["state"][:, 0, interior]
    rel = np.abs(e_got - e_ref) / np.maximum(np.abs(e_ref), 1e-300)
    lne = np.abs(res["u"][:, 0, interior] - gold["u"][:, 0, interior]) / np.maximum(np.abs(gold["u"][:, 0, interior]), 1e-300)
    print("1-D oracle vs Julia reference: largest relative error State e", float(rel.max()), "lne", float(lne.max()))
    assert lne.max() <= REL_TOL and rel.max() <= 1e-5
