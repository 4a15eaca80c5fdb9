"""A third, separately written restatement of the adaptive Runge-Kutta stepping the reference gets from
OrdinaryDiffEq.jl (`step!(integ, DT, true)`, mapping_2D.jl:152; `auto_dt_reset!`, mapping_2D.jl:95,103,110) —
plain Python float64 loops written from the published algorithm, NOT from
oracle/picles_oracle.c or physics.h.  It shares only the right-hand side with the oracle (oracle.rhs, itself held
to 50-digit arithmetic by test_independent_math.py).  The Dormand-Prince coefficients are not typed here at all:
they are scipy's own copy (scipy.integrate RK45: A, B, C, E), a third-party-held tableau.

What it guards: the oracle's and the device's integrators were written by the same builder; a slip in the
loop logic (stage times, FSAL reuse, error norm, PI controller memory, step-size clipping, tstop snapping, the Hairer
initial step) that both share would pass every bit-exact test.  Here the oracle has to reproduce the substep
COUNT, the step-size sequence's end (`dt`, `qold`) and the final state of an implementation that never saw its
code, to rounding level (1e-11: the two differ in operation order and in the `pow` of the controller only).

PARITY UNPINNED still holds: this is a restatement of the published algorithm, not OrdinaryDiffEq itself."""
import json
import math
import os

import numpy as np
import pytest

import oracle
from common import default_params

M = (1 / 2000.0, 0.0, 0.0, 1 / 2000.0)

# Tsitouras 2011, "Runge-Kutta pairs of order 5(4) satisfying only the first column simplifying assumption",
# the coefficients OrdinaryDiffEq's Tsit5 uses (b = last row of A: first-same-as-last)
_TS_C = [0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0]
_TS_A = [
    [],
    [0.161],
    [-0.008480655492356989, 0.335480655492357],
    [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
    [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
    [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
    [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774],
]
_TS_BT = [-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
          0.5823571654525552, -0.45808210592918697, 0.015151515151515152]


def _tableau(solver):
    """(c, a, btilde, beta1, beta2): 7 stages, the 7th evaluated at the new point (FSAL)."""
    if solver == "Tsit5":
        return _TS_C, _TS_A, _TS_BT, 7.0 / 50.0, 2.0 / 25.0
    from scipy.integrate._ivp.rk import RK45
    A, B, Cc, E = np.asarray(RK45.A), np.asarray(RK45.B), np.asarray(RK45.C), np.asarray(RK45.E)
    a = [list(A[i, :i]) for i in range(6)] + [list(B)]
    return list(Cc) + [1.0], a, list(E), 0.2 - 3 * 0.04 / 4, 0.04


def _rms(v):
    return math.sqrt(sum(x * x for x in v) / len(v))


def jmin(a, b):
    """Julia's min: a NaN operand wins (Python's min keeps whichever came first)."""
    return math.nan if (a != a or b != b) else min(a, b)


def jmax(a, b):
    return math.nan if (a != a or b != b) else max(a, b)


def _initdt(f, u0, t, abstol, reltol, dtmin, dtmax, order=5):
    """Hairer, Norsett, Wanner I, II.4 'starting step size' as OrdinaryDiffEq's initdt.jl applies it."""
    f0 = f(t, u0)
    sk = [abstol + abs(x) * reltol for x in u0]
    d0 = _rms([x / s for x, s in zip(u0, sk)])
    d1 = _rms([x / s for x, s in zip(f0, sk)])
    dt0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * (d0 / d1)
    dt0 = min(dt0, dtmax)
    u1 = [x + dt0 * k for x, k in zip(u0, f0)]
    f1 = f(t + dt0, u1)
    d2 = _rms([(b - a) / s for a, b, s in zip(f0, f1, sk)]) / dt0
    dm = max(d1, d2)
    dt1 = max(1e-6, dt0 * 1e-3) if dm <= 1e-15 else 10.0 ** (-(2.0 + math.log10(dm)) / order)
    return max(dtmin, min(100.0 * dt0, dt1, dtmax)), f0, 2


def integrate_python(f, u, t, DT, dt, qold, solver, abstol, reltol, dtmin, dtmax, force_dtmin, dt_reset=False,
                     gamma=0.9, qmin=0.2, qmax=10.0, qoldinit=1e-4, nan_eest_rejects=False):
    """`step!(integ, DT, true)`: adaptive steps until the added tstop t + DT is reached."""
    c, a, bt, beta1, beta2 = _tableau(solver)
    tstop = t + DT
    n_rhs = n_acc = n_rej = 0
    u = list(u)
    k1 = None
    if dt_reset:
        dt, k1, n = _initdt(f, u, t, abstol, reltol, dtmin, dtmax)
        n_rhs += n - 1               # f(u0, t) is the integrator's first stage: counted below as such
    if k1 is None:
        k1 = f(t, u)
    n_rhs += 1
    retcode, u_trial = "Success", None
    while t < tstop:
        dt = jmax(jmin(dt, dtmax), dtmin)
        dt = jmin(dt, tstop - t)
        if dt != dt:                  # check_error!: "NaN dt detected"
            retcode = "DtNaN"
            break
        K = [k1]
        for s in range(1, 7):
            us = [u[j] + dt * sum(a[s][i] * K[i][j] for i in range(s)) for j in range(5)]
            K.append(f(t + c[s] * dt, us))
            n_rhs += 1
        unew = us                     # stage 7 is evaluated AT the new point: us of s = 6 is u_{n+1}
        err = [dt * sum(bt[i] * K[i][j] for i in range(7)) for j in range(5)]
        EEst = _rms([e / (abstol + max(abs(x), abs(y)) * reltol) for e, x, y in zip(err, u, unew)])
        if EEst == 0.0:
            q = q11 = 1.0 / qmax
        else:
            q11 = EEst ** beta1
            q = jmax(1.0 / qmax, jmin(1.0 / qmin, (q11 / qold ** beta2) / gamma))
        if EEst <= 1.0 or (force_dtmin and abs(dt) <= dtmin):
            n_acc += 1
            if q == 1.0:              # qsteady_min = qsteady_max = 1
                q = 1.0
            qold = max(EEst, qoldinit)
            tn = t + dt
            t = tstop if abs(tstop - tn) <= 100 * np.spacing(tstop) else tn
            u, k1 = unew, K[6]
            dt = dt / q
        else:
            n_rej += 1
            if nan_eest_rejects and q11 != q11:
                # OrdinaryDiffEq's fastpow / FastPower.fastpower: Float32(NaN) read as bits is 2^(β1·128.56), a large finite
                # q11, so min(1/qmin, q11/γ) = 1/qmin — the step is rejected by the full factor and the run goes on
                dt = dt / (1.0 / qmin)
            else:
                dt = dt / jmin(1.0 / qmin, q11 / gamma)
            u_trial = unew            # an in-place integrator's `u` holds the rejected trial until the next attempt
    return dict(u=np.array(u), t=t, dt=dt, qold=qold, n_acc=n_acc, n_rej=n_rej, n_rhs=n_rhs, retcode=retcode,
                u_trial=u_trial)


def _compare(P, solver, wind, z0, n_DT, dt_reset_each=False, dt0=1e-3, DT=600.0):
    f = lambda t, z: list(oracle.rhs(P, np.asarray(z, np.float64), wind[0], wind[1], M=M))
    ro = dict(u=np.asarray(z0, np.float64), t=0.0, dt=dt0, qold=1e-4, iter=0)
    rp = dict(u=np.asarray(z0, np.float64), t=0.0, dt=dt0, qold=1e-4)
    tot = 0
    for _ in range(n_DT):
        ro = oracle.integrate_one(P, ro["u"], t=ro["t"], dt=ro["dt"], qold=ro["qold"], it=ro["iter"], wind0=wind,
                                  DT=DT, M=M, dt_reset=dt_reset_each)
        rp = integrate_python(f, rp["u"], rp["t"], DT, rp["dt"], rp["qold"], solver, P.abstol, P.reltol, P.dtmin,
                              P.dtmax if P.dtmax > 0 else math.inf, bool(P.force_dtmin), dt_reset=dt_reset_each)
        assert ro["status"] == 0
        co = ro["counters"]
        assert (co["n_substeps"], co["n_rejects"]) == (rp["n_acc"], rp["n_rej"])
        assert co["n_rhs"] == rp["n_rhs"]
        assert ro["t"] == rp["t"]
        scale = np.maximum(np.abs(rp["u"]), 1e-6)
        assert np.max(np.abs(ro["u"] - rp["u"]) / scale) < 1e-11
        # qold is the last error estimate: a sum of seven stage derivatives weighted by b - b̂ that cancels to ~1e-7 of
        # its terms, so operation order shows at 1e-9 … 1e-8 relative there (and, through q ~ EEst^β1, at a tenth of
        # that in the proposed dt); the state itself carries none of it
        assert ro["dt"] == pytest.approx(rp["dt"], rel=1e-7)
        assert ro["qold"] == pytest.approx(rp["qold"], rel=1e-6)
        tot += rp["n_acc"] + rp["n_rej"]
    return tot


@pytest.mark.parametrize("solver", ["Tsit5", "DP5"])
@pytest.mark.parametrize("wind", [(10.0, 10.0), (10.0, 4.0), (-7.0, 3.0)])
def test_oracle_integrator_matches_an_independent_restatement(solver, wind):
    """The example's settings (dt = 1e-3 governs the first substep after seeding; the controller carries dt and
    qold from one model step into the next, as branch D of the remesh leaves them)."""
    P = default_params(solver=solver)
    z0, _, _ = oracle.windsea(wind[0], wind[1], 600)
    assert _compare(P, solver, wind, z0, 3) > 15


@pytest.mark.parametrize("solver", ["Tsit5", "DP5"])
def test_oracle_initial_step_matches_an_independent_restatement(solver):
    """`auto_dt_reset!` before every model step (what every remesh branch but D does)."""
    P = default_params(solver=solver)
    z0, _, _ = oracle.windsea(10.0, 4.0, 600)
    _compare(P, solver, (10.0, 4.0), z0, 3, dt_reset_each=True)


@pytest.mark.parametrize("solver", ["Tsit5", "DP5"])
def test_oracle_rejections_match_an_independent_restatement(solver):
    """A first step far too long for the tolerances: rejected steps, the reject rule of the PI controller
    (dt / min(1/qmin, q11/gamma)) and the controller memory across them."""
    P = default_params(solver=solver, dt=20.0, dtmin=1e-4)
    P.reltol, P.abstol = 1e-7, 1e-9
    z0, _, _ = oracle.windsea(10.0, 4.0, 600)
    f = lambda t, z: list(oracle.rhs(P, np.asarray(z, np.float64), 10.0, 4.0, M=M))
    rp = integrate_python(f, z0, 0.0, 600.0, 20.0, 1e-4, solver, P.abstol, P.reltol, P.dtmin, math.inf, True)
    assert rp["n_rej"] >= 1 and rp["retcode"] == "Success"
    _compare(P, solver, (10.0, 4.0), z0, 2, dt0=20.0)


def test_bench06_settings_dp5():
    """bench06's solver settings (DP5, dt = 10, dtmin = 1; benchmarks/bench06…jl:88-103): dtmin large enough for
    `dt = max(dt, dtmin)` and the forced acceptance at dtmin to matter."""
    P = default_params(solver="DP5", dt=10.0, dtmin=1.0, log_energy_maximum=math.log(27))
    z0, _, _ = oracle.windsea(10.0, 10.0, 1800)
    _compare(P, "DP5", (10.0, 10.0), z0, 3, dt0=10.0)


# ---- the blow-up of a trial step, and the probe a machine with Julia can answer (julia/probe_integrator.jl) ----------
HERE = os.path.dirname(os.path.abspath(__file__))
PROBE = os.path.join(HERE, "golden", "julia", "integrator_probe.json")

# (name, solver, wind, dt, dtmin, number of model steps, time scale of the seed state); julia/probe_integrator.jl
# runs the same list on the real OrdinaryDiffEq
PROBE_CASES = [
    ("tsit5_u10_v10", "Tsit5", (10.0, 10.0), 1e-3, 1e-4, 3, 600.0),
    ("tsit5_u10_v4", "Tsit5", (10.0, 4.0), 1e-3, 1e-4, 3, 600.0),
    ("dp5_u10_v4", "DP5", (10.0, 4.0), 1e-3, 1e-4, 3, 600.0),
    ("dp5_bench06", "DP5", (10.0, 10.0), 10.0, 1.0, 3, 1800.0),
    ("dp5_blowup_u0_v14", "DP5", (0.0, 14.0), 1e-3, 1e-4, 2, 600.0),
    ("dp5_blowup_um7_v12", "DP5", (-7.0, 12.0), 1e-3, 1e-4, 2, 600.0),
    # the default solver on the stiff parameter set of tests/T04_2D_reg_test.jl (C_φ = c_β = 4e-2), and on the usual one
    ("auto_stiff_um10_vm10", "AutoTsit5", (-10.0, -10.0), 1e-3, 1e-4, 3, 600.0, 4e-2),
    ("auto_nonstiff_u10_v10", "AutoTsit5", (10.0, 10.0), 1e-3, 1e-4, 3, 600.0),
]


def _num(x):
    return float(x) if not isinstance(x, str) else float(x.replace("inf", "inf"))


def run_probe_oracle():
    """the probe's cases on the oracle, in the layout julia/probe_integrator.jl writes (counters cumulative)"""
    out = []
    for name, solver, wind, dt, dtmin, nDT, ts, *rest in PROBE_CASES:
        P = default_params(solver=solver, dt=dt, dtmin=dtmin, **({"C_phi": rest[0]} if rest else {}))
        z0, _, _ = oracle.windsea(wind[0], wind[1], ts)
        r = dict(u=np.asarray(z0, np.float64), t=0.0, dt=dt, qold=1e-4, iter=0, status=0, as_state=(0, 0))
        acc = rej = nf = 0
        steps = []
        for _ in range(nDT):
            r = oracle.integrate_one(P, r["u"], t=r["t"], dt=r["dt"], qold=r["qold"], it=r["iter"], wind0=wind, DT=600.0,
                                     M=M, status=r["status"], as_state=r["as_state"])
            acc += r["counters"]["n_substeps"]; rej += r["counters"]["n_rejects"]; nf += r["counters"]["n_rhs"]
            steps.append(dict(u=[float(x) for x in r["u"]], t=r["t"], dt=r["dt"], qold=r["qold"], naccept=acc, nreject=rej,
                              nf=nf, iter=r["iter"], current_alg=2 if r["as_state"][1] else 1, retcode="Success" if r["status"] == 0 else "status %d" % r["status"]))
        out.append(dict(name=name, solver=solver, wind=list(wind), dt=dt, dtmin=dtmin, timescale=ts,
                        z0=[float(x) for x in z0], steps=steps))
    return out


def compare_probe(ref_cases, got_cases, rtol_u=1e-6):
    """per case and model step: same retcode class, same accepted/rejected counts, state within rtol_u"""
    report = {}
    for ref, got in zip(ref_cases, got_cases):
        assert ref["name"] == got["name"]
        assert np.allclose([_num(x) for x in ref["z0"]], got["z0"], rtol=1e-12, atol=0), ref["name"]
        worst, same = 0.0, True
        for a, b in zip(ref["steps"], got["steps"]):
            ok_a, ok_b = a["retcode"] in ("Success", "Default"), b["retcode"] in ("Success", "Default")
            same &= (ok_a == ok_b) and (a["naccept"], a["nreject"]) == (b["naccept"], b["nreject"])
            if ok_a and ok_b:
                ua, ub = np.array([_num(x) for x in a["u"]]), np.array([_num(x) for x in b["u"]])
                worst = max(worst, float(np.max(np.abs(ua - ub) / np.maximum(np.abs(ua), 1e-6))))
        report[ref["name"]] = dict(same_step_sequence=bool(same), max_rel_u=worst, ref_retcodes=[s["retcode"] for s in ref["steps"]],
                                   got_retcodes=[s["retcode"] for s in got["steps"]])
    return report


@pytest.mark.parametrize("wind", [(0.0, 14.0), (-7.0, 12.0)])
def test_a_trial_step_that_overflows_ends_the_integrator(wind):
    """DP5 at the reference's tolerances (abstol 1e-4, reltol 1e-3) proposes, near |wind| = 14 m/s, a second substep of
    ~85 s whose stages overflow: EEst = NaN.  With exact powers in the PI controller Julia's `min(1/qmin, NaN/γ)` is NaN,
    the rejected step leaves dt = NaN and check_error! returns DtNaN.  The independent
    restatement (Julia min/max semantics) and the oracle agree on where that happens; the oracle flags the particle
    UNSTABLE, keeps its last accepted state and never integrates it again.  Tsit5 has no such case on a 21 x 21 scan
    of winds up to 20 m/s (profiles/README.md)."""
    from picles_b200._abi import PST_UNSTABLE
    P = default_params(solver="DP5")
    z0, _, _ = oracle.windsea(wind[0], wind[1], 600)
    f = lambda t, z: list(oracle.rhs(P, np.asarray(z, np.float64), wind[0], wind[1], M=M))
    rp = integrate_python(f, z0, 0.0, 600.0, 1e-3, 1e-4, "DP5", P.abstol, P.reltol, P.dtmin, math.inf, True)
    assert rp["retcode"] == "DtNaN" and rp["n_rej"] == 1 and not all(math.isfinite(x) for x in rp["u_trial"])
    ro = oracle.integrate_one(P, z0, t=0.0, dt=1e-3, DT=600.0, wind0=wind, M=M)
    assert ro["status"] == PST_UNSTABLE and ro["dt"] != ro["dt"]
    assert (ro["counters"]["n_substeps"], ro["counters"]["n_rejects"]) == (rp["n_acc"], rp["n_rej"])
    # t is a sum of proposed step sizes, which carry the error estimate's cancellation (see _compare): 1e-8, not 1e-12
    assert ro["t"] == pytest.approx(rp["t"], rel=1e-8) and 0.0 < ro["t"] < 600.0
    assert np.max(np.abs(ro["u"] - rp["u"]) / np.maximum(np.abs(rp["u"]), 1e-6)) < 1e-8   # the last ACCEPTED state
    again = oracle.integrate_one(P, ro["u"], t=ro["t"], dt=ro["dt"], qold=ro["qold"], it=ro["iter"], status=ro["status"],
                                 DT=600.0, wind0=wind, M=M)
    assert again["counters"]["n_rhs"] == 0 and again["t"] == ro["t"]
    # the same winds under Tsit5 integrate cleanly
    Pt = default_params(solver="Tsit5")
    assert oracle.integrate_one(Pt, z0, t=0.0, dt=1e-3, DT=600.0, wind0=wind, M=M)["status"] == 0


def test_probe_layout_round_trips(tmp_path):
    """the file julia/probe_integrator.jl writes, built here from the oracle's own run: loads and compares clean, the
    two blow-up cases are told apart from the clean ones"""
    got = run_probe_oracle()
    f = tmp_path / "integrator_probe.json"
    enc = lambda x: x if (isinstance(x, (int, str)) or math.isfinite(x)) else ("nan" if x != x else ("inf" if x > 0 else "-inf"))
    doc = json.loads(json.dumps({"generator": "test", "DT": 600.0, "cases": got}, default=str))
    for c in doc["cases"]:
        for s in c["steps"]:
            s["dt"], s["qold"] = enc(s["dt"]), enc(s["qold"])
    f.write_text(json.dumps(doc))
    back = json.loads(f.read_text())
    rep = compare_probe(back["cases"], got)
    assert all(r["same_step_sequence"] and r["max_rel_u"] == 0.0 for r in rep.values())
    assert [n for n, r in rep.items() if r["got_retcodes"][-1] != "Success"] == ["dp5_blowup_u0_v14", "dp5_blowup_um7_v12"]
    stiff = {c["name"]: [s["current_alg"] for s in c["steps"]] for c in got if c["name"].startswith("auto_")}
    assert stiff["auto_stiff_um10_vm10"] == [2, 2, 2] and stiff["auto_nonstiff_u10_v10"][0] == 1


@pytest.mark.skipif(not os.path.isfile(PROBE), reason="no integrator probe: run julia/probe_integrator.jl on a machine with Julia")
def test_oracle_against_the_julia_integrator_probe():
    with open(PROBE) as fh:
        ref = json.load(fh)
    rep = compare_probe(ref["cases"], run_probe_oracle())
    print("oracle vs OrdinaryDiffEq", ref.get("OrdinaryDiffEq"), json.dumps(rep, indent=1))
    for name, r in rep.items():
        if name.startswith("dp5_blowup") or name.startswith("auto_"):
            # reported, not asserted: what the reference does on an overflowing trial step depends on its OrdinaryDiffEq
            # version, and the oracle's AutoSwitch carries documented simplifications
            continue
        assert r["same_step_sequence"], (name, r)
        assert r["max_rel_u"] <= 1e-6, (name, r)


def test_wind_scan_where_trial_steps_overflow():
    """Where that happens: an isolated particle from its seed state, 6 model steps, winds on a 21 x 21 grid up to
    ±20 m/s, the reference's tolerances.  Tsit5 and AutoTsit5 (the `ODESettings` default): nowhere.  DP5 (bench06's
    solver): in three narrow bands of wind speed — the resonance of its proposed second substep with the right-hand
    side's stiff relaxation."""
    def scan(solver):
        P = default_params(solver=solver)
        bad = set()
        for U in np.linspace(-20, 20, 21):
            for V in np.linspace(-20, 20, 21):
                if U * U + V * V < 4:
                    continue
                z0, _, _ = oracle.windsea(U, V, 600)
                r = dict(u=z0, t=0.0, dt=1e-3, qold=1e-4, iter=0)
                for k in range(6):
                    r = oracle.integrate_one(P, r["u"], t=r["t"], dt=r["dt"], qold=r["qold"], it=r["iter"], DT=600.0,
                                             wind0=(U, V), M=M)
                    if r["status"]:
                        bad.add((round(math.hypot(U, V), 2), k, r["status"]))
                        break
        return bad
    assert scan("Tsit5") == set() and scan("AutoTsit5") == set()
    assert scan("DP5") == {(14.0, 0, 4), (24.08, 1, 4), (24.41, 1, 4)}


@pytest.mark.parametrize("wind", [(0.0, 14.0), (-7.0, 12.0), (20.0, 14.0)])
def test_nan_error_estimate_rejected_as_fastpow_does(wind):
    """picles_params_t::nan_eest_rejects = 1: the reading of OrdinaryDiffEq versions whose PI controller forms its powers
    with `fastpow` (DiffEqBase) or `fastpower` (FastPower.jl) — a NaN error estimate comes out of them as a large finite
    number, the overflowing trial step is rejected by 1/qmin and the integration goes on.  Oracle against the independent
    restatement on the winds where DP5 otherwise ends with DtNaN: same accepted / rejected counts, state to rounding level,
    the tstop reached."""
    P = default_params(solver="DP5", nan_eest_rejects=True)
    assert P.nan_eest_rejects == 1
    z0, _, _ = oracle.windsea(wind[0], wind[1], 600)
    f = lambda t, z: list(oracle.rhs(P, np.asarray(z, np.float64), wind[0], wind[1], M=M))
    ro = dict(u=np.asarray(z0, np.float64), t=0.0, dt=1e-3, qold=1e-4, iter=0)
    rp = dict(u=np.asarray(z0, np.float64), t=0.0, dt=1e-3, qold=1e-4)
    nan_rejects = 0
    for k in range(3):
        ro = oracle.integrate_one(P, ro["u"], t=ro["t"], dt=ro["dt"], qold=ro["qold"], it=ro["iter"], wind0=wind, DT=600.0, M=M)
        rp = integrate_python(f, rp["u"], rp["t"], 600.0, rp["dt"], rp["qold"], "DP5", P.abstol, P.reltol, P.dtmin, math.inf,
                              True, nan_eest_rejects=True)
        assert ro["status"] == 0 and rp["retcode"] == "Success" and ro["t"] == rp["t"] == 600.0 * (k + 1)
        assert (ro["counters"]["n_substeps"], ro["counters"]["n_rejects"]) == (rp["n_acc"], rp["n_rej"])
        assert np.max(np.abs(ro["u"] - rp["u"]) / np.maximum(np.abs(rp["u"]), 1e-6)) < 1e-9
        nan_rejects += rp["n_rej"]
    assert nan_rejects >= 1
    # and with the switch off the same winds end the integrator (the default: exact powers)
    P0 = default_params(solver="DP5")
    r = dict(u=np.asarray(z0, np.float64), t=0.0, dt=1e-3, qold=1e-4, iter=0, status=0)
    for k in range(3):
        r = oracle.integrate_one(P0, r["u"], t=r["t"], dt=r["dt"], qold=r["qold"], it=r["iter"], wind0=wind, DT=600.0, M=M,
                                 status=r["status"])
    assert r["status"] != 0
