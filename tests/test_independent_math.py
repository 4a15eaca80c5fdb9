"""An independent statement of the reference's formulas in 50-digit arithmetic (mpmath), written from the
Julia sources — NOT from picles_b200/csrc/physics.h or the oracle — against which the oracle's float64
results are checked.

Why: oracle and device share pmath.h and windsea() has the same text in both: a
transcription error there would be common-mode and invisible to every bit-exact CUDA-vs-oracle test.  The
functions below restate, symbol by symbol,
    FetchRelations.jl:107-111,128-130,165-167,184-186,201-203,314-359,381-386,412-415
    ParticleSystems/particle_waves_v5.jl:212,215-225,242-249,271-275,281-297,317-346,479-556
    Operators/core_2D.jl:69-78,121-128
with exact real arithmetic in between, so agreement to ~1e-13 also bounds the rounding error of the float64
path (division, sqrt, exp, tanh, sech, pow kernels of pmath.h included).
"""
import math

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
mp.mp.dps = 50

import oracle  # noqa: E402
from common import default_params  # noqa: E402

F = mp.mpf


# ---- FetchRelations.jl ------------------------------------------------------------------------
def mp_windsea(U10, V10, time_scale):
    U10, V10, time_scale = F(U10), F(V10), F(time_scale)
    g = F("9.81")
    U_amp = mp.sqrt(U10 ** 2 + V10 ** 2)                       # :315
    if U_amp < F("0.1"):                                       # :316
        U_amp = F("0.1")
    tau = g * abs(time_scale) / abs(U_amp)                     # :318-319
    q_x, A, xi_0x = F("0.2748"), F("22.8013"), F("2.4097")     # Dulov_fetch_constants :107-111
    X_tilde = (tau / (A * xi_0x)) ** (1 / (1 - q_x))           # X_tilde_from_tau :128-130
    f_m = F("3.5") * (g / U_amp) * X_tilde ** F("-0.33")       # fₘ_from_X_tilde :165-167
    alpha_j = F("0.033") * (f_m * U_amp / g) ** F("0.67")      # alpha_j :184-186
    E = F("0.31") * g ** 2 * alpha_j * (f_m * 2 * mp.pi) ** (-4)  # E_JONSWAP :201-203
    f_peak = f_m * g / U_amp                                   # :331
    T_bar = F("0.9") * (1 / f_peak)                            # :340
    cg = g * T_bar / (4 * mp.pi)                               # :341
    return mp.log(E), cg * U10 / U_amp, cg * V10 / U_amp, E, cg


def mp_minimal_state(u, v, T):
    """MinimalWindsea normalises the wind to unit speed (:381-386); MinimalState = [E, m_x^2 + m_y^2] (:412-415)"""
    u, v = F(u), F(v)
    a = mp.sqrt(u ** 2 + v ** 2)
    lne, cx, cy, E, cg = mp_windsea(u / a, v / a, T)
    Ua = mp.sqrt((u / a) ** 2 + (v / a) ** 2)
    mx, my = (u / a) / Ua * E / (2 * cg), (v / a) / Ua * E / (2 * cg)
    return E, mx ** 2 + my ** 2


# ---- particle_waves_v5.jl ------------------------------------------------------------------------
def mp_rhs(P, z, u, v, M, pc):
    lne, cx, cy = F(z[0]), F(z[1]), F(z[2])
    u, v = F(u), F(v)
    r_g, C_a, C_e, C_phi = F(P.r_g), F(P.C_alpha), F(P.C_e), F(P.C_varphi)
    p, n, e_T = F(P.p), F(P.n), F(P.e_T)
    g = F("9.81")
    speed = lambda a, b: mp.sqrt(a ** 2 + b ** 2)              # :297
    cbar, us = speed(cx, cy), speed(u, v)
    c_gp = abs(cbar) / r_g                                     # :289-295
    kp = g / (4 * max(c_gp ** 2, F("1e-2")))                   # :283
    wp = g / (2 * max(abs(c_gp), F("0.1")))                    # :284
    gx, gy = cx / r_g, cy / r_g
    alpha = min(us / (2 * c_gp), F(500))                       # :215-225
    alpha_p = (u * gx + v * gy) / (2 * max(speed(gx, gy), F("1e-4")) ** 2)   # :212
    Hp = F("0.5") * (1 + mp.tanh(p * (alpha_p - F("0.85"))))   # :274
    Dp = 1 - F("1.25") * mp.sech(10 * (alpha_p - F("0.85"))) ** 2  # :275
    It = C_e * Hp * alpha ** 2 if P.input else F(0)            # :317-321
    Dt = mp.exp(n * lne) * (kp / e_T) ** (2 * n) if P.dissipation else F(0)   # :331-335
    Scg = C_a * Dp * kp ** 4 * mp.exp(2 * lne) if P.peak_shift else F(0)      # :340
    sg = speed(gx, gy)
    if us * sg == 0:                                           # :242-249
        s2 = F(0)
    else:
        s2 = (2 / (us * sg) ** 2) * (u * v * (2 * gy ** 2 - sg ** 2) - gx * gy * (2 * v ** 2 - us ** 2))
    Sdir = min(us / (2 * sg), F(500)) ** 2 * C_phi * Hp * s2 if P.direction else F(0)   # :345-346
    Ssph = cx * F(pc)                                          # :521
    terms = [[wp * r_g * Scg, wp * It, -wp * Dt],
             [-cx * wp * r_g * Scg, cy * Sdir, cy * Ssph],
             [-cy * wp * r_g * Scg, -cx * Sdir, -cx * Ssph]]   # :526-530
    if P.propagation:
        terms += [[F(M[0]) * cx, F(M[1]) * cy], [F(M[2]) * cx, F(M[3]) * cy]]    # :536
    else:
        terms += [[F(0)], [F(0)]]
    # error scale of a float64 evaluation of these lines: the terms with the two factors that are formed by
    # cancellation — H_p = (1 + tanh x)/2 (exactly 0 in float64 for x < -19) and Delta_p = 1 - 1.25 sech^2 —
    # at their ceilings 1 and 1.25: their ABSOLUTE error is an ulp of 1, whatever is left of them
    hp, dp = (1 / Hp if Hp != 0 else F(0)), (F("1.25") / abs(Dp) if Dp != 0 else F(0))
    ceil_ = [[abs(wp * r_g * Scg) * max(dp, 1), abs(wp * It) * max(hp, 1), abs(wp * Dt)],
             [abs(cx * wp * r_g * Scg) * max(dp, 1), abs(cy * Sdir) * max(hp, 1), abs(cy * Ssph)],
             [abs(cy * wp * r_g * Scg) * max(dp, 1), abs(cx * Sdir) * max(hp, 1), abs(cx * Ssph)],
             [abs(x) for x in terms[3]], [abs(x) for x in terms[4]]]
    return [sum(t) for t in terms], [sum(t) for t in ceil_]


REL = 2e-13   # float64 path against exact arithmetic: a few ulp through ~20 roundings per term


def close(a, b, scale=None, rel=REL):
    scale = abs(b) if scale is None else scale
    return abs(F(float(a)) - b) <= rel * scale + F("1e-300")


def test_windsea_against_exact_arithmetic():
    rng = np.random.default_rng(11)
    cases = [(10, 10, 600), (10, 10, 1200), (2, 2, 600), (0.05, 0.02, 600), (-8, 5, 900), (15, -10, 1800), (0, 3, 600),
             (1e-3, 1e-3, 600), (40, -35, 300)]
    cases += [(float(a), float(b), float(c)) for a, b, c in zip(rng.normal(0, 12, 200), rng.normal(0, 12, 200),
                                                                rng.uniform(60, 7200, 200))]
    for u, v, T in cases:
        out, E, cg = oracle.windsea(u, v, T)
        lne, cx, cy, Em, cgm = mp_windsea(u, v, T)
        assert close(E, Em) and close(cg, cgm), (u, v, T)
        assert close(out[1], cx, scale=cgm) and close(out[2], cy, scale=cgm)
        assert abs(F(float(out[0])) - lne) <= REL * max(abs(lne), 1), (u, v, T)   # log: absolute error in lne
    # SURVEY Appendix C spot values were derived from the same Julia lines with Python floats: a third statement
    out, E, cg = oracle.windsea(10, 10, 600)
    assert abs(E - 9.05035257588395e-4) < 1e-17 and abs(out[1] - 0.7412606338387) < 1e-12


def test_minimal_state_against_exact_arithmetic():
    for u, v, T in [(2, 2, 600), (10, 10, 600), (2, 2, 1200), (3, -1, 900)]:
        ms, _ = oracle.minimal_state(u, v, T)
        E, m2 = mp_minimal_state(u, v, T)
        assert close(ms[0], E) and close(ms[1], m2), (u, v, T)


@pytest.mark.parametrize("switches", [{}, {"direction": False}, {"input": False, "peak_shift": False}])
def test_right_hand_side_against_exact_arithmetic(switches):
    P = default_params(**switches)
    rng = np.random.default_rng(5)
    n = 300
    lne = rng.uniform(-14.0, 2.5, n)
    cg = np.exp(rng.uniform(np.log(0.01), np.log(15.0), n))     # |c̄| from the minimal particle to long swell
    th = rng.uniform(0, 2 * np.pi, n)
    U = rng.uniform(0.0, 30.0, n)
    tw = rng.uniform(0, 2 * np.pi, n)
    M = (1 / 2000.0, 3e-5, -2e-5, 1 / 1500.0)
    worst = 0.0
    for k in range(n):
        z = [lne[k], cg[k] * math.cos(th[k]), cg[k] * math.sin(th[k]), 0.3, -0.2]
        u, v = U[k] * math.cos(tw[k]), U[k] * math.sin(tw[k])
        pc = 1e-7 if k % 3 == 0 else 0.0
        d = oracle.rhs(P, z, u, v, M=M, pc=pc)
        ref, scale = mp_rhs(P, z, u, v, M, pc)
        for c in range(5):
            err = abs(F(float(d[c])) - ref[c])
            if scale[c] > 0:
                worst = max(worst, float(err / scale[c]))
            assert err <= 1e-12 * scale[c] + F("1e-300"), (k, c, float(d[c]), float(ref[c]))
    print("right-hand side: worst error relative to the error scale:", worst)


def test_right_hand_side_guards_against_exact_arithmetic():
    """the max(...) floors, the alpha cap at 500 and the us*sg == 0 branch (:215-225, :242-249, :283-284)"""
    P = default_params()
    M = (5e-4, 0.0, 0.0, 5e-4)
    for z, u, v in [([-7.0, 1e-3, 2e-3, 0, 0], 12.0, 3.0),      # alpha capped at 500, c_gp^2 < 1e-2, |c_gp| < 0.1
                    ([-5.0, 0.7, 0.7, 0, 0], 0.0, 0.0),          # calm: s2 = 0 branch, alpha = 0
                    ([-9.0, 3e-5, 4e-5, 0, 0], 10.0, -10.0),     # speed(c_gp) < 1e-4 floor
                    ([1.0, -6.0, 2.0, 0, 0], -25.0, 14.0)]:
        d = oracle.rhs(P, z, u, v, M=M)
        ref, scale = mp_rhs(P, z, u, v, M, 0.0)
        for c in range(5):
            assert abs(F(float(d[c])) - ref[c]) <= 1e-12 * scale[c] + F("1e-300"), (z, c)


def test_particle_node_maps_against_exact_arithmetic():
    """GetParticleEnergyMomentum (core_2D.jl:69-78) and GetVariablesAtVertex (:121-128), and that they invert each other"""
    rng = np.random.default_rng(2)
    for _ in range(200):
        lne, cx, cy = rng.uniform(-14, 2), rng.normal(0, 4), rng.normal(0, 4)
        ch = oracle.particle_to_charge([lne, cx, cy, 0, 0])
        e = mp.exp(F(lne))
        c2 = mp.sqrt(F(cx) ** 2 + F(cy) ** 2) ** 2
        assert close(ch[0], e) and close(ch[1], F(cx) * e / c2 / 2, scale=e / mp.sqrt(c2)) and close(ch[2], F(cy) * e / c2 / 2, scale=e / mp.sqrt(c2))
        back = oracle.vertex_to_particle(ch)
        m2 = F(float(ch[1])) ** 2 + F(float(ch[2])) ** 2
        assert abs(F(float(back[0])) - mp.log(F(float(ch[0])))) <= REL * max(abs(lne), 1)
        assert close(back[1], F(float(ch[1])) * F(float(ch[0])) / (2 * m2), scale=abs(F(cx)) + abs(F(cy)))
        assert abs(back[1] - cx) <= 1e-12 * (abs(cx) + abs(cy)) and abs(back[2] - cy) <= 1e-12 * (abs(cx) + abs(cy))


def test_grid_metric_against_exact_arithmetic():
    """TripolarGridMOM6.jl:448-459 (`cos.(angle_dx * pi / 180)` — a product rounded to float64 BEFORE the cosine, not cosd)
    and spherical_grid_corrections.jl:13 (`sign(φ)·min(sign(φ)·tand(φ), 60)/R`), from the Julia text in 50 digits, against
    oracle.grid_metric (whose sin/cos/tand are this repository's own pmath_trig.h, shared with k_grid_metric): random
    spacings, rotation angles over ±180°, latitudes over ±90° including the clamp near the poles, the equator and ±45°."""
    rng = np.random.default_rng(11)
    n = 4000
    dx = rng.uniform(500.0, 2e5, n)
    dy = rng.uniform(500.0, 2e5, n)
    ang = np.concatenate([rng.uniform(-180.0, 180.0, n - 8), [0.0, 90.0, -90.0, 180.0, 45.0, 1e-9, -1e-9, 30.0]])
    lat = np.concatenate([rng.uniform(-90.0, 90.0, n - 8), [0.0, 45.0, -45.0, 89.0, -89.0, 89.99, -89.99, 60.0]])
    Mo, pco = oracle.grid_metric(dx, dy, ang, lat)
    R = F("6.3710e6")
    worst_M = worst_pc = 0.0
    for k in range(n):
        arg = F(float(ang[k]) * math.pi / 180)            # the float64 product the reference hands to cos / sin
        ca, sa = mp.cos(arg), mp.sin(arg)
        ref = [ca / F(float(dx[k])), sa / F(float(dy[k])), -sa / F(float(dx[k])), ca / F(float(dy[k]))]
        for j in range(4):
            # on the scale of the row's larger entry: cos(90°·π/180) is 6e-17, not 0, in both
            scale = max(abs(ref[j]), 1 / F(float(max(dx[k], dy[k]))) * F("1e-3"))
            worst_M = max(worst_M, float(abs(F(float(Mo[j][k])) - ref[j]) / scale))
        ph = F(float(lat[k]))
        sg = mp.sign(ph)
        td = mp.tan(ph * mp.pi / 180)                      # tand: exact degrees
        refpc = sg * min(sg * td, F(60)) / R
        if refpc == 0:
            assert pco[k] == 0.0
        else:
            worst_pc = max(worst_pc, float(abs(F(float(pco[k])) - refpc) / abs(refpc)))
    assert worst_M < 5e-16 * 4 and worst_pc < 5e-16 * 4, (worst_M, worst_pc)
