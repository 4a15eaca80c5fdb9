"""Host-side mirror of PiCLES `FetchRelations` (JONSWAP / Dulov fetch laws).

Used by the host to build model defaults (minimal_state, minimal_particle, default
particles) exactly where the reference does (WaveGrowthModels2D.jl:234-246); in-loop
re-seeding runs on the device (physics.h, windsea()).  Reference:
/root/reference/src/FetchRelations.jl:107-415.
"""
from __future__ import annotations

import math

q_x = 0.2748
A = 22.8013
xi_0x = 2.4097
u_min = 1.0


def X_tilde_from_tau(tau):  # :128-130
    return (tau / (A * xi_0x)) ** (1 / (1 - q_x))


def fₘ_from_X_tilde(U10, X_tilde, g=9.81, fgp=3.5):  # :165-167
    return fgp * (g / U10) * X_tilde ** (-0.33)


def alpha_j(U10, f_m, g=9.81):  # :184-186
    return 0.033 * (f_m * U10 / g) ** 0.67


def E_JONSWAP(f_m, alpha_j_):  # :201-203
    return 0.31 * 9.81 ** 2 * alpha_j_ * (f_m * 2 * math.pi) ** (-4)


def get_initial_windsea(U10, V10, time_scale, type="JONSWAP", particle_state=False):
    """FetchRelations.jl:314-359."""
    U_amp = math.sqrt(U10 ** 2 + V10 ** 2)
    U_amp = 0.1 if U_amp < 0.1 else U_amp
    time_scale = abs(time_scale)
    tau = 9.81 * time_scale / abs(U_amp)
    X_tilde_ = X_tilde_from_tau(tau)
    f_m_ = fₘ_from_X_tilde(U_amp, X_tilde_)
    alpha_j_ = alpha_j(U_amp, f_m_)
    if type == "JONSWAP":
        E_ = E_JONSWAP(f_m_, alpha_j_)
        Hs_ = 4 * math.sqrt(E_)
        f_peak = f_m_ * 9.81 / U_amp
    elif type == "PM":
        f_peak = 0.816 * 9.81 / (2 * math.pi * U_amp)
        Hs_ = 0.0246 * U_amp ** 2
        E_ = (Hs_ / 4) ** 2
    else:
        raise ValueError(type)
    T_bar = 0.9 * (1 / f_peak)
    cg_bar_amp = 9.81 * T_bar / (4 * math.pi)
    cg_bar_x = cg_bar_amp * U10 / U_amp
    cg_bar_y = cg_bar_amp * V10 / U_amp
    if particle_state:
        return [math.log(E_), cg_bar_x, cg_bar_y, 0.0, 0.0]
    mom_x = (U10 / U_amp) * E_ / (2 * cg_bar_amp)
    mom_y = (V10 / U_amp) * E_ / (2 * cg_bar_amp)
    return {"E": E_, "lne": math.log(E_), "Hs": Hs_, "cg_bar_x": cg_bar_x, "cg_bar_y": cg_bar_y,
            "cg_bar": cg_bar_amp, "f_peak": f_peak, "T_bar": T_bar, "X_tilde": X_tilde_, "m_x": mom_x, "m_y": mom_y}


def _nz(x):
    # the reference draws rand_sign() for an exactly-zero component (FetchRelations.jl:365,
    # 382-383), which is non-deterministic; the B200 path fixes the sign to +1.
    return 1.0 if x == 0 else x


def MinimalWindsea(U10, V10, time_scale, type="JONSWAP"):  # :381-386
    U10, V10 = _nz(U10), _nz(V10)
    Uamp = math.sqrt(U10 ** 2 + V10 ** 2)
    return get_initial_windsea(u_min * U10 / Uamp, u_min * V10 / Uamp, time_scale, type=type)


def MinimalParticle(U10, V10, time_scale, type="JONSWAP"):  # :401-404
    w = MinimalWindsea(U10, V10, time_scale, type)
    return [math.log(w["E"]), w["cg_bar_x"], w["cg_bar_y"], 0, 0]


def MinimalState(U10, V10, time_scale, type="JONSWAP"):  # :412-415
    w = MinimalWindsea(U10, V10, time_scale, type)
    return [w["E"], w["m_x"] ** 2 + w["m_y"] ** 2]
