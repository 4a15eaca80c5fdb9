"""Flatten ODESettings + ODEParameters + particle_equations constants + WaveGrowth2D
keyword arguments into the POD `picles_params_t` that crosses the C ABI."""
from __future__ import annotations

from ._abi import PiclesParams


def make_params(ODEsets, ODEsys, minimal_state, defaults=None, periodic_boundary=True, on_persist=False,
                nan_eest_rejects=False):
    """ODEsets: ParticleSystems.particle_waves_v5.ODESettings; ODEsys: ParticleSystem
    (result of particle_equations); minimal_state: [E_min, |m|²_min]; defaults: None
    ("wind_sea") or a 5-sequence (ParticleDefaults lne, c̄_x, c̄_y, x, y)."""
    P = PiclesParams()
    par = ODEsets.Parameters
    P.r_g = float(par["r_g"])
    P.C_alpha = float(par["C_α"])
    P.C_varphi = float(par.get("C_φ", 0.0))  # absent from the 1-D parameter set (r_g, C_α, C_e)
    P.C_e = float(par["C_e"])
    P.g = float(par.get("g", 9.81))
    P.p, P.q, P.n, P.e_T = float(ODEsys.p), float(ODEsys.q), float(ODEsys.n), float(ODEsys.e_T)
    P.propagation = int(bool(ODEsys.propagation))
    P.input = int(bool(ODEsys.input))
    P.dissipation = int(bool(ODEsys.dissipation))
    P.peak_shift = int(bool(ODEsys.peak_shift))
    P.direction = int(bool(ODEsys.direction))
    P.solver = ODEsets.solver_id()
    P.abstol = float(ODEsets.abstol)
    P.reltol = float(ODEsets.reltol)
    P.dt = float(ODEsets.dt)
    P.dtmin = float(ODEsets.dtmin)
    P.dtmax = float(ODEsets.total_time)  # OrdinaryDiffEq default dtmax = tspan length
    P.force_dtmin = int(bool(ODEsets.force_dtmin))
    if not ODEsets.adaptive:
        raise ValueError("adaptive=false is not supported on the B200 path")
    P.adaptive = 1
    P.maxiters = int(ODEsets.maxiters)
    P.log_energy_minimum = float(ODEsets.log_energy_minimum)
    P.log_energy_maximum = float(ODEsets.log_energy_maximum)
    P.wind_min_squared = float(ODEsets.wind_min_squared)
    P.seed_timescale = float(ODEsets.timestep)
    P.minimal_state[0] = float(minimal_state[0])
    P.minimal_state[1] = float(minimal_state[1])
    if defaults is None:
        P.has_defaults = 0
    else:
        P.has_defaults = 1
        for k in range(5):
            P.defaults[k] = float(defaults[k])
    P.periodic_boundary = int(bool(periodic_boundary))
    P.on_persist = int(bool(on_persist))
    # a NaN error estimate (overflowing trial step): False = DtNaN, exact powers; True = rejected by 1/qmin, as
    # OrdinaryDiffEq's fastpow / FastPower.fastpower make of it (include/picles_b200.h)
    P.nan_eest_rejects = int(bool(nan_eest_rejects))
    return P
