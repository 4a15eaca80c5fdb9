"""Host-side mirror of PiCLES `ParticleSystems.particle_waves_v5` (settings + constants).

Only the configuration objects live here; the right-hand side itself is evaluated by
the sm_100a advance kernel (picles_b200/csrc/physics.h, rhs()).  Reference:
/root/reference/src/ParticleSystems/particle_waves_v5.jl.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, Callable, Optional

SOLVERS = {"Tsit5": 0, "AutoTsit5": 2, "AutoTsit5(Rosenbrock23())": 2, "DP5": 1}


def magic_fractions(q: float = -1 / 4.0):
    """particle_waves_v5.jl:87-92 — universal exponent relations [p, q, n]."""
    p = (-1 - 10 * q) / 2
    n = 2 * q / (p + 4 * q)
    return [p, q, n]


@dataclass
class IDConstants:
    """particle_waves_v5.jl:107-128."""

    c_D: float
    c_β: float
    c_e: float
    c_alpha: float
    r_w: float
    C_e: float
    γ: float
    p: float
    q: float
    n: float

    @classmethod
    def make(cls, r_g=0.85, c_D=2e-3, c_β=4e-2, c_e=1.3e-6, c_alpha=11.8, r_w=2.35, q=-1 / 4):
        p = (-1 - 10 * q) / 2
        n = 2 * q / (p + 4 * q)
        C_e = r_w * c_β * c_D / r_g
        γ = 1 - (p - q) / (c_alpha ** 4 * C_e * 2)
        return cls(c_D, c_β, c_e, c_alpha, r_w, C_e, γ, p, q, n)


@dataclass
class ScgConstants:
    """particle_waves_v5.jl:154-162."""

    C_alpha: float = -1.41
    C_varphi: float = 1.81e-5


def ODEParameters(r_g=0.85, q=-0.25, g=9.81):
    """particle_waves_v5.jl:184-196 → (parset, Const_ID, Const_Scg)."""
    Const_ID = IDConstants.make(r_g=r_g, q=q)
    Const_Scg = ScgConstants()
    parset = dict(r_g=r_g, C_α=Const_Scg.C_alpha, C_φ=Const_Scg.C_varphi, C_e=Const_ID.C_e, g=g)
    return parset, Const_ID, Const_Scg


def e_T_func(γ, p, q, n, c_β=2.16e-4, c_D=2e-3, c_e=1.3e-6, c_α=11.8):
    """particle_waves_v5.jl:271 (eq. A2.4 in Kudr. 2021 2D)."""
    return math.sqrt(c_e * c_α ** (-p / q) / (γ * c_β * c_D) ** (1 / n))


@dataclass
class ParticleSystem:
    """What `particle_equations(u, v; γ, q, ...)` returns in the reference is a closure
    (particle_waves_v5.jl:382-563); here it is the descriptor the device kernel needs:
    the derived constants (p, n, e_T) and the term switches."""

    u_wind: Callable
    v_wind: Callable
    γ: float
    q: float
    p: float
    n: float
    e_T: float
    propagation: bool = True
    input: bool = True
    dissipation: bool = True
    peak_shift: bool = True
    direction: bool = True


_ONE_D = object()  # marker: particle_equations(u_wind; ...) with a single forcing field = the 1-D system


def particle_equations(u_wind, v_wind=_ONE_D, γ=0.88, q=-1 / 4.0, IDConstants_=None, propagation=True, input=True,
                       dissipation=True, peak_shift=True, direction=True, debug_output=False, static=False):
    """particle_waves_v5.jl:382-395 (two forcing fields: the 2-D system) and :584-650 (one forcing field: the 1-D
    system [lne, c̄_x, x], no directional term).  `debug_output`/`static` select host-only variants in the
    reference and are not part of the stepping path."""
    if debug_output or static:
        raise NotImplementedError("debug_output/static variants are outside the B200 path")
    idc = IDConstants_ if IDConstants_ is not None else IDConstants.make()
    p, q, n = magic_fractions(q)
    e_T = e_T_func(γ, p, q, n, c_β=idc.c_β, c_D=idc.c_D, c_e=idc.c_e, c_α=idc.c_alpha)
    if v_wind is _ONE_D:
        return ParticleSystem(u_wind, None, γ, q, p, n, e_T, propagation, input, dissipation, peak_shift, False)
    return ParticleSystem(u_wind, v_wind, γ, q, p, n, e_T, propagation, input, dissipation, peak_shift, direction)


@dataclass
class ODESettings:
    """particle_waves_v5.jl:34-75 (same field names and defaults)."""

    Parameters: dict
    log_energy_minimum: float
    saving_step: float
    timestep: float
    total_time: float
    log_energy_maximum: float = math.log(17)
    wind_min_squared: float = 4.0
    solver: Any = "AutoTsit5(Rosenbrock23())"
    abstol: float = 1e-4
    reltol: float = 1e-3
    maxiters: int = int(1e4)
    adaptive: bool = True
    dt: float = 60 * 6
    dtmin: float = 60 * 5
    force_dtmin: bool = False
    callbacks: Optional[Any] = None
    save_everystep: bool = False

    def solver_id(self) -> int:
        name = self.solver if isinstance(self.solver, str) else getattr(self.solver, "__name__", str(self.solver))
        if name not in SOLVERS:
            raise ValueError(f"solver {name!r} is not available on the B200 path (Tsit5 / AutoTsit5 / DP5)")
        return SOLVERS[name]
