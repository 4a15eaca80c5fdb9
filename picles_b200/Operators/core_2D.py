"""Host mirror of the small value types / conversions of `src/Operators/core_2D.jl` that
scripts use around the stepping path (the per-particle versions run on the device)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class ParticleDefaults:
    """core_2D.jl:40-58."""

    lne: float
    c̄_x: float
    c̄_y: float
    x: float = 0.0
    y: float = 0.0

    def as_list(self):
        return [self.lne, self.c̄_x, self.c̄_y, self.x, self.y]


def speed(x, y):
    return np.sqrt(x ** 2 + y ** 2)


def GetParticleEnergyMomentum(z0):
    """core_2D.jl:69-78 — (e, m_x, m_y) from [lne, c̄_x, c̄_y, x, y]."""
    if isinstance(z0, ParticleDefaults):
        z0 = z0.as_list()
    lne, cx, cy = z0[0], z0[1], z0[2]
    e = np.exp(lne)
    c_speed = speed(cx, cy)
    return np.array([e, cx * e / c_speed ** 2 / 2, cy * e / c_speed ** 2 / 2])


def GetVariablesAtVertex(i_State, x, y):
    """core_2D.jl:121-128."""
    e, m_x, m_y = i_State
    m_amp = speed(m_x, m_y)
    return np.array([np.log(e), m_x * e / (2 * m_amp ** 2), m_y * e / (2 * m_amp ** 2), x, y])


def GetGroupVelocity(i_State):
    """core_2D.jl:138-147 — (c_x, c_y) fields from a State array (Nx, Ny, 3)."""
    e, m_x, m_y = i_State[:, :, 0], i_State[:, :, 1], i_State[:, :, 2]
    m_amp = speed(m_x, m_y)
    with np.errstate(divide="ignore", invalid="ignore"):
        c_x = m_x * e / (2 * m_amp ** 2)
        c_y = m_y * e / (2 * m_amp ** 2)
    return dict(c_x=c_x, c_y=c_y)


def check_boundary_point(imesh, periodic_boundary):
    """core_2D.jl:360-366."""
    return imesh == 2 if periodic_boundary else imesh >= 2
