"""Host mirror of `src/Grids/spherical_grid_corrections.jl`: the great-circle coefficient
that the advance kernel multiplies with c̄_x (PropagationCorrection(c̄_x) = c̄_x * coef)."""
from __future__ import annotations

import numpy as np


def SphericalPropagationCorrection(phi, R=6.3710e6):
    """spherical_grid_corrections.jl:3-21 — coefficient sign(φ)·min(sign(φ)·tand(φ), 60)/R
    (φ latitude in degrees).  Returns the coefficient array; the closure of the reference
    is `cg_x_bar -> cg_x_bar * coefficient`."""
    phi = np.asarray(phi, dtype=np.float64)
    sgn = np.sign(phi)
    return (sgn * np.minimum(sgn * np.tan(np.deg2rad(phi)), 60.0)) / R


def SphericalPropagationCorrection_dummy(phi=None):
    """spherical_grid_corrections.jl:54-57 — x -> 0.0 (Cartesian grids)."""
    return None
