"""Host mirror of `src/ParticleMesh.jl:102-146` — the grids of the one-dimensional model."""
from __future__ import annotations

import numpy as np


class OneDGrid:
    """OneDGrid(xmin, xmax, Nx): Nx nodes, Ndx = Nx - 1 cells, dx = (xmax - xmin)/Ndx (ParticleMesh.jl:102-118)."""

    def __init__(self, xmin, xmax, Nx):
        self.Nx = int(Nx)
        self.Ndx = self.Nx - 1
        self.xmin = float(xmin)
        self.xmax = float(xmax)
        self.dimx = self.xmax - self.xmin
        self.dx = self.dimx / self.Ndx

    def __repr__(self):
        return f"OneDGrid(xmin={self.xmin}, xmax={self.xmax}, Nx={self.Nx}, dx={self.dx})"


class OneDGridNotes:
    """OneDGridNotes(grid): x = collect(LinRange(0, grid.dimx, grid.Nx)) — the node coordinates start at 0 whatever
    grid.xmin is (ParticleMesh.jl:121-134).  LinRange(a, b, n)[i] = (1 - t) a + t b with t = (i - 1)/(n - 1)."""

    def __init__(self, grid: OneDGrid):
        self.Nx, self.Ndx = grid.Nx, grid.Ndx
        self.xmin, self.xmax, self.dimx, self.dx = grid.xmin, grid.xmax, grid.dimx, grid.dx
        t = np.arange(grid.Nx, dtype=np.float64) / float(grid.Nx - 1)
        self.x = (1.0 - t) * 0.0 + t * grid.dimx


def get_x(m: OneDGrid, i):
    """get_x(mesh, i), ParticleMesh.jl:143-145 (1-based i)."""
    return m.xmin + (i - 1) * m.dx
