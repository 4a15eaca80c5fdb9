"""Host mirror of `src/Grids/CartesianGrid.jl` (TwoDCartesianGridStatistics / Mesh,
ProjetionKernel).  Arrays are numpy (Nx, Ny), Fortran-ordered, indexed [i, j] like the
reference; they reach the device as i-fastest planes."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..Architectures import N_NonPeriodic, N_Periodic
from .mask_utils import make_boundaries


@dataclass
class MeshData:
    """`grid.data` StructArray: component arrays with shape (Nx, Ny)."""

    x: np.ndarray
    y: np.ndarray
    mask: np.ndarray
    angle_dx: np.ndarray = None
    dx: np.ndarray = None
    dy: np.ndarray = None
    area: np.ndarray = None


class TwoDCartesianGridStatistics:
    """CartesianGrid.jl:26-64."""

    def __init__(self, xmin, xmax, Nx, ymin, ymax, Ny, angle=0.0, periodic_boundary=(False, False)):
        self.xmin, self.xmax, self.ymin, self.ymax = float(xmin), float(xmax), float(ymin), float(ymax)
        self.dimx = self.xmax - self.xmin
        self.dimy = self.ymax - self.ymin
        self.Ndx = int(Nx) - 1
        self.Ndy = int(Ny) - 1
        self.dx = self.dimx / self.Ndx
        self.dy = self.dimy / self.Ndy
        self.area = self.dx * self.dy
        self.Nx = N_Periodic(Nx) if periodic_boundary[0] else N_NonPeriodic(Nx)
        self.Ny = N_Periodic(Ny) if periodic_boundary[1] else N_NonPeriodic(Ny)
        self.angle_dx = float(angle)


def ProjetionKernel(stats):
    """CartesianGrid.jl:115-136 (sic) — M = [1/dx 0; 0 1/dy], or the rotated variant the
    reference writes WITHOUT a minus sign (SURVEY B-8).  Returned row-major (M11,M12,M21,M22)."""
    if stats.angle_dx == 0.0:
        return np.array([1 / stats.dx, 0.0, 0.0, 1 / stats.dy])
    cosa = np.cos(stats.angle_dx * np.pi / 180)
    sina = np.sin(stats.angle_dx * np.pi / 180)
    return np.array([cosa / stats.dx, sina / stats.dy, sina / stats.dx, cosa / stats.dy])


class TwoDCartesianGridMesh:
    """CartesianGrid.jl:67-112.  Call forms of the reference:
    TwoDCartesianGridMesh(dimx, nx, dimy, ny; angle, periodic_boundary)
    TwoDCartesianGridMesh(xmin, xmax, Nx, ymin, ymax, Ny; mask, angle, periodic_boundary)"""

    def __init__(self, *args, mask=None, angle=0.0, periodic_boundary=(False, False), total_mask=None):
        if len(args) == 4:
            dimx, nx, dimy, ny = args
            xmin, xmax, Nx, ymin, ymax, Ny = 0.0, dimx, nx, 0.0, dimy, ny
        elif len(args) == 6:
            xmin, xmax, Nx, ymin, ymax, Ny = args
        else:
            raise TypeError("TwoDCartesianGridMesh(dimx, nx, dimy, ny) or (xmin, xmax, Nx, ymin, ymax, Ny)")
        self.stats = TwoDCartesianGridStatistics(xmin, xmax, Nx, ymin, ymax, Ny, angle=angle,
                                                 periodic_boundary=periodic_boundary)
        x = np.linspace(self.stats.xmin, self.stats.xmax, int(Nx))
        y = np.linspace(self.stats.ymin, self.stats.ymax, int(Ny))
        XX = np.asfortranarray(np.broadcast_to(x[:, None], (int(Nx), int(Ny))))
        YY = np.asfortranarray(np.broadcast_to(y[None, :], (int(Nx), int(Ny))))
        if mask is None:
            mask = np.ones(XX.shape, dtype=bool)
        if total_mask is None:
            total_mask = make_boundaries(mask, self.stats.Nx, self.stats.Ny)
        self.data = MeshData(x=XX, y=YY, mask=np.asfortranarray(total_mask))
        self.ProjetionKernel = ProjetionKernel
        self.PropagationCorrection = None  # SphericalPropagationCorrection_dummy

    # what the engine needs
    def device_metric(self):
        return dict(M=None, M_const=ProjetionKernel(self.stats), pc=None)
