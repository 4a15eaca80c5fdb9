"""Host mirror of `src/Grids/SphericalGrid.jl`: the regular longitude/latitude grid
(`TwoDSphericalGridStatistics` :99-138, `TwoDSphericalGridMesh` :156-210, per-node
`ProjetionKernel(Gi, stats)` :229-239) with the great-circle term of
`spherical_grid_corrections.jl`.  Arrays are (Nx, Ny), indexed [i, j] like the reference.

The device path needs nothing new for this grid: the per-node kernel planes and the great-circle
coefficient go through `picles_set_grid` (M, pc_coef), the same planes the oracle is given.

Kept as the reference has it: the kernel takes its "cos(lat)" from
`cos(Gi.dy * pi / 180)` — `dy` is the meridional spacing in METRES there, not the latitude — and
the matrix form `ProjetionKernel(Gdata::StructArray)` (:207-226) reads an undefined `R`, so only
the per-node form exists here (it is the one SeedParticle calls, core_2D.jl:477).
"""
from __future__ import annotations

import numpy as np

from ..Architectures import N_NonPeriodic, N_Periodic
from .CartesianGrid import MeshData
from .mask_utils import make_boundaries
from .spherical_grid_corrections import SphericalPropagationCorrection

R_EARTH = 6371.0e3  # SphericalGrid.jl:58,79


def cal_dx_degree(XX):
    """SphericalGrid.jl:26-32: centred differences, one-sided at the ends (axis 1 of the reference = i)."""
    dx = np.zeros(XX.shape)
    dx[1:-1, :] = (XX[2:, :] - XX[:-2, :]) / 2
    dx[0, :] = XX[1, :] - XX[0, :]
    dx[-1, :] = XX[-1, :] - XX[-2, :]
    return dx


def cal_dy_degree(YY):
    """SphericalGrid.jl:34-40."""
    dy = np.zeros(YY.shape)
    dy[:, 1:-1] = (YY[:, 2:] - YY[:, :-2]) / 2
    dy[:, 0] = YY[:, 1] - YY[:, 0]
    dy[:, -1] = YY[:, -1] - YY[:, -2]
    return dy


def cal_dx_meters(XX, YY):
    """SphericalGrid.jl:56-61."""
    return cal_dx_degree(XX) * np.pi / 180 * (R_EARTH * np.cos(YY * np.pi / 180))


def cal_dy_meters(YY):
    """SphericalGrid.jl:78-81."""
    return cal_dy_degree(YY) * np.pi / 180 * R_EARTH


class TwoDSphericalGridStatistics:
    """SphericalGrid.jl:99-138."""

    def __init__(self, xmin, xmax, Nx, ymin, ymax, Ny, mask_value=1, angle=0.0, periodic_boundary=(False, False)):
        self.xmin, self.xmax, self.ymin, self.ymax = float(xmin), float(xmax), float(ymin), float(ymax)
        self.dimx = self.xmax - self.xmin
        self.dimy = self.ymax - self.ymin
        self.Ndx, self.Ndy = int(Nx) - 1, int(Ny) - 1
        self.Nx = N_Periodic(Nx) if periodic_boundary[0] else N_NonPeriodic(Nx)
        self.Ny = N_Periodic(Ny) if periodic_boundary[1] else N_NonPeriodic(Ny)
        self.dx_deg = self.dimx / self.Ndx
        self.dy_deg = self.dimy / self.Ndy
        self.angle_dx = float(angle)
        self.mask_value = int(mask_value)


def ProjetionKernel(Gi_dx, Gi_dy):
    """ProjetionKernel(Gi::NamedTuple, stats), SphericalGrid.jl:229-239, for arrays of nodes:
    [1/(cos(dy*pi/180)*dx) 0; 0 1/dy] as planes (M11, M12, M21, M22)."""
    cos_lat = np.cos(Gi_dy * np.pi / 180)
    z = np.zeros_like(Gi_dx)
    return np.stack([1.0 / (cos_lat * Gi_dx), z, z, 1.0 / Gi_dy])


class TwoDSphericalGridMesh:
    """TwoDSphericalGridMesh(xmin, xmax, Nx, ymin, ymax, Ny; mask, angle, periodic_boundary),
    SphericalGrid.jl:200-204; data fields x, y (degrees), dx, dy (metres), area, mask."""

    def __init__(self, xmin, xmax, Nx, ymin, ymax, Ny, mask=None, angle=0.0, periodic_boundary=(False, False),
                 total_mask=None):
        self.stats = TwoDSphericalGridStatistics(xmin, xmax, Nx, ymin, ymax, Ny, angle=angle,
                                                 periodic_boundary=periodic_boundary)
        st = self.stats
        # collect(range(xmin, stop=xmax, step=dx_deg)): Nx points (the reference relies on the range hitting xmax)
        x = st.xmin + st.dx_deg * np.arange(int(Nx))
        y = st.ymin + st.dy_deg * np.arange(int(Ny))
        XX = np.asfortranarray(np.broadcast_to(x[:, None], (int(Nx), int(Ny))))
        YY = np.asfortranarray(np.broadcast_to(y[None, :], (int(Nx), int(Ny))))
        dx = cal_dx_meters(XX, YY)
        dy = cal_dy_meters(YY)
        if mask is None:
            mask = np.ones(XX.shape, dtype=bool)
        if total_mask is None:
            total_mask = make_boundaries(mask, st.Nx, st.Ny)
        self.data = MeshData(x=XX, y=YY, mask=np.asfortranarray(total_mask), dx=np.asfortranarray(dx),
                             dy=np.asfortranarray(dy), area=np.asfortranarray(dx * dy))
        self.ProjetionKernel = ProjetionKernel
        self.PropagationCorrection = SphericalPropagationCorrection

    def device_metric(self):
        """per-node kernel planes and great-circle coefficient, host-evaluated (numpy) and handed to
        the library as they are (picles_set_grid M / pc_coef)"""
        return dict(M=ProjetionKernel(self.data.dx, self.data.dy), M_const=None,
                    pc=SphericalPropagationCorrection(self.data.y))
