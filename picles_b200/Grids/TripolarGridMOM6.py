"""Host mirror of `src/Grids/TripolarGridMOM6.jl`: MOM6 super-grid → T-point mesh, per-node
metric (dx, dy, angle) and the tripolar-north boundary type.

The reference reads `ocean_hgrid_221123.nc` / `ocean_topo_tx2_3v2_240501.nc`, which are not
in the checkout (`.MISSING_LARGE_BLOBS`); NetCDF ingestion is out of scope here.  The
functions below take the super-grid arrays directly (any reader can supply them) and
`synthetic_supergrid` builds an analytic stand-in of the same shapes.

Arrays are numpy, indexed [i, j] like the Julia arrays."""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

from ..Architectures import N_NonPeriodic, N_Periodic, N_TripolarNorth
from .CartesianGrid import MeshData
from .mask_utils import make_boundaries, mask_circle_
from .spherical_grid_corrections import SphericalPropagationCorrection


def extract_grid_points(x, y, angle_dx, k, mask=None):
    """TripolarGridMOM6.jl:42-103 — T/U/V/Q point coordinates from the super-grid."""
    if x.shape != y.shape:
        raise ValueError("x and y have different shapes")
    khalf = int(k / 2)
    cx, cy = slice(khalf, None, k), slice(khalf, None, k)   # khalf+1:k:end
    qx, qy = slice(0, None, k), slice(0, None, k)           # 1:k:end
    if mask is not None:
        if k == 2:
            mask = mask == 1
        elif k in (4, 6, 8):
            mask = mask[::int(k / 2), ::int(k / 2)] == 1
        else:
            raise ValueError("k must be 2, 4, 6, or 8")
    return SimpleNamespace(
        Tpoint=SimpleNamespace(lon=x[cx, cy], lat=y[cx, cy]),
        u=SimpleNamespace(lon=x[qx, cy], lat=y[qx, cy]),
        v=SimpleNamespace(lon=x[cx, qy], lat=y[cx, qy]),
        Qpoint=SimpleNamespace(lon=x[qx, qy], lat=y[qx, qy]),
        angle=angle_dx[cx, cy], k=k, khalf=khalf, mask=mask)


def calculate_distances(area, dx, dy, k, khalf):
    """TripolarGridMOM6.jl:203-264 — T-cell areas and the x/y distances between u/v, q and
    t points (sums of k super-grid edge lengths), incl. the north-seam fix-up of dyCv."""
    shp_a = (area.shape[0] // k, area.shape[1] // k)
    shp_x = (dx.shape[0] // k, dx.shape[1] // k)
    shp_y = (dy.shape[0] // k, dy.shape[1] // k)
    tarea = np.zeros(shp_a)
    for i in range(k):
        for j in range(k):
            tarea += area[i::k, j::k]
    dxt = np.zeros(shp_x)
    dyt = np.zeros(shp_y)
    dxCv = np.zeros(shp_x)
    dyCu = np.zeros(shp_y)
    dxCu = np.zeros(shp_x)
    dyCv = np.zeros(shp_y)
    dxt[...] = sum(dx[j::k, khalf::k] for j in range(k))
    dyt[...] = sum(dy[khalf::k, i::k] for i in range(k))
    dxCv[...] = sum(dx[j::k, k::k] for j in range(k))
    dyCu[...] = sum(dy[k::k, i::k] for i in range(k))
    dxr = np.roll(dx, -khalf, axis=0)
    for j in range(k):
        dxCu += dxr[j::k, khalf::k]
    dyr = np.roll(dy, -khalf, axis=1).copy()
    dyr[:, -1] = dyr[::-1, -4]    # dyr[:, end]   = dyr[end:-1:1, end-3]
    dyr[:, -2] = dyr[::-1, -3]    # dyr[:, end-1] = dyr[end:-1:1, end-2]
    for i in range(k):
        dyCv += dyr[khalf::k, i::k]
    return SimpleNamespace(Tarea=tarea, UVdist=SimpleNamespace(dxt=dxt, dyt=dyt),
                           Qdist=SimpleNamespace(dxCv=dxCv, dyCu=dyCu), Tdist=SimpleNamespace(dxCu=dxCu, dyCv=dyCv))


class MOM6GridStatistic:
    """TripolarGridMOM6.jl:288-329 — Nx periodic, Ny tripolar-north unless periodic."""

    def __init__(self, Grid, mask_value=1, file="unknown", periodic_boundary=(True, False)):
        Nx, Ny = Grid.Tpoint.lon.shape
        self.file = file
        self.Ndx, self.Ndy = Nx - 1, Ny - 1
        self.Nx = N_Periodic(Nx) if periodic_boundary[0] else N_NonPeriodic(Nx)
        self.Ny = N_Periodic(Ny) if periodic_boundary[1] else N_TripolarNorth(Ny)
        self.xmin, self.xmax = float(Grid.Tpoint.lon.min()), float(Grid.Tpoint.lon.max())
        self.ymin, self.ymax = float(Grid.Tpoint.lat.min()), float(Grid.Tpoint.lat.max())
        self.dimx, self.dimy = self.xmax - self.xmin, self.ymax - self.ymin
        self.mask_value = mask_value


def TripolarGrid_mask_pols_(mask, Nx, Ny, lons, lats, dx, radius_deg):
    """TripolarGrid_mask_pols!, TripolarGridMOM6.jl:469-486."""
    for pp_ij in [(1, Ny.N), (Nx.N, Ny.N), (int(np.round(Nx.N / 2)), Ny.N)]:
        mask_circle_(mask, lons, lats, pp_ij, radius_deg)
    dx_deg = float(np.mean(dx)) / 110e3
    Ny_mask = int(math.ceil(radius_deg / dx_deg))
    mask[:, :Ny_mask] = False
    return mask


def ProjetionKernel(data):
    """TripolarGridMOM6.jl:448-459 — per node [cosα/dx sinα/dy; -sinα/dx cosα/dy]; returns
    4 planes (M11, M12, M21, M22) of shape (Nx, Ny)."""
    cosa = np.cos(data.angle_dx * np.pi / 180)
    sina = np.sin(data.angle_dx * np.pi / 180)
    return np.stack([cosa / data.dx, sina / data.dy, -sina / data.dx, cosa / data.dy])


class MOM6GridMesh:
    """TripolarGridMOM6.jl:332-432 (constructor from extracted points + distances)."""

    def __init__(self, G, GA, mask=None, file="unknown", total_mask=None, mask_radius=3):
        self.stats = MOM6GridStatistic(G, file=file)
        if mask is None:
            mask = np.ones(G.Tpoint.lon.shape, dtype=bool)
            TripolarGrid_mask_pols_(mask, self.stats.Nx, self.stats.Ny, G.Tpoint.lon, G.Tpoint.lat, GA.Tdist.dyCv,
                                    mask_radius)
        elif mask.shape != G.Tpoint.lon.shape:
            raise ValueError("Mask size must be the same as the grid size")
        if total_mask is None:
            total_mask = make_boundaries(mask, self.stats.Nx, self.stats.Ny)
        F = np.asfortranarray
        self.data = MeshData(x=F(G.Tpoint.lon), y=F(G.Tpoint.lat), angle_dx=F(G.angle), dx=F(GA.Tdist.dxCu),
                             dy=F(GA.Tdist.dyCv), area=F(GA.Tarea), mask=F(total_mask))
        self.ProjetionKernel = ProjetionKernel
        self.PropagationCorrection = SphericalPropagationCorrection

    def device_metric(self):
        """M / pc: host evaluation of the reference formulas (numpy); raw: the mesh planes the
        library turns into the same quantities on the device (picles_set_grid_metric), which is
        what the B200 model uses."""
        return dict(M=ProjetionKernel(self.data), M_const=None, pc=SphericalPropagationCorrection(self.data.y),
                    raw=dict(dx=self.data.dx, dy=self.data.dy, angle_dx=self.data.angle_dx, lat=self.data.y,
                             R_earth=6.3710e6))


def synthetic_supergrid(nx, ny, k=2, lat_min=-78.0, lat_max=89.5, cap_lat=65.0, R=6.371e6):
    """Analytic stand-in for a MOM6 super-grid (shapes as in ocean_hgrid: x,y,angle_dx at the
    (k*nx+1, k*ny+1) super-grid vertices, dx (k*nx, k*ny+1), dy (k*nx+1, k*ny), area (k*nx, k*ny)).
    Regular lat-lon south of `cap_lat`; north of it the grid lines rotate smoothly (angle_dx up to
    ±40°) to imitate the bipolar cap."""
    nxs, nys = k * nx, k * ny
    lon = -280.0 + 360.0 * np.arange(nxs + 1) / nxs
    lat = lat_min + (lat_max - lat_min) * np.arange(nys + 1) / nys
    X, Y = np.meshgrid(lon, lat, indexing="ij")
    cap = np.clip((Y - cap_lat) / (90.0 - cap_lat), 0.0, 1.0)
    angle = 40.0 * cap * np.sin(np.deg2rad(2.0 * (X + 280.0)))
    dlon = np.deg2rad(360.0 / nxs)
    dlat = np.deg2rad((lat_max - lat_min) / nys)
    latv = 0.5 * (Y[:-1, :] + Y[1:, :])
    dx = np.maximum(R * np.cos(np.deg2rad(latv)) * dlon, 500.0)         # (nxs, nys+1)
    dy = np.full((nxs + 1, nys), R * dlat)                               # (nxs+1, nys)
    area = 0.5 * (dx[:, :-1] + dx[:, 1:]) * dy[:-1, :]                   # (nxs, nys)
    return dict(x=X, y=Y, dx=dx, dy=dy, area=area, angle_dx=angle)


def synthetic_MOM6GridMesh(nx, ny, k=2, mask=None, mask_radius=5, **kw):
    """MOM6GridMesh(GridFile, k) with the file replaced by `synthetic_supergrid`."""
    sg = synthetic_supergrid(nx, ny, k=k, **kw)
    G = extract_grid_points(sg["x"], sg["y"], sg["angle_dx"], k)
    GA = calculate_distances(sg["area"], sg["dx"], sg["dy"], G.k, G.khalf)
    return MOM6GridMesh(G, GA, mask=mask, file="synthetic", mask_radius=mask_radius)
