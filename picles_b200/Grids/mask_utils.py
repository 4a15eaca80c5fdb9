"""Host mirror of `src/Grids/mask_utils.jl` — integer masks, bit-exact with the reference.

Arrays are indexed [i, j] like the Julia (Nx, Ny) arrays (numpy shape (Nx, Ny))."""
from __future__ import annotations

import numpy as np

from ..Architectures import N_NonPeriodic


def interior_boundary(mask):
    """mask_utils.jl:14-22 — land cells (mask==0) with an ocean neighbour; `circshift`
    wraps around, exactly as in the reference."""
    mask = np.asarray(mask, dtype=bool)
    bmask = np.zeros(mask.shape, dtype=int)
    for dims in [(1, 0), (-1, 0), (0, 1), (0, -1)]:
        bmask += np.roll(mask, dims, axis=(0, 1)) & ~mask
    return bmask != 0


def make_boundaries(mask, Nx, Ny):
    """mask_utils.jl:38-55 — total mask: 0 land, 1 ocean, 2 land boundary, 3 grid boundary."""
    mask = np.asarray(mask, dtype=bool)
    bmask = interior_boundary(mask)
    total_mask = mask.astype(np.int64) + 2 * bmask.astype(np.int64)
    if isinstance(Nx, N_NonPeriodic):
        total_mask[0, :] = 3
        total_mask[-1, :] = 3
    if isinstance(Ny, N_NonPeriodic):
        total_mask[:, 0] = 3
        total_mask[:, -1] = 3
    return total_mask


def _findall(cond):
    """Julia `findall` on a matrix: CartesianIndices in column-major order (i fastest),
    returned as an (n, 2) array of 1-based (i, j)."""
    jj, ii = np.nonzero(np.asarray(cond).T)
    return np.stack([ii + 1, jj + 1], axis=1)


def make_boundary_lists(total_mask):
    """mask_utils.jl:71-82."""
    total_mask = np.asarray(total_mask)
    return dict(ocean=_findall(total_mask == 1), land_boundary=_findall(total_mask == 2),
                grid_boundary=_findall(total_mask == 3))


def mask_circle_(mask, xx, yy, pp_ij, radius):
    """mask_circle!(mask, xx, yy, pp_ij, radius), mask_utils.jl:125-139; pp_ij 1-based."""
    px, py = xx[pp_ij[0] - 1, pp_ij[1] - 1], yy[pp_ij[0] - 1, pp_ij[1] - 1]
    inside = (xx - px) ** 2 + (yy - py) ** 2 < radius ** 2
    mask[inside] = False
    return mask
