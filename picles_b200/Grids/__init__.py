from . import CartesianGrid, SphericalGrid, TripolarGridMOM6, mask_utils, spherical_grid_corrections  # noqa: F401
