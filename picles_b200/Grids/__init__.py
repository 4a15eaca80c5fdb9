from . import CartesianGrid, TripolarGridMOM6, mask_utils, spherical_grid_corrections  # noqa: F401
