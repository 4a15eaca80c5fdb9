from . import particle_waves_v5  # noqa: F401
