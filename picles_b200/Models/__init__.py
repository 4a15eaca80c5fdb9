from . import WaveGrowthModels2D  # noqa: F401
