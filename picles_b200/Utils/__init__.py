from . import WindEmulator  # noqa: F401
