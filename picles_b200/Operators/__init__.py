from . import TimeSteppers, core_2D  # noqa: F401
