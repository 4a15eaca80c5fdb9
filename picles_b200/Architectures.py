"""Architecture and boundary types.

The reference's `src/Architectures.jl:13-69` holds only abstract types; nothing dispatches
on a device.  The B200 path adds the concept the north-star asks for: an architecture
object selected with `WaveGrowth2D(...; architecture=B200())`.  `CPU()` exists only so the
keyword reads like the reference would; this package has no CPU compute path and raises
if asked to run on it.

Boundary types mirror `src/custom_structures.jl:51-61`.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple


class AbstractArchitecture:
    pass


@dataclass(frozen=True)
class CPU(AbstractArchitecture):
    """Placeholder: the Julia reference is the CPU implementation."""


@dataclass(frozen=True)
class B200(AbstractArchitecture):
    """One or more B200 GPUs of one node; `devices` are CUDA ordinals.  With more than one
    device the grid is split in contiguous y-strips, one per device/process
    (picles_b200.distributed)."""

    devices: Tuple[int, ...] = (0,)
    #: wind levels staged per model step (2..5).  The reference calls the wind closures at every
    #: Runge-Kutta stage time; here they are evaluated on the host mesh at `wind_levels` equally
    #: spaced times of each step and interpolated in time on the device (2: linear, exact for
    #: steady winds; 5: within 1e-9 of the closure for the time-varying winds of the reference's
    #: scripts, tests/test_wind_levels.py).
    wind_levels: int = 2
    #: a trial step whose stages overflow (EEst = NaN): False = the integrator ends with DtNaN (exact powers in the PI
    #: controller, the default); True = rejected by 1/qmin, as OrdinaryDiffEq's fastpow / FastPower.fastpower make of it
    #: (include/picles_b200.h: picles_params_t::nan_eest_rejects)
    nan_eest_rejects: bool = False

    def __post_init__(self):
        if not 2 <= int(self.wind_levels) <= 5:
            raise ValueError("wind_levels must be between 2 and 5")


class AbstractBoundary:
    """Integer-like axis length tagged with its boundary rule (`AbstractBoundary <: Integer`)."""

    code = -1

    def __init__(self, N: int):
        self.N = int(N)

    def __int__(self):
        return self.N

    def __index__(self):
        return self.N

    def __repr__(self):
        return f"Int={self.N} {type(self).__name__}"


class N_NonPeriodic(AbstractBoundary):
    code = 0


class N_Periodic(AbstractBoundary):
    code = 1


class N_TripolarNorth(AbstractBoundary):
    code = 2
