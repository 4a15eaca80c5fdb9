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
