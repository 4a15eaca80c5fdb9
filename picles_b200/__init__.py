"""picles_b200 — B200-native implementation of the PiCLES per-timestep particle-in-cell
path (adaptive-RK particle advance → ParticleToNode projection → NodeToParticle remesh)
behind the reference's WaveGrowth2D / Simulation / run! API.

The compute lives in libpicles_b200.so (hand-written sm_100a CUDA kernels behind the
C ABI in include/picles_b200.h); this package is the host-side mirror of the reference's
Julia interface.  There is no CPU fallback.
"""
from . import FetchRelations  # noqa: F401
from ._abi import PiclesError  # noqa: F401

__version__ = "0.1.0"
