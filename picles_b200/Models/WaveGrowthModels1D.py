"""`WaveGrowth1D` — the model container of `src/Models/WaveGrowthModels1D.jl:28-197` with the same keyword
constructor, backed by one picles1d handle on a B200.  `time_step!` / `init_particles!` for it are
`time_step_1D` / `init_particles_1D` below (TimeSteppers.jl:51-92, run.jl:268-302); `Simulations.run` accepts the
model as it is."""
from __future__ import annotations

import numpy as np

from .. import FetchRelations
from ..Architectures import B200, CPU
from ..ParticleMesh import OneDGrid, OneDGridNotes
from ..params import make_params
from .WaveGrowthModels2D import Clock


def eval_wind_1d(f, x, t):
    try:
        out = np.asarray(f(x, t), dtype=np.float64)
        if out.shape == ():
            return np.full(x.shape, float(out))
        if out.shape == x.shape:
            return np.ascontiguousarray(out)
    except Exception:
        pass
    return np.array([float(f(float(a), t)) for a in x], dtype=np.float64)


class WaveGrowth1D:
    dims = 1

    def __init__(self, *, grid, winds, ODEsys, ODEvars=None, layers=1, clock=None, ODEsets=None, ODEinit_type="wind_sea",
                 minimal_particle=None, minimal_state=None, currents=None, periodic_boundary=True, boundary_type="same",
                 CBsets=None, architecture=None, nan_eest_rejects=False):
        if not isinstance(grid, OneDGrid):
            raise TypeError("grid must be a OneDGrid")
        if layers != 1:
            raise NotImplementedError("layers > 1: nothing on the reference's stepping path can index it (SURVEY B-12)")
        if ODEsets is None:
            raise ValueError("ODEsets (ODESettings) is required")
        if ODEinit_type != "wind_sea":
            raise NotImplementedError("the B200 1-D path seeds from the wind sea (ODEinit_type = 'wind_sea'); a ParticleDefaults "
                                      "instance would put every particle at defaults.x in the reference (core_1D.jl:215-218)")
        arch = architecture if architecture is not None else B200()
        if isinstance(arch, CPU):
            raise RuntimeError("picles_b200 has no CPU compute path: the Julia reference is the CPU implementation")
        self.architecture = arch
        self.grid, self.layers, self.timestepper = grid, layers, None
        self.clock = clock if clock is not None else Clock(0.0)
        self.ODEvars, self.ODEsystem, self.ODEsettings, self.ODEdefaults = ODEvars, ODEsys, ODEsets, None
        self.winds, self.currents = winds, currents
        self.periodic_boundary = bool(periodic_boundary)
        self.boundary = [] if self.periodic_boundary else [1, grid.Nx]
        self.boundary_defaults = None
        # WaveGrowthModels1D.jl:130-142; an exactly-zero component draws rand_sign() in the reference: +1 here (B-9)
        self.minimal_particle = (FetchRelations.MinimalParticle(2, 0, ODEsets.timestep) if minimal_particle is None
                                 else minimal_particle)
        self.minimal_state = FetchRelations.MinimalState(2, 0, ODEsets.timestep) if minimal_state is None else minimal_state
        self.FailedCollection = []
        self.gridnotes = OneDGridNotes(grid)
        P = make_params(ODEsets, ODEsys, self.minimal_state, defaults=None, periodic_boundary=self.periodic_boundary,
                        nan_eest_rejects=nan_eest_rejects or getattr(architecture, "nan_eest_rejects", False))
        from ..engine1d import B200Engine1D
        self.engine = B200Engine1D(grid.Nx, grid.xmin, grid.dx, self.gridnotes.x, P, device=arch.devices[0])
        self._seeded = False

    def _wind_nodes(self, t):
        return eval_wind_1d(self.winds, self.gridnotes.x, float(t))

    @property
    def State(self):
        """(Nx, 3) like the reference's SharedMatrix: [e, m_x, 0] per node."""
        return np.ascontiguousarray(self.engine.state().T)

    @property
    def ParticleCollection(self):
        return self.engine.particles()

    def counters(self):
        return self.engine.counters()


def fields(model):
    return {"State": model.State}


def reset_boundary(model):
    model.boundary = [] if model.periodic_boundary else [1, model.grid.Nx]


def init_particles_1D(model, defaults=None, verbose=False):
    """init_particles!(model::Abstract1DModel), run.jl:268-302: SeedParticle! for every node with the wind at t = 0."""
    model.engine.seed(model._wind_nodes(0.0))
    model._seeded = True


def time_step_1D(model, Δt, callbacks=None, debug=False):
    """State .= 0 (run.jl:72-80) and time_step!(model::Abstract1DModel, Δt) (TimeSteppers.jl:51-92): advance! every
    particle, merge the charges onto the nodes, remesh!, tick! — one picles1d_step."""
    if not model._seeded:
        raise RuntimeError("init_particles! must run before time_step!")
    t = model.clock.time
    model.engine.step(t, float(Δt), model._wind_nodes(t), model._wind_nodes(t + float(Δt)))
    if debug:
        c = model.counters()
        model.FailedCollection = [c] if c["n_failed"] else []
    model.clock.tick(float(Δt))
