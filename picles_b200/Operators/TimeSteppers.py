"""`time_step!` / `movie_time_step!` of `src/Operators/TimeSteppers.jl:109-166,212-247` on the
B200 architecture: each is one `picles_step` through the C ABI plus the clock tick."""
from __future__ import annotations

import numpy as np


def _stage_and_step(model, Δt, accumulate):
    eng = model.engine
    t = model.clock.time
    n_mid = int(getattr(model.architecture, "wind_levels", 2)) - 2
    eng.set_accumulate(accumulate)
    if model._gridded_winds is not None:
        # wind ingestion: every level is sampled on the device from the resident wind mesh
        eng.step_wind_mesh(t, Δt, n_mid)
        model._wind_level_time = t + Δt
        return
    if n_mid > 0:
        # intermediate levels at t + Δt*k/(n_mid+1) (the same expression the library uses)
        lv = [model._wind_planes(t + Δt * float(k) / float(n_mid + 1)) for k in range(1, n_mid + 1)]
        eng.set_wind_midlevels([a for a, _ in lv], [b for _, b in lv])
    # wind at the pre-step clock time t (remesh uses it, TimeSteppers.jl:146) and at t+Δt
    if model._wind_level_time == t:
        u_t = v_t = None  # the device's t+Δt level of the previous step is this step's t level
    else:
        u_t, v_t = model._wind_planes(t)
    u_t1, v_t1 = model._wind_planes(t + Δt)
    eng.step(t, Δt, u_t, v_t, u_t1, v_t1)
    model._wind_level_time = t + Δt


def time_step(model, Δt, callbacks=None, debug=False, zero_state_first=False):
    """time_step!(model, Δt): advance! all ocean_points, project, remesh!, tick!.
    Like the reference it adds the deposits to whatever State holds; `run` zeroes State
    first (run.jl:75-79), which is `zero_state_first=True` here."""
    if not model._seeded:
        raise RuntimeError("init_particles! / initialize_simulation! must run before time_step!")
    _stage_and_step(model, float(Δt), accumulate=not zero_state_first)
    if debug:
        c = model.counters()
        model.FailedCollection = [c] if c["n_failed"] else []
    model.clock.tick(float(Δt))


def movie_time_step(model, Δt, callbacks=None, debug=False):
    """movie_time_step!: advance; MovieState = copy(State); remesh; State .= 0; tick!."""
    if not model._seeded:
        raise RuntimeError("init_particles! / initialize_simulation! must run before movie_time_step!")
    _stage_and_step(model, float(Δt), accumulate=True)
    model.MovieState = np.array(model.State, copy=True)
    model.engine.zero_state()
    if debug:
        c = model.counters()
        model.FailedCollection = [c] if c["n_failed"] else []
    model.clock.tick(float(Δt))
