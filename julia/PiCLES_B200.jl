# PiCLES_B200.jl — the binding a PiCLES maintainer adds to run `init_particles!` / `time_step!`
# on a B200 through libpicles_b200.so (C ABI: include/picles_b200.h).
#
# STATUS: written against the C header, NOT executed — Julia is not installed in the build
# image.  tests/test_abi.py checks the struct layouts used below (field order, offsets, sizes)
# against the compiled header; picles_b200/{engine,params}.py is the executed 1:1 mirror.
#
# Where it plugs into the reference (paths relative to the PiCLES checkout):
#   src/Architectures.jl            add `AbstractArchitecture`, `CPU`, `B200` (below) and export them
#   src/Models/WaveGrowthModels2D.jl:194   add kwarg `architecture = CPU()`, store it in the struct
#   src/Simulations/run.jl:75-82    `State .= 0; time_step!(model, Δt)` dispatches on model.architecture
#   src/Simulations/run.jl:136,167  `init_particles!(model; defaults)` dispatches likewise
# Host code stays Julia: wind closures u(x,y,t), v(x,y,t) are broadcast over the mesh every step
# and handed to the library as plain `Matrix{Float64}` (column-major (Nx,Ny) == the library's
# row pitch Nx, i fastest — no transposition).

module PiCLES_B200

export B200, B200Context, b200_init_particles!, b200_time_step!, b200_fetch_state!, b200_counters

const LIB = get(ENV, "PICLES_B200_LIB", joinpath(@__DIR__, "..", "picles_b200", "libpicles_b200.so"))

# ---- src/Architectures.jl additions ---------------------------------------------------------
abstract type AbstractArchitecture end
struct CPU <: AbstractArchitecture end
"""
    B200(; device=0, terms=(propagation=true, input=true, dissipation=true, peak_shift=true, direction=true),
          γ, q, on_persist=false)

`particle_equations(u, v; γ, q, propagation, input, …)` returns a closure, so the term switches
and (γ, q) cannot be read back from `model.ODEsystem`; they are repeated here.
"""
Base.@kwdef struct B200 <: AbstractArchitecture
    device::Int = 0
    γ::Float64
    q::Float64 = -1 / 4
    propagation::Bool = true
    input::Bool = true
    dissipation::Bool = true
    peak_shift::Bool = true
    direction::Bool = true
    on_persist::Bool = false      # SURVEY B-1: false = `on` frozen at seed, as the reference runs
    nan_eest_rejects::Bool = false  # a NaN error estimate: false = DtNaN (exact powers), true = rejected by 1/qmin (fastpow / fastpower)
    # multi-GPU: this process owns rows j0+1 : j0+ny_local of the global grid
    rank::Int = 0
    nranks::Int = 1
    halo::Int = 2
    # wind ingestion (include/picles_b200.h: picles_set_wind_midlevels / picles_set_wind_mesh)
    wind_levels::Int = 2          # 2..5 levels staged per step; the closures are called at wind_levels
                                  # equally spaced times of [t, t+Δt] and interpolated in time on the device
    wind_mesh = nothing           # (x=, y=, t=, u=U[ix,iy,it], v=V[ix,iy,it]): gridded winds kept on the device
                                  # (the data behind LinearInterpolation((x,y,t), U, extrapolation_bc=Periodic()),
                                  # tests/T03_PIC_tripolar_realistic.jl:61-73); model.winds is then not called
end

# ---- C structs (include/picles_b200.h) -------------------------------------------------------
struct PiclesParams                # picles_params_t, 264 bytes
    r_g::Cdouble; C_alpha::Cdouble; C_varphi::Cdouble; C_e::Cdouble; g::Cdouble
    p::Cdouble; q::Cdouble; n::Cdouble; e_T::Cdouble
    propagation::Int32; input::Int32; dissipation::Int32; peak_shift::Int32; direction::Int32
    solver::Int32
    abstol::Cdouble; reltol::Cdouble; dt::Cdouble; dtmin::Cdouble; dtmax::Cdouble
    force_dtmin::Int32; adaptive::Int32
    maxiters::Int64
    log_energy_minimum::Cdouble; log_energy_maximum::Cdouble; wind_min_squared::Cdouble; seed_timescale::Cdouble
    minimal_state::NTuple{2,Cdouble}
    has_defaults::Int32
    defaults::NTuple{5,Cdouble}
    periodic_boundary::Int32; on_persist::Int32; nan_eest_rejects::Int32
end

struct PiclesCounters              # picles_counters_t
    n_active::Int64; n_integrated::Int64; n_substeps::Int64; n_rejects::Int64; n_rhs::Int64
    n_reseed_advance::Int64; n_fixups::Int64; n_failed::Int64; n_deposited::Int64
    n_remesh_A::Int64; n_remesh_B::Int64; n_remesh_C::Int64; n_remesh_D::Int64
    reach::Int32; max_attempts::Int32
    ms_advance::Cdouble; ms_project::Cdouble; ms_remesh::Cdouble
    n_stiff_switches::Int64; n_stiff_attempts::Int64
end

mutable struct B200Context
    handle::Ptr{Cvoid}
    Nx::Int; Ny::Int; j0::Int; ny::Int
    lo::Int; hi::Int                 # neighbour ranks (-1: none)
    u_t1::Matrix{Float64}; v_t1::Matrix{Float64}   # staging of the wind at t+Δt
    first_step::Bool
    n_mid::Int                       # intermediate wind levels per step (arch.wind_levels - 2)
    mesh_winds::Bool                 # winds sampled on the device from the resident wind mesh
end

function check(h::Ptr{Cvoid}, rc::Integer)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:picles_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    error("libpicles_b200: status $rc: $msg")
end

# ---- flattening, exactly as the reference derives its constants --------------------------------
# magic_fractions / e_T_func: src/ParticleSystems/particle_waves_v5.jl:87-92,271
magic_fractions(q) = (p=(-1 - 10q) / 2, q=q, n=2q / ((-1 - 10q) / 2 + 4q))
e_T_func(γ, p, q, n; C_e=2.16e-4, c_β=4e-2, c_D=2e-3, c_e=1.3e-6, c_α=11.8) =
    sqrt(c_e * c_α^(-p / q) / (γ * c_β * c_D)^(1 / n))

boundary_code(N) = occursin("TripolarNorth", string(typeof(N))) ? 2 : occursin("N_Periodic", string(typeof(N))) ? 1 : 0
# DP5 -> 1; AutoTsit5(Rosenbrock23()) (a CompositeAlgorithm, the ODESettings default) -> 2: Tsit5 with
# OrdinaryDiffEq's AutoSwitch monitor and the Rosenbrock23 branch; Tsit5 -> 0
solver_code(s) = (n = string(typeof(s)); occursin("Composite", n) ? 2 : occursin("DP5", n) ? 1 : 0)

function flatten_params(model, arch::B200)
    S = model.ODEsettings
    par = S.Parameters
    mf = magic_fractions(arch.q)
    d = model.ODEdefaults
    PiclesParams(par.r_g, par.C_α, get(par, :C_φ, 0.0), par.C_e, get(par, :g, 9.81),   # the 1-D parameter set has no C_φ
        mf.p, mf.q, mf.n, e_T_func(arch.γ, mf.p, mf.q, mf.n),
        arch.propagation, arch.input, arch.dissipation, arch.peak_shift, arch.direction,
        solver_code(S.solver),
        S.abstol, S.reltol, S.dt, S.dtmin, S.total_time,        # OrdinaryDiffEq dtmax default = tspan length
        S.force_dtmin, S.adaptive, S.maxiters,
        S.log_energy_minimum, S.log_energy_maximum, S.wind_min_squared, S.timestep,
        (model.minimal_state[1], model.minimal_state[2]),
        d === nothing ? 0 : 1,
        d === nothing ? (0.0, 0.0, 0.0, 0.0, 0.0) : (d.lne, d.c̄_x, d.c̄_y, d.x, d.y),
        model.periodic_boundary, arch.on_persist, arch.nan_eest_rejects)
end

# wind closures -> mesh arrays of this strip (north-star: evaluated on the host every step)
stage_wind(f, grid, rows, t) = Float64[f(grid.data.x[i, j], grid.data.y[i, j], t) for i in axes(grid.data.x, 1), j in rows]

# ---- init_particles!(model) for B200: src/Simulations/run.jl:199-247 -----------------------------
function b200_init_particles!(model, arch::B200; nccl_id::Union{Nothing,Vector{UInt8}}=nothing)
    grid = model.grid
    Nx, Ny = grid.stats.Nx.N, grid.stats.Ny.N
    bounds = [(Ny * r ÷ arch.nranks, Ny * (r + 1) ÷ arch.nranks) for r in 0:arch.nranks-1]
    j0, j1 = bounds[arch.rank+1]
    rows = j0+1:j1
    href = Ref{Ptr{Cvoid}}(C_NULL)
    check(C_NULL, ccall((:picles_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Cint), href, arch.device))
    h = href[]
    mask = UInt8.(grid.data.mask[:, rows])
    halo = arch.nranks > 1 ? arch.halo : 0
    if hasproperty(grid.data, :angle_dx)            # MOM6GridMesh: per-node kernel + great-circle term
        # library derives M = [cosα/dx sinα/dy; -sinα/dx cosα/dy] and sign(φ)·min(sign(φ)·tand(φ),60)/R on the device
        check(h, ccall((:picles_set_grid_metric, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Ptr{UInt8}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble),
            h, Nx, Ny, boundary_code(grid.stats.Nx), boundary_code(grid.stats.Ny), j0, j1 - j0, halo, mask,
            Float64.(grid.data.dx[:, rows]), Float64.(grid.data.dy[:, rows]), Float64.(grid.data.angle_dx[:, rows]),
            Float64.(grid.data.y[:, rows]), 6.3710e6))
    else                                             # TwoDCartesianGridMesh: uniform kernel
        # the grid's own ProjetionKernel(stats) (CartesianGrid.jl:115-129): diag(1/dx, 1/dy), or — for a rotated grid,
        # angle_dx != 0 — [cosα/dx sinα/dy; sinα/dx cosα/dy] (no minus sign: SURVEY B-8).  Row-major M11, M12, M21, M22
        Mk = grid.ProjetionKernel(grid.stats)
        Mc = Float64[Mk[1, 1], Mk[1, 2], Mk[2, 1], Mk[2, 2]]
        check(h, ccall((:picles_set_grid, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, Cint, Cint, Ptr{UInt8}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
            h, Nx, Ny, boundary_code(grid.stats.Nx), boundary_code(grid.stats.Ny), j0, j1 - j0, halo, mask, C_NULL, Mc, C_NULL))
    end
    P = flatten_params(model, arch)
    check(h, ccall((:picles_set_params, LIB), Cint, (Ptr{Cvoid}, Ref{PiclesParams}), h, P))
    2 <= arch.wind_levels <= 5 || error("wind_levels must be between 2 and 5")
    mesh_winds = arch.wind_mesh !== nothing
    if mesh_winds
        w = arch.wind_mesh           # U[ix, iy, it] column-major == nt slices of ny*nx, x fastest
        check(h, ccall((:picles_set_wind_mesh, LIB), Cint,
            (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
            h, length(w.x), length(w.y), length(w.t), Float64.(collect(w.x)), Float64.(collect(w.y)), Float64.(collect(w.t)),
            Float64.(w.u), Float64.(w.v), Float64.(grid.data.x[:, rows]), Float64.(grid.data.y[:, rows])))
        check(h, ccall((:picles_seed_wind_mesh, LIB), Cint, (Ptr{Cvoid}, Cdouble), h, 0.0))
        u0 = v0 = zeros(0, 0)
    else
        u0 = stage_wind(model.winds.u, grid, rows, 0.0)
        v0 = stage_wind(model.winds.v, grid, rows, 0.0)
        check(h, ccall((:picles_seed, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), h, u0, v0))
    end
    lo = hi = -1
    if arch.nranks > 1
        # rank 0 creates the id with picles_comm_unique_id and broadcasts it (MPI.Bcast! / Distributed)
        nccl_id === nothing && error("multi-GPU: pass the 128-byte NCCL id created by rank 0")
        check(h, ccall((:picles_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint, Cstring), h, nccl_id, arch.rank, arch.nranks, C_NULL))
        per = boundary_code(grid.stats.Ny) == 1
        lo = arch.rank > 0 ? arch.rank - 1 : (per ? arch.nranks - 1 : -1)
        hi = arch.rank < arch.nranks - 1 ? arch.rank + 1 : (per ? 0 : -1)
    end
    ctx = B200Context(h, Nx, Ny, j0, j1 - j0, lo, hi, u0, v0, true, arch.wind_levels - 2, mesh_winds)
    b200_fetch_state!(model, ctx)                   # State after seeding (init_z0_to_State!)
    return ctx
end

b200_nccl_unique_id() = (id = zeros(UInt8, 128); check(C_NULL, ccall((:picles_comm_unique_id, LIB), Cint, (Ptr{UInt8}, Cstring), id, C_NULL)); id)

# ---- State .= 0 ; time_step!(model, Δt) for B200: run.jl:75-82, TimeSteppers.jl:109-166 ----------
function b200_time_step!(model, ctx::B200Context, Δt::Float64; fetch_state::Bool=true)
    t = model.clock.time
    rows = ctx.j0+1:ctx.j0+ctx.ny
    if ctx.mesh_winds
        # every wind level of the step is sampled on the device from the resident wind mesh: no upload
        check(ctx.handle, ccall((:picles_step_wind_mesh, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cint, Cint, Cint),
            ctx.handle, t, Δt, ctx.n_mid, ctx.lo, ctx.hi))
        fetch_state && b200_fetch_state!(model, ctx)
        return nothing
    end
    if ctx.n_mid > 0
        # intermediate levels at t + Δt*k/(n_mid+1), k = 1..n_mid (the expression the library documents)
        um = Array{Float64,3}(undef, ctx.Nx, ctx.ny, ctx.n_mid)
        vm = similar(um)
        for k in 1:ctx.n_mid
            tk = t + Δt * Float64(k) / Float64(ctx.n_mid + 1)
            um[:, :, k] = stage_wind(model.winds.u, model.grid, rows, tk)
            vm[:, :, k] = stage_wind(model.winds.v, model.grid, rows, tk)
        end
        check(ctx.handle, ccall((:picles_set_wind_midlevels, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Ptr{Cdouble}),
            ctx.handle, ctx.n_mid, um, vm))
    end
    # wind at t+Δt on the mesh; the level uploaded last step becomes this step's t level inside the library
    ctx.u_t1 = stage_wind(model.winds.u, model.grid, rows, t + Δt)
    ctx.v_t1 = stage_wind(model.winds.v, model.grid, rows, t + Δt)
    if ctx.lo < 0 && ctx.hi < 0 && ctx.ny == ctx.Ny
        check(ctx.handle, ccall((:picles_step, LIB), Cint,
            (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
            ctx.handle, t, Δt, C_NULL, C_NULL, ctx.u_t1, ctx.v_t1))
    else
        check(ctx.handle, ccall((:picles_step_strip, LIB), Cint,
            (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Cint),
            ctx.handle, t, Δt, C_NULL, C_NULL, ctx.u_t1, ctx.v_t1, ctx.lo, ctx.hi))
    end
    fetch_state && b200_fetch_state!(model, ctx)
    # tick!(model.clock, Δt) stays with the caller exactly as in TimeSteppers.jl:163
    return nothing
end

# model.State[:, rows, 1:3] <- device planes e, m_x, m_y (only needed when a store or a plot reads State)
function b200_fetch_state!(model, ctx::B200Context)
    S = Array{Float64,3}(undef, ctx.Nx, ctx.ny, 3)
    check(ctx.handle, ccall((:picles_get_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.handle, S))
    model.State[:, ctx.j0+1:ctx.j0+ctx.ny, :] .= S
    return nothing
end

function b200_counters(ctx::B200Context)
    c = Ref{PiclesCounters}()
    check(ctx.handle, ccall((:picles_get_counters, LIB), Cint, (Ptr{Cvoid}, Ref{PiclesCounters}), ctx.handle, c))
    return c[]
end

b200_destroy!(ctx::B200Context) = (ccall((:picles_destroy, LIB), Cint, (Ptr{Cvoid},), ctx.handle); ctx.handle = C_NULL; nothing)

# ---- the one-dimensional model (WaveGrowth1D): picles1d_* ---------------------------------------------------
# init_particles!(model::Abstract1DModel) (run.jl:268-302) and State .= 0 ; time_step!(model::Abstract1DModel, Δt)
# (TimeSteppers.jl:51-92).  The parameter struct is the 2-D one (flatten_params).
mutable struct B200Context1D
    handle::Ptr{Cvoid}
    Nx::Int
    x::Vector{Float64}      # OneDGridNotes.x: where the wind closure is evaluated
end

function check1d(h::Ptr{Cvoid}, rc::Integer)
    rc == 0 && return nothing
    error("picles_b200 (1-D) error $rc: " * unsafe_string(ccall((:picles1d_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))
end

function b200_init_particles_1d!(model, arch::B200)
    model.ODEdefaults === nothing || error("the B200 1-D path seeds from the wind sea (ODEinit_type = \"wind_sea\")")
    href = Ref{Ptr{Cvoid}}(C_NULL)
    check1d(C_NULL, ccall((:picles1d_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Cint), href, arch.device))
    h = href[]
    g = model.grid
    x = collect(LinRange(0, g.dimx, g.Nx))                      # OneDGridNotes(grid).x, ParticleMesh.jl:131
    check1d(h, ccall((:picles1d_set_grid, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Ptr{Cdouble}), h, g.Nx, g.xmin, g.dx, x))
    P = flatten_params(model, arch)                             # same struct; periodic_boundary = the model kwarg
    check1d(h, ccall((:picles1d_set_params, LIB), Cint, (Ptr{Cvoid}, Ref{PiclesParams}), h, P))
    u0 = Float64[model.winds(xi, 0.0) for xi in x]
    check1d(h, ccall((:picles1d_seed, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h, u0))
    ctx = B200Context1D(h, g.Nx, x)
    b200_fetch_state_1d!(model, ctx)
    return ctx
end

function b200_time_step_1d!(model, ctx::B200Context1D, Δt::Float64)
    t = model.clock.time
    u_t = Float64[model.winds(xi, t) for xi in ctx.x]
    u_t1 = Float64[model.winds(xi, t + Δt) for xi in ctx.x]
    check1d(ctx.handle, ccall((:picles1d_step, LIB), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
        ctx.handle, t, Δt, u_t, u_t1))
    b200_fetch_state_1d!(model, ctx)
    return nothing                                              # tick!(model.clock, Δt) stays with the caller
end

function b200_fetch_state_1d!(model, ctx::B200Context1D)
    S = Matrix{Float64}(undef, ctx.Nx, 3)                       # (Nx, 3) column-major, as model.State
    check1d(ctx.handle, ccall((:picles1d_get_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), ctx.handle, S))
    model.State[:, :] .= S
    return nothing
end

function b200_counters(ctx::B200Context1D)
    c = Ref{PiclesCounters}()
    check1d(ctx.handle, ccall((:picles1d_get_counters, LIB), Cint, (Ptr{Cvoid}, Ref{PiclesCounters}), ctx.handle, c))
    return c[]
end

b200_destroy!(ctx::B200Context1D) = (ccall((:picles1d_destroy, LIB), Cint, (Ptr{Cvoid},), ctx.handle); ctx.handle = C_NULL; nothing)

end # module
