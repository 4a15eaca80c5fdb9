# generate_golden.jl — reference-held golden vectors for the hot path, for a machine that has Julia.
#
# STATUS: written against the reference sources, NOT executed (no Julia in the build image).  It drives the
# UNMODIFIED reference — examples/example_00_minimal.jl:17-67, statement for statement — and, instead of
# plotting, dumps what the parity tests compare:
#     State (Nx, Ny, 3) after the seed and after every model step          (run.jl:55-57, 104-106: the CashStore pushes)
#     the particles' integrator state u[1:5], t, and the `on` flag after every step
# in the layout tests/golden/*.npz use: little-endian Float64, i (x) fastest = Julia's own memory order, so
# every array is written with `write(io, A)` and read back with numpy.fromfile(...).reshape(..., Ny, Nx).
#
#   julia --project=<PiCLES checkout> julia/generate_golden.jl [outdir]        (default: tests/golden/julia/example_00_minimal)
#
# tests/test_reference_golden.py loads the directory when it exists and reports the largest relative error of
# the oracle (and, with -m gpu, of the CUDA path) on State, lne and c̄ against these vectors — the number the
# north-star's 1e-6 tolerance is about.  Until then parity stays "unpinned" (DESIGN.md §2).
#
# The script also records what the survey could only infer (SURVEY.md Appendix B, A.2):
#     - whether `PI.on = …` persists in the StructArray (B-1),
#     - the OrdinaryDiffEq version, algorithm and controller in use (A.2),
#     - per-step counts of integrated / seeded-off particles.

using PiCLES
using PiCLES.Operators.core_2D: ParticleDefaults
using PiCLES.Models.WaveGrowthModels2D: WaveGrowth2D
using PiCLES.Simulations
using PiCLES.Grids.CartesianGrid: TwoDCartesianGridMesh, ProjetionKernel, TwoDCartesianGridStatistics
using PiCLES.ParticleSystems: particle_waves_v5 as PW
using PiCLES.Operators.TimeSteppers: time_step!
using Oceananigans.Units
import Pkg

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "julia", "example_00_minimal")
mkpath(outdir)

# ---- examples/example_00_minimal.jl:17-67, verbatim except plotting ---------------------------------
U10, V10 = 10.0, 10.0
DT = 10minutes
r_g0 = 0.85
u(x, y, t) = U10
v(x, y, t) = V10
winds = (u=u, v=v)
grid = TwoDCartesianGridMesh(100e3, 51, 100e3, 51)
ODEpars, Const_ID, Const_Scg = PW.ODEParameters(r_g=r_g0)
particle_system = PW.particle_equations(u, v, γ=Const_ID.γ, q=Const_ID.q)
WindSeamin = FetchRelations.MinimalWindsea(U10, V10, DT)
default_particle = ParticleDefaults(WindSeamin["lne"], WindSeamin["cg_bar_x"], WindSeamin["cg_bar_y"], 0.0, 0.0)
ODE_settings = PW.ODESettings(
    Parameters=ODEpars,
    log_energy_minimum=WindSeamin["lne"],
    saving_step=DT,
    timestep=DT,
    total_time=T = 6days,
    dt=1e-3,
    dtmin=1e-4,
    force_dtmin=true)
wave_model = WaveGrowth2D(; grid=grid,
    winds=winds,
    ODEsys=particle_system,
    ODEsets=ODE_settings,
    periodic_boundary=false,
    minimal_particle=FetchRelations.MinimalParticle(U10, V10, DT),
    movie=true)
wave_simulation = Simulation(wave_model, Δt=DT, stop_time=2hour)

# ---- run!(wave_simulation, cash_store=true) unrolled (run.jl:36-122) so the particles can be dumped per step ----
initialize_simulation!(wave_simulation)
model = wave_simulation.model
Nx, Ny = size(model.State, 1), size(model.State, 2)

function dump_particles(io_u, io_t, io_on, model)
    U = fill(NaN, Nx, Ny, 5); Tm = fill(NaN, Nx, Ny); On = zeros(UInt8, Nx, Ny)
    for ij in CartesianIndices((Nx, Ny))
        PI = model.ParticleCollection[ij]
        (PI.ODEIntegrator === nothing) && continue
        try
            U[ij, :] = PI.ODEIntegrator.u[1:5]
            Tm[ij] = PI.ODEIntegrator.t
            On[ij] = PI.on ? 0x01 : 0x00
        catch
        end
    end
    write(io_u, U); write(io_t, Tm); write(io_on, On)
end

io_S = open(joinpath(outdir, "state.f64"), "w")
io_u = open(joinpath(outdir, "particles_u.f64"), "w")
io_t = open(joinpath(outdir, "particles_t.f64"), "w")
io_on = open(joinpath(outdir, "particles_on.u8"), "w")
write(io_S, Array(model.State))               # after the seed (run.jl:55-57)
dump_particles(io_u, io_t, io_on, model)
nsteps = 0
clock_times = Float64[model.clock.time]
running = wave_simulation.stop_time >= model.clock.time
while running                                  # run.jl:72-115
    model.State .= 0.0
    time_step!(model, wave_simulation.Δt)
    global nsteps += 1
    write(io_S, Array(model.State))
    dump_particles(io_u, io_t, io_on, model)
    push!(clock_times, model.clock.time)
    global running = wave_simulation.stop_time >= model.clock.time
end
close(io_S); close(io_u); close(io_t); close(io_on)

# ---- B-1: does a mutation of `on` through the StructArray persist? ---------------------------------------
ij = CartesianIndex(10, 10)
before = model.ParticleCollection[ij].on
tmp = model.ParticleCollection[ij]; tmp.on = !before
on_persists = (model.ParticleCollection[ij].on == !before)
on_persists && (tmp2 = model.ParticleCollection[ij]; tmp2.on = before)

integ = model.ParticleCollection[CartesianIndex(10, 10)].ODEIntegrator
deps = Pkg.dependencies()
ver(name) = (v = [string(d.version) for d in values(deps) if d.name == name]; isempty(v) ? "absent" : v[1])
open(joinpath(outdir, "manifest.json"), "w") do io
    println(io, "{")
    println(io, "  \"generator\": \"julia/generate_golden.jl\", \"source\": \"examples/example_00_minimal.jl:17-67\",")
    println(io, "  \"Nx\": $Nx, \"Ny\": $Ny, \"nsteps\": $nsteps, \"DT\": $(Float64(DT)),")
    println(io, "  \"clock_times\": [", join(clock_times, ", "), "],")
    println(io, "  \"layout\": \"state.f64: (nsteps+1) x 3 x Ny x Nx float64 LE (Julia (Nx,Ny,3) memory order); particles_u.f64: (nsteps+1) x 5 x Ny x Nx; particles_t.f64: (nsteps+1) x Ny x Nx; particles_on.u8: (nsteps+1) x Ny x Nx\",")
    println(io, "  \"on_flag_persists_in_structarray\": $(on_persists),")
    println(io, "  \"integrator_alg\": \"$(typeof(integ.alg))\", \"controller\": \"$(typeof(integ.opts.controller))\",")
    println(io, "  \"abstol\": $(integ.opts.abstol), \"reltol\": $(integ.opts.reltol), \"qoldinit\": $(integ.opts.qoldinit),")
    println(io, "  \"julia\": \"$(VERSION)\", \"OrdinaryDiffEq\": \"$(ver("OrdinaryDiffEq"))\", \"DifferentialEquations\": \"$(ver("DifferentialEquations"))\", \"StructArrays\": \"$(ver("StructArrays"))\"")
    println(io, "}")
end
@info "golden vectors written" outdir nsteps on_persists
