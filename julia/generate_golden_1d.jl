# generate_golden_1d.jl — reference-held golden vectors for the ONE-DIMENSIONAL model, for a machine that has Julia.
#
# STATUS: written against the reference sources, NOT executed (no Julia in the build image).  It drives the unmodified
# reference's WaveGrowth1D the way tests/T03_PIC_propagation_1d.jl:100-182 and tests/B01_1D_regtest_wave_growth.jl do
# (u10 = 15 m/s, DT = 10 min, Nx = 51: a B01 case; the grid offset xmin = 1 km of the T03 script), with the solver set
# to Tsit5() so that the run is the one the picles1d_* path and its oracle reproduce (scenario "steady_nonperiodic" of
# tests/scenarios_1d.py), and dumps
#     State (Nx, 3) after the seed and after every model step,
#     the particles' integrator state u[1:3], t, and the `on` flag after every step
# as little-endian Float64 / UInt8 in Julia's own memory order.
#
#   julia --project=<PiCLES checkout> julia/generate_golden_1d.jl [outdir]     (default: tests/golden/julia/oned_steady)
#
# tests/test_reference_golden.py loads the directory when it exists.  What to look at first: whether the merge rule
# (ParticleInCell.jl:228-252) really turns later charges away as the text reads (DESIGN.md §4.6) — State after step 2
# tells — and whether `PI.on` persists in the plain-vector ParticleCollection of the 1-D model.

using PiCLES
using PiCLES.ParticleMesh: OneDGrid, OneDGridNotes
using PiCLES: Simulation, WaveGrowthModels1D, FetchRelations
using PiCLES.Simulations
using PiCLES.ParticleSystems: particle_waves_v5 as PW
using PiCLES.Operators.TimeSteppers: time_step!
using DifferentialEquations
using Oceananigans.Units
import Pkg

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "julia", "oned_steady")
mkpath(outdir)

U10 = 15.0
DT = Float64(10minutes)
Nx = 51
nsteps_wanted = 8
u(x, t) = U10 + x * 0 + t * 0
grid1d = OneDGrid(1e3, 1500e3, Nx)
ODEpars, Const_ID, Const_Scg = PW.ODEParameters(r_g=0.85)
particle_system = PW.particle_equations(u, γ=Const_ID.γ, q=Const_ID.q)
default_ODE_parameters = (r_g=0.85, C_α=Const_Scg.C_alpha, C_e=Const_ID.C_e)
WindSeamin = FetchRelations.MinimalWindsea(10.0, 0.0, DT)
ODE_settings = PW.ODESettings(
    Parameters=default_ODE_parameters,
    log_energy_minimum=WindSeamin["lne"],
    log_energy_maximum=log(17),
    saving_step=DT,
    timestep=DT,
    total_time=6days,
    solver=Tsit5(),
    adaptive=true,
    dt=1e-3,
    dtmin=1e-4,
    force_dtmin=true)
wave_model = WaveGrowthModels1D.WaveGrowth1D(; grid=grid1d, winds=u, ODEsys=particle_system, ODEvars=nothing, layers=1,
    ODEsets=ODE_settings, ODEinit_type="wind_sea", periodic_boundary=false, boundary_type="same")
wave_simulation = Simulation(wave_model, Δt=DT, stop_time=(nsteps_wanted - 1) * DT)
initialize_simulation!(wave_simulation)
model = wave_simulation.model

function dump_particles(io_u, io_t, io_on, model)
    U = fill(NaN, Nx, 3); Tm = fill(NaN, Nx); On = zeros(UInt8, Nx)
    for (i, PI) in enumerate(model.ParticleCollection)
        U[i, :] = PI.ODEIntegrator.u[1:3]
        Tm[i] = PI.ODEIntegrator.t
        On[i] = PI.on ? 0x01 : 0x00
    end
    write(io_u, U); write(io_t, Tm); write(io_on, On)
end

io_S = open(joinpath(outdir, "state.f64"), "w")
io_u = open(joinpath(outdir, "particles_u.f64"), "w")
io_t = open(joinpath(outdir, "particles_t.f64"), "w")
io_on = open(joinpath(outdir, "particles_on.u8"), "w")
write(io_S, Array(model.State))
dump_particles(io_u, io_t, io_on, model)
nsteps = 0
running = wave_simulation.stop_time >= model.clock.time
while running                                  # run.jl:72-115
    model.State[:, :, :] .= 0.0
    time_step!(model, wave_simulation.Δt)
    global nsteps += 1
    write(io_S, Array(model.State))
    dump_particles(io_u, io_t, io_on, model)
    global running = wave_simulation.stop_time >= model.clock.time
end
close(io_S); close(io_u); close(io_t); close(io_on)

before = model.ParticleCollection[10].on
model.ParticleCollection[10].on = !before
on_persists = (model.ParticleCollection[10].on == !before)
model.ParticleCollection[10].on = before

integ = model.ParticleCollection[10].ODEIntegrator
deps = Pkg.dependencies()
ver(name) = (v = [string(d.version) for d in values(deps) if d.name == name]; isempty(v) ? "absent" : v[1])
gn = OneDGridNotes(grid1d)
open(joinpath(outdir, "manifest.json"), "w") do io
    println(io, "{")
    println(io, "  \"generator\": \"julia/generate_golden_1d.jl\", \"Nx\": $Nx, \"nsteps\": $nsteps, \"DT\": $DT, \"U10\": $U10,")
    println(io, "  \"xmin\": $(grid1d.xmin), \"dx\": $(grid1d.dx), \"x_first\": $(gn.x[1]), \"x_last\": $(gn.x[end]),")
    println(io, "  \"layout\": \"state.f64: (nsteps+1) x 3 x Nx float64 LE (Julia (Nx,3) memory order); particles_u.f64: (nsteps+1) x 3 x Nx; particles_t.f64: (nsteps+1) x Nx; particles_on.u8: (nsteps+1) x Nx\",")
    println(io, "  \"on_flag_persists\": $(on_persists), \"integrator_alg\": \"$(typeof(integ.alg))\",")
    println(io, "  \"julia\": \"$(VERSION)\", \"OrdinaryDiffEq\": \"$(ver("OrdinaryDiffEq"))\", \"DifferentialEquations\": \"$(ver("DifferentialEquations"))\"")
    println(io, "}")
end
@info "1-D golden vectors written" outdir nsteps on_persists
