# probe_integrator.jl — one particle, one integrator: what OrdinaryDiffEq really does on this path.
#
# STATUS: written against the reference sources, NOT executed (no Julia in the build image).
#
# The adaptive Runge-Kutta arithmetic of the hot path lives in OrdinaryDiffEq.jl, which the reference does not pin
# (Project.toml:16, no [compat], Manifest git-ignored).  The oracle restates the published algorithm; this script
# asks the installed version the questions that restatement had to decide (DESIGN.md §2, quirk table):
#   * substep counts, proposed dt and controller memory (qold) after each `step!(integ, DT, true)`;
#   * what a trial step that overflows does (`blowup` cases: DP5 at the reference's tolerances near |wind| = 14 m/s
#     takes a trial step whose stages reach Inf).  With exact powers in the PI controller, EEst = NaN is rejected,
#     `dt /= min(1/qmin, NaN)` = NaN and check_error! ends the integrator with DtNaN, `integ.u` still holding the NaN
#     trial (so advance!'s NaN fix-up, mapping_2D.jl:196-211, reseeds a dead particle).  With the Float32 `fastpow`
#     of older versions NaN^β1 is a large finite number, the step is rejected by a factor 5 and the run goes on.
#     The oracle takes the first reading and keeps the last accepted state (status UNSTABLE).
#
#   julia --project=<PiCLES checkout> julia/probe_integrator.jl [outfile]     (default: tests/golden/julia/integrator_probe.json)
#
# tests/test_independent_integrator.py::test_oracle_against_the_julia_integrator_probe loads the file when present.

using PiCLES
using PiCLES.Operators.core_2D: ParticleDefaults, InitParticleInstance
using PiCLES.ParticleSystems: particle_waves_v5 as PW
using DifferentialEquations
import Pkg

outfile = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "julia", "integrator_probe.json")
mkpath(dirname(outfile))

# the cases of tests/test_independent_integrator.py::PROBE_CASES, same order
cases = [
    (name="tsit5_u10_v10", solver=Tsit5(), U=10.0, V=10.0, dt=1e-3, dtmin=1e-4, nDT=3, timescale=600.0),
    (name="tsit5_u10_v4", solver=Tsit5(), U=10.0, V=4.0, dt=1e-3, dtmin=1e-4, nDT=3, timescale=600.0),
    (name="dp5_u10_v4", solver=DP5(), U=10.0, V=4.0, dt=1e-3, dtmin=1e-4, nDT=3, timescale=600.0),
    (name="dp5_bench06", solver=DP5(), U=10.0, V=10.0, dt=10.0, dtmin=1.0, nDT=3, timescale=1800.0),
    (name="dp5_blowup_u0_v14", solver=DP5(), U=0.0, V=14.0, dt=1e-3, dtmin=1e-4, nDT=2, timescale=600.0),
    (name="dp5_blowup_um7_v12", solver=DP5(), U=-7.0, V=12.0, dt=1e-3, dtmin=1e-4, nDT=2, timescale=600.0),
    # the ODESettings default solver on the parameter set of tests/T04_2D_reg_test.jl:56-58 (C_φ = c_β = 4e-2): stiff,
    # the AutoSwitch hands the particle to Rosenbrock23 within its first model step (C_phi overrides ODEParameters')
    (name="auto_stiff_um10_vm10", solver=AutoTsit5(Rosenbrock23()), U=-10.0, V=-10.0, dt=1e-3, dtmin=1e-4, nDT=3, timescale=600.0, C_phi=4e-2),
    (name="auto_nonstiff_u10_v10", solver=AutoTsit5(Rosenbrock23()), U=10.0, V=10.0, dt=1e-3, dtmin=1e-4, nDT=3, timescale=600.0),
]

DT = 600.0
jnum(x) = isfinite(x) ? string(x) : (isnan(x) ? "\"nan\"" : (x > 0 ? "\"inf\"" : "\"-inf\""))
jvec(v) = "[" * join(jnum.(v), ", ") * "]"

records = String[]
for c in cases
    u(x, y, t) = c.U
    v(x, y, t) = c.V
    ODEpars, Const_ID, Const_Scg = PW.ODEParameters(r_g=0.85)
    if haskey(c, :C_phi)
        ODEpars = merge(ODEpars, (C_φ=c.C_phi,))
    end
    particle_system = PW.particle_equations(u, v, γ=Const_ID.γ, q=Const_ID.q)
    sets = PW.ODESettings(Parameters=ODEpars, log_energy_minimum=FetchRelations.MinimalWindsea(10.0, 10.0, DT)["lne"],
        saving_step=DT, timestep=DT, total_time=6 * 86400.0, solver=c.solver, dt=c.dt, dtmin=c.dtmin, force_dtmin=true)
    # SeedParticle (core_2D.jl:434-488): initial state from the fetch relations, projection kernel of a 2 km Cartesian cell
    ui = FetchRelations.get_initial_windsea(c.U, c.V, c.timescale; particle_state=true)
    z0 = ParticleDefaults(ui[1], ui[2], ui[3], 0.0, 0.0)
    sets.Parameters = (; sets.Parameters..., M=[1/2000.0 0.0; 0.0 1/2000.0], PC=0.0)
    PI = InitParticleInstance(particle_system, z0, sets, (1, 1), (0.0, 0.0), false, true)
    integ = PI.ODEIntegrator
    steps = String[]
    for k in 1:c.nDT
        step!(integ, DT, true)
        st = hasproperty(integ, :stats) ? integ.stats : integ.destats   # renamed in SciMLBase 1.7x
        current = hasproperty(integ.cache, :current) ? integ.cache.current : 1      # CompositeCache: 1 = Tsit5, 2 = Rosenbrock23
        push!(steps, "{\"u\": $(jvec(integ.u[1:5])), \"t\": $(jnum(integ.t)), \"dt\": $(jnum(integ.dt)), \"qold\": $(jnum(integ.qold)), \"current_alg\": $current, " *
                     "\"naccept\": $(st.naccept), \"nreject\": $(st.nreject), \"nf\": $(st.nf), \"iter\": $(integ.iter), " *
                     "\"retcode\": \"$(integ.sol.retcode)\"}")
    end
    # what reset_PI_u! (mapping_2D.jl:91-96) does to an integrator that has failed: does the next step! integrate again
    # (old DiffEqBase: check_error re-evaluated every call) or return at once (SciMLBase: a bad retcode is sticky)?
    t_before = integ.t
    set_u!(integ, [ui[1], ui[2], ui[3], 0.0, 0.0]); u_modified!(integ, true); auto_dt_reset!(integ)
    dt_after_reset = integ.dt
    step!(integ, DT, true)
    after = "{\"t_before\": $(jnum(t_before)), \"dt_after_reset\": $(jnum(dt_after_reset)), \"t_after\": $(jnum(integ.t)), " *
            "\"retcode\": \"$(integ.sol.retcode)\", \"u\": $(jvec(integ.u[1:5]))}"
    push!(records, "  {\"name\": \"$(c.name)\", \"after_reset\": $after, \"solver\": \"$(typeof(c.solver).name.name)\", \"wind\": [$(c.U), $(c.V)], " *
                   "\"dt\": $(c.dt), \"dtmin\": $(c.dtmin), \"timescale\": $(c.timescale), \"z0\": $(jvec([ui[1], ui[2], ui[3], 0.0, 0.0])),\n" *
                   "   \"steps\": [" * join(steps, ",\n             ") * "]}")
end

deps = Pkg.dependencies()
ver(name) = (v = [string(d.version) for d in values(deps) if d.name == name]; isempty(v) ? "absent" : v[1])
open(outfile, "w") do io
    println(io, "{\"generator\": \"julia/probe_integrator.jl\", \"DT\": $DT, \"counters\": \"cumulative over the integrator's life (integ.stats)\",")
    println(io, " \"julia\": \"$(VERSION)\", \"OrdinaryDiffEq\": \"$(ver("OrdinaryDiffEq"))\", \"DifferentialEquations\": \"$(ver("DifferentialEquations"))\",")
    println(io, " \"cases\": [")
    println(io, join(records, ",\n"))
    println(io, " ]}")
end
@info "integrator probe written" outfile
