#!/usr/bin/env python
"""bench.py — particle-steps/s of the PiCLES per-timestep particle-in-cell loop on B200.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on all host cores

Workload (BASELINE.json configs[1]): T04_2D_reg_test_large_grid / bench06 homogeneous box
scaled to 4096x4096 Cartesian, dx=dy=2 km, constant wind u=v=10 m/s, model step 10 min,
example_00_minimal ODE settings (Tsit5 controller, dt=1e-3, dtmin=1e-4, force_dtmin).
A "step" is one model step: State .= 0; advance! (adaptive RK over DT for every particle);
ParticleToNode! projection; remesh!.  With N GPUs every rank owns a 4096x4096 y-strip of a
4096 x (4096*N) box (weak scaling) and exchanges a halo of particle records per step.

One JSON line is printed by rank 0.  See DESIGN.md §measurement for every field.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the 4096^2 bench step, from the
# committed ncu --set full captures (profiles/r01_ncu_final.txt); never measured under the timer
NCU_TRAFFIC = {"k_advance": 3.61e9, "k_project_remesh": 1.80e9, "k_wind_sample": 0.49e9}
NCU_FP64_PIPE_PCT = 60.1  # sm__pipe_fp64_cycles_active of k_advance (profiles/r01_ncu_final.txt)

METRIC = "particle-steps/s"
UNIT = "particle-steps/s"

# Algorithmic FP64 work (DESIGN.md §advance kernel): source-level operation counts of
# physics.h with fma = 2, add/mul/compare-select = 1, div = sqrt = 8, exp = 30,
# tanh = sech = log = 35, pow = 70 flop.
F_RHS = 329          # one right-hand side (rhs3 + prop + wind interpolation)
F_ATTEMPT = 511      # one RK attempt minus its 6 RHS: stage sums, error norm, PI controller
F_INITDT = 319       # Hairer initial step minus its RHS
F_DEPOSIT = 105      # charge + weights of the deposit record
# Algorithmic HBM bytes per node of the fused gather + remesh kernel: read one
# deposit record (5 f64 + packed cell) and the particle flags; write 3 f64 of State, the
# particle's u[5] and its flags (remesh branch A, the steady state of an all-ocean box)
B_PROJECT_REMESH = 44 + 1 + 24 + 40 + 1


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------
def workload(nx, ny_per_gpu, n_gpus, rank):
    """Global homogeneous box Nx x (ny_per_gpu*n_gpus), both axes non-periodic, all ocean."""
    Ny = ny_per_gpu * n_gpus
    j0 = rank * ny_per_gpu
    mask = np.ones((ny_per_gpu, nx), np.uint8)
    mask[:, 0] = 3
    mask[:, -1] = 3
    if j0 == 0:
        mask[0, :] = 3
    if j0 + ny_per_gpu == Ny:
        mask[-1, :] = 3
    return dict(Nx=nx, Ny=Ny, j0=j0, ny=ny_per_gpu, mask=mask, M_const=np.array([1 / 2000.0, 0.0, 0.0, 1 / 2000.0]))


SOLVER = "Tsit5"


def params(solver=None):
    """example_00_minimal.jl:17-67 settings (T04_2D_reg_test uses the same), flattened by the host
    mirror of the reference API — no test or oracle module is involved on the b200 arm."""
    from picles_b200 import FetchRelations as FR
    from picles_b200.ParticleSystems import particle_waves_v5 as PW
    from picles_b200.params import make_params
    DT = 600.0
    pars, cid, _ = PW.ODEParameters(r_g=0.85)
    ps = PW.particle_equations(None, None, γ=cid.γ, q=cid.q)
    sets = PW.ODESettings(Parameters=pars, log_energy_minimum=FR.MinimalWindsea(10, 10, DT)["lne"], saving_step=DT,
                          timestep=DT, total_time=6 * 86400.0, dt=1e-3, dtmin=1e-4, force_dtmin=True,
                          solver=solver or SOLVER)
    return make_params(sets, ps, FR.MinimalState(2, 2, DT), defaults=None, periodic_boundary=False, on_persist=False)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (driver-measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


# ---------------------------------------------------------------------------------------
def cpu_arm(steps, warmup, sample_n=512, threads=None):
    """The reference arm / cpu_baseline: the oracle port (OpenMP over particles in the ODE
    phase, serial canonical-order deposit) on the host cores, on a bounded sample of the
    same workload: a sample_n x sample_n homogeneous box, the same step indices."""
    import oracle
    from common import cartesian_grid, make_oracle
    threads = threads or os.cpu_count() or 1
    g = cartesian_grid(sample_n, sample_n)
    o = make_oracle(g, params(), variant="omp", threads=threads)
    o.seed(10.0, 10.0)
    t = 0.0
    for _ in range(warmup):
        o.step(t, 600.0, 10.0, 10.0, 10.0, 10.0)
        t += 600.0
    n_active = (sample_n - 2) ** 2
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(t, 600.0, 10.0, 10.0, 10.0, 10.0)
        t += 600.0
    dt = time.perf_counter() - t0
    c = o.counters()
    return dict(value=n_active * steps / dt, unit=UNIT, cores=threads, kind="port",
                sample=f"{sample_n}x{sample_n} homogeneous box (same physics/settings), steps {warmup + 1}..{warmup + steps}, "
                       f"oracle/picles_oracle.c with OpenMP over particles; {dt:.2f} s wall",
                ms_per_step=dt / steps * 1e3, substeps_per_particle_step=c["n_substeps"] / max(c["n_integrated"], 1))


_REAL_STDOUT = None


def claim_stdout():
    """Everything libraries print on fd 1 (NCCL's version banner, …) goes to stderr; the one
    JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=4096)
    ap.add_argument("--ny", type=int, default=4096, help="rows per GPU")
    ap.add_argument("--halo", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--solver", default="Tsit5", choices=["Tsit5", "DP5", "AutoTsit5"],
                    help="ODESettings.solver; AutoTsit5 = the reference's default AutoTsit5(Rosenbrock23()): the same "
                         "arithmetic as Tsit5 on this workload (the stiffness monitor never fires) plus the monitor")
    args = ap.parse_args()
    global SOLVER
    SOLVER = args.solver
    if args.warmup < 3:
        log("warmup raised to 3 (timing rules)")
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"homogeneous box {args.nx}x{args.ny * max(world, 1)} Cartesian (BASELINE configs[1]: "
                          f"T04_2D_reg_test_large_grid/bench06 scaled), dx=dy=2km, u=v=10 m/s, DT=600 s, "
                          f"{args.solver} abstol=1e-4 reltol=1e-3 dt=1e-3 dtmin=1e-4 force_dtmin",
              "nx": args.nx, "ny_per_gpu": args.ny, "parallelism": f"y-strips x{max(world, 1)}",
              "l2_policy": "inputs_exceed_l2 (3.3 GB of per-node planes per GPU >> 126 MB L2)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb = cpu_arm(args.steps, args.warmup, args.cpu_sample)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "Julia is not installed on this image: the reference's CPU path is timed as its C port "
                        "(oracle/), all host threads; BASELINE.md quotes 4-7e4 particle-steps/s for the Julia original"}
        emit(line)
        return 0

    # ---- our arm -------------------------------------------------------------------
    import build_lib
    if rank == 0 or world == 1:
        build_lib.build()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    from picles_b200.engine import B200Engine
    W = workload(args.nx, args.ny, world, rank)
    P = params()
    halo = args.halo if world > 1 else 0
    eng = B200Engine(W["Nx"], W["Ny"], 0, 0, W["mask"], P, M_const=W["M_const"], device=local_rank, j0=W["j0"],
                     ny_local=W["ny"], halo=halo)
    n_nodes = W["Nx"] * W["ny"]
    stepper = None
    if world > 1:
        from picles_b200.distributed import StripStepper
        stepper = StripStepper(eng, rank, world, periodic_y=False)

    fp64_peak = eng.measure_fp64_peak()
    hbm_meas = eng.measure_hbm_copy(2048)
    hbm_peak, hbm_src = measured_peaks()

    def do_step(t, host_ptrs=None):
        if stepper is None:
            if host_ptrs is None:
                eng.step(t, 600.0)
            else:
                eng.step_raw(t, 600.0, None, None, host_ptrs[0], host_ptrs[1])
        else:
            stepper.step(t, 600.0, host_ptrs)

    def barrier():
        eng.synchronize()
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    eng.seed(10.0, 10.0)
    t = 0.0
    for _ in range(args.warmup):
        do_step(t)
        t += 600.0

    # ---- timed region A: inputs resident in HBM -----------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    per_step = []
    barrier()
    eng.timer_start()
    for _ in range(args.steps):
        do_step(t)
        t += 600.0
        per_step.append(eng.counters())
    ms_total = eng.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    n_active = sum(c["n_active"] for c in per_step)  # particle-steps of this rank in the region

    # ---- the reference's default solver on the same workload ---------------------------------
    # ODESettings.solver defaults to AutoTsit5(Rosenbrock23()) (particle_waves_v5.jl:47).  On this
    # workload its stiffness monitor never fires (n_stiff_switches = 0, tested), so its results are
    # the Tsit5 results bit for bit; the monitor-carrying kernel is timed here for the record.
    auto_variant = None
    if world == 1 and args.solver == "Tsit5" and not args.no_e2e:
        eng2 = B200Engine(W["Nx"], W["Ny"], 0, 0, W["mask"], params("AutoTsit5"),
                          M_const=W["M_const"], device=local_rank)
        eng2.seed(10.0, 10.0)
        t2 = 0.0
        for _ in range(args.warmup):
            eng2.step(t2, 600.0)
            t2 += 600.0
        k2 = min(args.steps, 5)
        rows2 = []
        eng2.synchronize()
        eng2.timer_start()
        for _ in range(k2):
            eng2.step(t2, 600.0)
            t2 += 600.0
            rows2.append(eng2.counters())
        ms2 = eng2.timer_stop()
        auto_variant = {"solver": "AutoTsit5(Rosenbrock23())", "steps": k2,
                        "value": sum(r["n_active"] for r in rows2) / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2 / k2,
                        "ms_advance": float(np.mean([r["ms_advance"] for r in rows2])),
                        "n_stiff_switches": int(sum(r["n_stiff_switches"] for r in rows2)),
                        "n_stiff_attempts": int(sum(r["n_stiff_attempts"] for r in rows2))}
        eng2.close()

    # ---- timed region B: end to end through host buffers ----------------------------
    e2e = None
    if not args.no_e2e:
        import torch
        # pinned host staging of the step's inputs: wind at t+DT evaluated on the host mesh
        hu = torch.full((W["ny"], W["Nx"]), 10.0, dtype=torch.float64).pin_memory()
        hv = torch.full((W["ny"], W["Nx"]), 10.0, dtype=torch.float64).pin_memory()
        ptrs = (ctypes.c_void_p(hu.data_ptr()), ctypes.c_void_p(hv.data_ptr()))
        do_step(t, ptrs)
        t += 600.0
        barrier()
        t0 = time.perf_counter()
        eng.timer_start()
        n_e2e = 0
        for _ in range(args.steps):
            do_step(t, ptrs)           # H2D of u,v at t+DT (2 planes), step, D2H of the counters
            t += 600.0
            _ = eng.energy_sum()       # D2H read of the step's result: sum of the energy plane
            n_e2e += eng.counters()["n_active"]
        ms_e2e = eng.timer_stop()
        barrier()
        wall_e2e = (time.perf_counter() - t0) * 1e3
        ms_e2e = max(ms_e2e, wall_e2e)  # host-side work (API, staging) counts end to end
        e2e = dict(ms=ms_e2e, n=n_e2e, h2d=2 * n_nodes * 8, d2h=ctypes.sizeof(ctypes.c_double) * 1024 + 88)

    # ---- timed region C: end to end with a device-resident wind mesh (wind ingestion) ----------
    # Same workload; the wind comes from a gridded field kept on the device (9 x 9 x 3 knots of
    # u = v = 10 m/s, the reference's LinearInterpolation((x,y,t), U, Periodic()) closure) and is
    # sampled at the nodes by k_wind_sample every step: no wind bytes cross PCIe.
    e2e_mesh = None
    if not args.no_e2e:
        xw = np.linspace(-1.0e4, 2000.0 * W["Nx"] + 1.0e4, 9)
        yw = np.linspace(-1.0e4, 2000.0 * W["Ny"] + 1.0e4, 9)
        tw = np.array([0.0, 1.0e6, 2.0e6])
        Uw = np.full((3, 9, 9), 10.0)
        nx_ = np.broadcast_to(np.arange(W["Nx"]) * 2000.0, (W["ny"], W["Nx"]))
        ny_ = np.broadcast_to(((W["j0"] + np.arange(W["ny"])) * 2000.0)[:, None], (W["ny"], W["Nx"]))
        eng.set_wind_mesh(xw, yw, tw, Uw, Uw, nx_, ny_)
        lo, hi = (stepper.lo, stepper.hi) if stepper is not None else (-1, -1)
        eng.step_wind_mesh(t, 600.0, 0, lo, hi)
        t += 600.0
        barrier()
        t0 = time.perf_counter()
        eng.timer_start()
        n_mesh = 0
        for _ in range(args.steps):
            eng.step_wind_mesh(t, 600.0, 0, lo, hi)
            t += 600.0
            _ = eng.energy_sum()
            n_mesh += eng.counters()["n_active"]
        ms_mesh = eng.timer_stop()
        barrier()
        ms_mesh = max(ms_mesh, (time.perf_counter() - t0) * 1e3)
        e2e_mesh = dict(ms=ms_mesh, n=n_mesh, ms_sample=eng.measure_wind_sample(t, 20))

    # ---- reduce over ranks: max time, sum of work ------------------------------------
    if dist is not None:
        import torch
        tt = torch.tensor([ms_total, e2e["ms"] if e2e else 0.0, e2e_mesh["ms"] if e2e_mesh else 0.0],
                          dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ww = torch.tensor([float(n_active), float(e2e["n"]) if e2e else 0.0, float(e2e_mesh["n"]) if e2e_mesh else 0.0],
                          dtype=torch.float64, device="cuda")
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
        ms_total, ms_e2e_all, ms_mesh_all = tt.tolist()
        n_active_all, n_e2e_all, n_mesh_all = ww.tolist()
    else:
        ms_e2e_all = e2e["ms"] if e2e else 0.0
        ms_mesh_all = e2e_mesh["ms"] if e2e_mesh else 0.0
        n_active_all, n_e2e_all = float(n_active), float(e2e["n"]) if e2e else 0.0
        n_mesh_all = float(e2e_mesh["n"]) if e2e_mesh else 0.0

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- rooflines from the live CUDA-event times of the timed region ------------------
    ms_adv = float(np.mean([c["ms_advance"] for c in per_step]))
    ms_prj = float(np.mean([c["ms_project"] for c in per_step]))
    std_size = (args.nx, args.ny) == (4096, 4096)  # the size the committed ncu captures were taken at
    rhs = float(np.mean([c["n_rhs"] for c in per_step]))
    att = float(np.mean([c["n_substeps"] + c["n_rejects"] for c in per_step]))
    integ = float(np.mean([c["n_integrated"] for c in per_step]))
    dep = float(np.mean([c["n_deposited"] for c in per_step]))
    flops = rhs * F_RHS + att * F_ATTEMPT + integ * F_INITDT + dep * F_DEPOSIT
    adv_tf = flops / (ms_adv * 1e-3) / 1e12
    prj_gbs = n_nodes * B_PROJECT_REMESH / (ms_prj * 1e-3) / 1e9
    roofline = {"kernel": "k_advance", "bound": "fp64", "achieved": adv_tf, "peak": fp64_peak,
                "unit": "TFLOP/s", "frac": adv_tf / fp64_peak, "traffic": NCU_TRAFFIC.get("k_advance") if std_size else None,
                "peak_source": "DFMA-chain microbenchmark run in this process (picles_measure_fp64_peak); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "flop_model": {"F_RHS": F_RHS, "F_ATTEMPT": F_ATTEMPT, "F_INITDT": F_INITDT, "F_DEPOSIT": F_DEPOSIT,
                               "rhs_per_launch": rhs, "attempts_per_launch": att},
                "ncu_fp64_pipe_active_pct": NCU_FP64_PIPE_PCT,
                "ms_per_launch": ms_adv, "share_of_step": ms_adv / (ms_adv + ms_prj)}
    roofline_hbm = [
        {"kernel": "k_project_remesh", "bound": "hbm", "achieved": prj_gbs, "peak": hbm_peak, "unit": "GB/s",
         "frac": prj_gbs / hbm_peak, "traffic": NCU_TRAFFIC.get("k_project_remesh") if std_size else None,
         "bytes_per_node": B_PROJECT_REMESH, "nodes_per_launch": n_nodes, "ms_per_launch": ms_prj,
         "share_of_step": ms_prj / (ms_adv + ms_prj), "peak_source": hbm_src},
    ]

    value = n_active_all / (ms_total * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "measured_here": {"fp64_dfma_tflops": fp64_peak, "hbm_copy_gbs": hbm_meas},
            "clocks": clocks,
            # ours, per step of the resident-input region: k_advance + k_project_remesh; on strips the
            # advance is three launches (two boundary blocks first, then the interior) plus
            # k_halo_pack and k_halo_unpack around NCCL's own send/recv kernel
            "gpu_launches": (2 if world == 1 else 6) * args.steps,
            "substeps_per_particle_step": float(np.mean([c["n_substeps"] / max(c["n_integrated"], 1) for c in per_step])),
            "max_attempts": int(max(c["max_attempts"] for c in per_step)),
            "rejects": int(sum(c["n_rejects"] for c in per_step)),
            "failed": int(sum(c["n_failed"] for c in per_step))}
    if e2e:
        line["e2e"] = {"value": n_e2e_all / (ms_e2e_all * 1e-3), "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                       "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": ms_e2e_all / args.steps,
                       "api": "picles_step through the C ABI with pinned host wind buffers + picles_state_energy_sum"}
    if auto_variant:
        line["default_solver_variant"] = auto_variant
    if e2e_mesh:
        # the same metric with the wind ingested on the device; k_wind_sample: 16 B of node
        # coordinates in + 16 B of wind out per node (the mesh itself is L2-resident)
        gbs = n_nodes * 32 / (e2e_mesh["ms_sample"] * 1e-3) / 1e9
        line["e2e_wind_mesh"] = {"value": n_mesh_all / (ms_mesh_all * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0,
                                 "d2h_bytes_per_step": e2e["d2h"] if e2e else 0, "ms_per_step": ms_mesh_all / args.steps,
                                 "api": "picles_step_wind_mesh (wind sampled on the device from a resident wind mesh) "
                                        "+ picles_state_energy_sum"}
        line["roofline_hbm"].append({"kernel": "k_wind_sample", "note": "k_wind_timeblend (mesh-sized, ~3 us) + k_wind_sample, timed as a pair",
                                     "bound": "hbm", "achieved": gbs, "peak": hbm_peak,
                                     "unit": "GB/s", "frac": gbs / hbm_peak,
                                     "traffic": NCU_TRAFFIC.get("k_wind_sample") if std_size else None, "bytes_per_node": 32,
                                     "nodes_per_launch": n_nodes, "ms_per_launch": e2e_mesh["ms_sample"],
                                     "share_of_step": e2e_mesh["ms_sample"] / (ms_adv + ms_prj + e2e_mesh["ms_sample"]),
                                     "peak_source": hbm_src})
    if not args.no_cpu_baseline:
        cb = cpu_arm(min(args.steps, 3), args.warmup, args.cpu_sample)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
