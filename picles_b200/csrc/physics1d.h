/*
 * physics1d.h — per-particle and per-node arithmetic of the reference's ONE-DIMENSIONAL model (WaveGrowth1D)
 * for the picles1d_* kernels.  SURVEY §8f-4, the last "next" row.  One-dimensional grids are small (tens to
 * thousands of nodes), so this path is written for exactness and brevity, not for the FP64 roofline: IEEE
 * operators throughout (the *_safe instantiation of pmath.h), state in local arrays.
 *
 * Reference call sites (paths relative to /root/reference/src):
 *   p1_rhs          particle_equations(u_wind; ...)            ParticleSystems/particle_waves_v5.jl:584-650
 *   p1_windsea      get_initial_windsea(U10, T)                FetchRelations.jl:254-287
 *   p1_windsea2x    get_initial_windsea(U10, V10, T)           FetchRelations.jl:314-359 (seeding, core_1D.jl:196-206)
 *   p1_charge       GetParticleEnergyMomentum                  Operators/core_1D.jl:98-107
 *   p1_vertex       GetVariablesAtVertex                       Operators/core_1D.jl:118-124
 *   p1_integrate    step!(integ, DT, true), auto_dt_reset!     Operators/mapping_1D.jl:107 [OrdinaryDiffEq]
 *   p1_advance      advance!                                   Operators/mapping_1D.jl:84-190
 *   p1_merge        merge!(grid_point, charge) V1              ParticleInCell.jl:228-252
 *   p1_gather_node  push_to_grid! (1-D) as a gather            ParticleInCell.jl:562-590, 163-172
 *   p1_remesh       NodeToParticle!                            Operators/mapping_1D.jl:225-283
 * Behaviour reproduced as written, and what is not modelled: DESIGN.md (quirk table, 1-D rows).
 */
#ifndef PICLES_PHYSICS1D_H
#define PICLES_PHYSICS1D_H

#include "../../include/picles_b200.h"
#include "pmath.h"

#if defined(__CUDACC__)
#define P1_HD __host__ __device__ inline
#else
#define P1_HD static inline
#endif

namespace picles1d {

/* device view of one 1-D model (all arrays of Nx elements unless noted) */
struct Arrays {
    int Nx;
    double xmin, dx;      /* OneDGrid.xmin, .dx: the frame of the weights */
    const double* xn;     /* OneDGridNotes.x: node coordinates, the particles' frame */
    double *z0, *z1, *z2; /* lne, c̄_x, x */
    double *t, *dt, *qold;
    int64_t* iter;
    uint8_t* flags;       /* PICLES_PF_ON | _BOUNDARY | _DT_RESET | _ACTIVE */
    int32_t* status;
    const double *w0, *w1; /* wind at the nodes, levels t and t + DT */
    double* S;             /* State (Nx, 3) column-major */
    /* deposit records written by the advance kernel */
    double *r_e, *r_m, *r_wf, *r_wc;
    int64_t* r_ifl;        /* floor node (1-based, not wrapped); INT64_MIN: no deposit */
};
#define P1_NO_DEPOSIT INT64_MIN

struct Counters {
    unsigned long long n_integrated, n_substeps, n_rejects, n_rhs, n_reseed_advance, n_fixups, n_failed, n_deposited, n_A,
        n_B, n_D;
    int reach, max_attempts;
};

struct Particle1 {
    double u[3];
    double t, dt, qold;
    int64_t iter;
    uint8_t flags;
    int32_t status;
};

/* ---- FetchRelations ---------------------------------------------------------------------- */
P1_HD void p1_windsea(double U10, double time_scale, double& lne, double& cg_bar) {
    time_scale = fabs(time_scale);
    const double aU = fabs(U10);
    const double tau = 9.81 * time_scale / aU;
    const double sgn = (U10 > 0.0) ? 1.0 : ((U10 < 0.0) ? -1.0 : 0.0);
    const double X_tilde = pm_pow(tau / (22.8013 * 2.4097), 1.0 / (1.0 - 0.2748));
    const double f_m = 3.5 * (9.81 / aU) * pm_pow(X_tilde, -0.33);
    const double a_j = 0.033 * pm_pow(f_m * aU / 9.81, 0.67);
    const double w = f_m * 2.0 * 3.141592653589793;
    const double iw = 1.0 / w;
    const double iw2 = iw * iw;
    const double E = 0.31 * (9.81 * 9.81) * a_j * (iw2 * iw2);
    const double f_peak = f_m * 9.81 / aU;
    const double T_bar = 0.9 * (1.0 / f_peak);
    const double cg_amp = 9.81 * T_bar / (4.0 * 3.141592653589793);
    cg_bar = sgn * cg_amp;
    lne = pm_log(E);
}
P1_HD void p1_windsea2x(double U10, double V10, double time_scale, double& lne, double& cgx) {
    double U_amp = sqrt(U10 * U10 + V10 * V10);
    U_amp = (U_amp < 0.1) ? 0.1 : U_amp;
    time_scale = fabs(time_scale);
    const double tau = 9.81 * time_scale / fabs(U_amp);
    const double X_tilde = pm_pow(tau / (22.8013 * 2.4097), 1.0 / (1.0 - 0.2748));
    const double f_m = 3.5 * (9.81 / U_amp) * pm_pow(X_tilde, -0.33);
    const double a_j = 0.033 * pm_pow(f_m * U_amp / 9.81, 0.67);
    const double w = f_m * 2.0 * 3.141592653589793;
    const double iw = 1.0 / w;
    const double iw2 = iw * iw;
    const double E = 0.31 * (9.81 * 9.81) * a_j * (iw2 * iw2);
    const double f_peak = f_m * 9.81 / U_amp;
    const double T_bar = 0.9 * (1.0 / f_peak);
    const double cg_amp = 9.81 * T_bar / (4.0 * 3.141592653589793);
    cgx = cg_amp * U10 / U_amp;
    lne = pm_log(E);
}

/* ---- particle <-> node ----------------------------------------------------------------------- */
P1_HD void p1_charge(const double* u, double& e, double& m) {
    e = pm_exp(u[0]);
    m = e / u[1] / 2.0;
}
P1_HD void p1_vertex(double e, double m, double x, double* u) {
    u[0] = pm_log(e);
    u[1] = e / 2.0 / m;
    u[2] = x;
}

/* ---- right-hand side -------------------------------------------------------------------------- */
P1_HD void p1_rhs(const picles_params_t& P, const double* z, double u, double* dz) {
    const double lne = z[0], cx = z[1];
    const double r_g = P.r_g;
    const double us = fabs(u);
    const double c_gp = fabs(cx) / r_g;
    const double kp = 9.81 / (4.0 * pm_max(c_gp * c_gp, 1e-2));
    const double wp = 9.81 / (2.0 * pm_max(fabs(c_gp), 0.1));
    const double a = us / (2.0 * c_gp);
    const double alpha = (a > 500.0) ? 500.0 : a;
    const double Hp = 0.5 * (1.0 + pm_tanh(P.p * (alpha - 0.85)));
    const double sch = pm_sech(10.0 * (alpha - 0.85));
    const double Dp = 1.0 - 1.25 * (sch * sch);
    double It = 0.0, Dt = 0.0, Scg = 0.0;
    if (P.input) It = P.C_e * Hp * (alpha * alpha);
    if (P.dissipation) {
        const double r = kp / P.e_T;
        const double twon = 2.0 * P.n;
        double pw;
        if (twon == 4.0) { const double r2 = r * r; pw = r2 * r2; }
        else if (twon == 2.0) pw = r * r;
        else pw = pm_pow(r, twon);
        Dt = pm_exp(P.n * lne) * pw;
    }
    if (P.peak_shift) {
        const double k2 = kp * kp;
        Scg = P.C_alpha * Dp * (k2 * k2) * pm_exp(2.0 * lne);
    }
    dz[0] = wp * r_g * Scg + wp * (It - Dt);
    dz[1] = -cx * wp * r_g * Scg;
    dz[2] = P.propagation ? cx : 0.0;
}

/* wind at the particle's position x and stage time ts from the two node levels of the step: linear in time
   (fraction of the particle's own step), then linear in x between the bracketing nodes; constant beyond the
   ends of a non-periodic grid, the wrap cell between node Nx and node 1 on a periodic one */
struct WindCtx {
    const double *w0, *w1, *xn;
    int Nx, periodic;
    double t_start, inv_DT;
};
P1_HD double p1_wind_at(const WindCtx& c, double x, double ts) {
    const int Nx = c.Nx;
    const double s = (ts - c.t_start) * c.inv_DT;
    const double dxn = c.xn[1] - c.xn[0];
    const double xi = (x - c.xn[0]) / dxn;
    const double fl = floor(xi);
    double fr = xi - fl;
    int64_t i0, i1;
    if (c.periodic) {
        double m = fmod(fl, (double)Nx);
        if (m < 0.0) m += (double)Nx;
        i0 = (int64_t)m;
        i1 = (i0 + 1 == Nx) ? 0 : i0 + 1;
    } else {
        if (fl < 0.0) { i0 = 0; fr = 0.0; }
        else if (fl > (double)(Nx - 2)) { i0 = Nx - 2; fr = 1.0; }
        else i0 = (int64_t)fl;
        i1 = i0 + 1;
    }
    if (!(xi == xi)) { i0 = 0; i1 = 1; fr = xi; }
    const double a = fma(c.w1[i0] - c.w0[i0], s, c.w0[i0]);
    const double b = fma(c.w1[i1] - c.w0[i1], s, c.w0[i1]);
    return fma(b - a, fr, a);
}

/* ---- adaptive Runge-Kutta over DT (Tsit5 / DP5; PI controller, Hairer initial step): the same restatement of
   OrdinaryDiffEq as physics.h / SURVEY A.2, on three components ---------------------------------------------- */
struct Tab1 {
    double c[7], a[8][7], bt[8], beta1, beta2;
};
#define P1_TABLEAUS                                                                                              \
    {                                                                                                            \
        {{0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0},                                                   \
         {{0}, {0},                                                                                              \
          {0, 0.161},                                                                                            \
          {0, -0.008480655492356989, 0.335480655492357},                                                         \
          {0, 2.8971530571054935, -6.359448489975075, 4.3622954328695815},                                       \
          {0, 5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525},                 \
          {0, 5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383}, \
          {0, 0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,              \
           2.324710524099774}},                                                                                  \
         {0, -0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,         \
          0.5823571654525552, -0.45808210592918697, 0.015151515151515152},                                       \
         0.14, 0.08},                                                                                            \
        {{0, 0.2, 0.3, 0.8, 8.0 / 9.0, 1.0, 1.0},                                                                \
         {{0}, {0},                                                                                              \
          {0, 0.2},                                                                                              \
          {0, 3.0 / 40.0, 9.0 / 40.0},                                                                           \
          {0, 44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0},                                                            \
          {0, 19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0},                            \
          {0, 9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0},                \
          {0, 35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}},                 \
         {0, -71.0 / 57600.0, 0.0, 71.0 / 16695.0, -71.0 / 1920.0, 17253.0 / 339200.0, -22.0 / 525.0,            \
          1.0 / 40.0},                                                                                           \
         0.17, 0.04}                                                                                             \
    }
#if defined(__CUDACC__)
__constant__ Tab1 d_tab1[2] = P1_TABLEAUS;
#endif
static const Tab1 h_tab1[2] = P1_TABLEAUS;
P1_HD const Tab1& p1_tableau(int solver) {
    const int k = (solver == PICLES_SOLVER_DP5) ? 1 : 0;
#if defined(__CUDA_ARCH__)
    return d_tab1[k];
#else
    return h_tab1[k];
#endif
}

struct Tally1 {
    int integrated, substeps, rejects, rhs, reseed, fixups, failed, deposited, A, B, D, reach, max_attempts;
};

P1_HD double p1_rms3(const double* x) {
    double s = 0.0;
    for (int i = 0; i < 3; i++) s += x[i] * x[i];
    return sqrt(s / 3.0);
}
P1_HD void p1_f(const picles_params_t& P, const WindCtx& w, const double* z, double ts, double* dz, int& nrhs) {
    p1_rhs(P, z, p1_wind_at(w, z[2], ts), dz);
    nrhs++;
}
P1_HD double p1_initdt(const picles_params_t& P, const WindCtx& w, const double* u0, double t, const double* f0, int& nrhs) {
    const double dtmin = pm_nextfloat_pos(P.dtmin);
    const double smalldt = 1e-6;
    double sk[3], tmp[3];
    for (int i = 0; i < 3; i++) sk[i] = fma(fabs(u0[i]), P.reltol, P.abstol);
    for (int i = 0; i < 3; i++) tmp[i] = u0[i] / sk[i];
    const double d0 = p1_rms3(tmp);
    for (int i = 0; i < 3; i++) tmp[i] = f0[i] / sk[i];
    const double d1 = p1_rms3(tmp);
    if (d1 != d1) return dtmin;
    double dt0 = ((d0 < 1e-5) | (d1 < 1e-5)) ? smalldt : (d0 / d1) / 100.0;
    dt0 = pm_min(dt0, P.dtmax);
    if (dt0 < 10.0 * 2.220446049250313e-16) return pm_max(smalldt, dtmin);
    double u1[3], f1v[3];
    for (int i = 0; i < 3; i++) u1[i] = fma(dt0, f0[i], u0[i]);
    p1_f(P, w, u1, t + dt0, f1v, nrhs);
    bool same = true;
    for (int i = 0; i < 3; i++) same = same && (f0[i] == f1v[i]);
    if (same) return pm_max(dtmin, 100.0 * dt0);
    for (int i = 0; i < 3; i++) tmp[i] = (f1v[i] - f0[i]) / sk[i];
    const double d2 = p1_rms3(tmp) / dt0;
    const double mx = pm_max(d1, d2);
    double dt1;
    if (mx <= 1e-15) dt1 = pm_max(1e-6, dt0 * 1e-3);
    else dt1 = pm_exp10(-(2.0 + pm_log10(mx)) / 5.0);
    return pm_max(dtmin, pm_min(pm_min(100.0 * dt0, dt1), P.dtmax));
}

P1_HD void p1_integrate(const picles_params_t& P, const WindCtx& w, double DT, Particle1& p, Tally1& c) {
    const Tab1& T = p1_tableau(P.solver);
    if (p.status & (PICLES_PST_MAXITERS | PICLES_PST_DTMIN | PICLES_PST_UNSTABLE)) return;
    double t = p.t;
    const double tstop = t + DT;
    double u[3], k[8][3];
    int nrhs = 0;
    for (int i = 0; i < 3; i++) u[i] = p.u[i];
    p1_f(P, w, u, t, k[1], nrhs);
    double dt = p.dt;
    if (p.flags & PICLES_PF_DT_RESET) {
        dt = p1_initdt(P, w, u, t, k[1], nrhs);
        p.flags &= (uint8_t)~PICLES_PF_DT_RESET;
    }
    double qold = p.qold;
    int64_t iter = p.iter;
    int attempts = 0;
    const double qmin = 0.2, qmax = 10.0, gamma = 0.9;
    while (t < tstop) {
        iter++;
        const double dtmin_t = pm_max(pm_eps(t), P.dtmin);
        dt = pm_min(P.dtmax, dt);
        dt = pm_max(dt, dtmin_t);
        dt = pm_min(dt, tstop - t);
        if (dt != dt) { p.status |= PICLES_PST_UNSTABLE; c.failed++; break; }
        if (iter > P.maxiters) { p.status |= PICLES_PST_MAXITERS; c.failed++; break; }
        if (!P.force_dtmin && dt <= P.dtmin && (t + dt < tstop)) { p.status |= PICLES_PST_DTMIN; c.failed++; break; }
        attempts++;
        double tmp[3], un[3];
        {
            const double a = dt * T.a[2][1];
            for (int i = 0; i < 3; i++) tmp[i] = fma(a, k[1][i], u[i]);
            p1_f(P, w, tmp, fma(T.c[1], dt, t), k[2], nrhs);
        }
        for (int s = 3; s <= 7; s++) {
            for (int i = 0; i < 3; i++) {
                double inner = T.a[s][1] * k[1][i];
                for (int j = 2; j < s; j++)
                    if (T.a[s][j] != 0.0) inner = fma(T.a[s][j], k[j][i], inner);
                tmp[i] = fma(dt, inner, u[i]);
            }
            const double ts = (s >= 6) ? (t + dt) : fma(T.c[s - 1], dt, t);
            p1_f(P, w, tmp, ts, k[s], nrhs);
            if (s == 7) for (int i = 0; i < 3; i++) un[i] = tmp[i];
        }
        double r[3];
        for (int i = 0; i < 3; i++) {
            double inner = T.bt[1] * k[1][i];
            for (int j = 2; j <= 7; j++)
                if (T.bt[j] != 0.0) inner = fma(T.bt[j], k[j][i], inner);
            const double ut = dt * inner;
            const double sc = fma(pm_max(fabs(u[i]), fabs(un[i])), P.reltol, P.abstol);
            r[i] = ut / sc;
        }
        const double EEst = p1_rms3(r);
        double q, q11 = 1.0;
        if (EEst == 0.0) {
            q = 1.0 / qmax;
        } else {
            const double t1 = T.beta1 * pm_log(EEst);
            q11 = pm_exp(t1);
            q = pm_exp(t1 - T.beta2 * pm_log(qold));
            q = pm_max(1.0 / qmax, pm_min(1.0 / qmin, q / gamma));
        }
        const bool accept = (EEst <= 1.0) || (P.force_dtmin && fabs(dt) <= dtmin_t);
        if (accept) {
            qold = pm_max(EEst, 1e-4);
            const double dtnew = dt / q;
            const double ttmp = t + dt;
            t = (fabs(ttmp - tstop) < 100.0 * pm_eps(tstop)) ? tstop : ttmp;
            double dtp = pm_min(P.dtmax, dtnew);
            dtp = pm_max(dtp, pm_max(pm_eps(t), P.dtmin));
            dt = dtp;
            for (int i = 0; i < 3; i++) { u[i] = un[i]; k[1][i] = k[7][i]; }
            c.substeps++;
            if ((u[0] != u[0]) | (u[1] != u[1]) | (u[2] != u[2])) { p.status |= PICLES_PST_UNSTABLE; c.failed++; break; }
        } else {
            dt = dt / pm_reject_factor(P.nan_eest_rejects, q11, qmin, gamma);
            c.rejects++;
        }
    }
    for (int i = 0; i < 3; i++) p.u[i] = u[i];
    p.t = t; p.dt = dt; p.qold = qold; p.iter = iter;
    c.rhs += nrhs;
    c.integrated++;
    if (attempts > c.max_attempts) c.max_attempts = attempts;
}

/* ---- advance! for particle i (0-based); leaves the deposit record ---------------------------------------- */
P1_HD void p1_reset_values(const Arrays& A, int i, double u, double DT, double* z) {
    p1_windsea(u, DT, z[0], z[1]);
    z[2] = A.xn[i];
}
P1_HD void p1_advance(const Arrays& A, const picles_params_t& P, int i, double DT, Particle1& p, Tally1& c, double& r_e,
                      double& r_m, double& r_wf, double& r_wc, int64_t& r_ifl) {
    WindCtx w;
    w.w0 = A.w0; w.w1 = A.w1; w.xn = A.xn; w.Nx = A.Nx; w.periodic = P.periodic_boundary;
    w.t_start = p.t; w.inv_DT = 1.0 / DT;
    const double t_start = p.t;
    const bool on = (p.flags & PICLES_PF_ON) != 0, boundary = (p.flags & PICLES_PF_BOUNDARY) != 0;
    r_ifl = P1_NO_DEPOSIT;
    r_e = r_m = r_wf = r_wc = 0.0;
    if (on && !boundary) {
        p1_integrate(P, w, DT, p, c);
    } else if (!on && !boundary) {
        const double wind_end = p1_wind_at(w, A.xn[i], t_start + DT);
        if (wind_end * wind_end >= P.wind_min_squared) {
            p1_reset_values(A, i, wind_end, DT, p.u);
            p.flags |= PICLES_PF_DT_RESET | PICLES_PF_ON;
            c.reseed++;
        }
    } else {
        p.flags &= (uint8_t)~PICLES_PF_ON;
        return;
    }
    const bool isn = (p.u[0] != p.u[0]) | (p.u[1] != p.u[1]) | (p.u[2] != p.u[2]);
    const bool isi = pm_isinf(p.u[0]) | pm_isinf(p.u[1]) | pm_isinf(p.u[2]);
    if (isn) {
        p1_reset_values(A, i, p1_wind_at(w, A.xn[i], t_start + DT), DT, p.u);
        p.flags |= PICLES_PF_DT_RESET; p.status |= PICLES_PST_NAN_RESET; c.fixups++;
    } else if (isi) {
        p1_reset_values(A, i, p1_wind_at(w, A.xn[i], t_start), DT, p.u);
        p.flags |= PICLES_PF_DT_RESET; p.status |= PICLES_PST_INF_RESET; c.fixups++;
    } else if (p.u[0] > P.log_energy_maximum) {
        p1_reset_values(A, i, p1_wind_at(w, A.xn[i], t_start), DT, p.u);
        p.flags |= PICLES_PF_DT_RESET; p.status |= PICLES_PST_EMAX_CLAMP; c.fixups++;
    }
    if (p.flags & PICLES_PF_ON) {
        /* ParticleToNode!: charge, floor node and weights; the merge itself is done per target node */
        const double xnrm = (p.u[2] - A.xmin) / A.dx;
        if (!(fabs(xnrm) < 1.0e9)) return;
        const double base = floor(xnrm);
        const int64_t ifl = (int64_t)(base + 1.0);
        const double wc = xnrm - base, wf = 1.0 - wc;
        p1_charge(p.u, r_e, r_m);
        r_wf = wf; r_wc = wc; r_ifl = ifl;
        /* does any corner land?  (counter only; reach for the gather's window) */
        bool any = false;
        for (int k = 0; k < 2; k++) {
            const int64_t im = ifl + k;
            any = any || P.periodic_boundary || (im <= A.Nx && im > 0);
        }
        if (any) c.deposited++;
        const int64_t home = i + 1;
        int64_t d0 = ifl - home, d1 = ifl + 1 - home;
        d0 = d0 < 0 ? -d0 : d0; d1 = d1 < 0 ? -d1 : d1;
        const int64_t d = d0 > d1 ? d0 : d1;
        const int dd = d > 1000000000 ? 1000000000 : (int)d;
        if (dd > c.reach) c.reach = dd;
    }
}

/* ---- merge!(grid_point, charge), as typed ------------------------------------------------------------------- */
P1_HD void p1_merge(double* g, const double* ch) {
    const double dE = g[0] - ch[0];
    double cosv;
    const double ng = sqrt(g[1] * g[1] + g[2] * g[2]);
    if (ng == 0.0) cosv = 1.0;
    else {
        const double nc = sqrt(ch[1] * ch[1] + ch[2] * ch[2]);
        cosv = g[1] * ch[1] + g[2] * g[2] / (ng * nc);
    }
    if (cosv >= 0.5) { g[0] += ch[0]; g[1] += ch[1]; g[2] += ch[2]; }
    else if ((cosv < 0.5) && (dE > 0.0)) { }
    else if ((cosv > 0.5) && (dE <= 0.0)) { g[0] = ch[0]; g[1] = ch[1]; g[2] = ch[2]; }
}
P1_HD int64_t p1_wrap_index(int64_t pos, int64_t N) {
    pos = pos % N;
    if (pos < 0) pos += N;
    else if (pos == 0) pos += N;
    return pos;
}
/* one candidate particle i (0-based) merged into node n (1-based) if one of its two corners lands there */
P1_HD void p1_merge_candidate(const Arrays& A, int periodic, int64_t i, int64_t n, double* g) {
    const int64_t ifl = A.r_ifl[i];
    if (ifl == P1_NO_DEPOSIT) return;
    for (int k = 0; k < 2; k++) {
        int64_t im = ifl + k;
        if (periodic) im = p1_wrap_index(im, A.Nx);
        else if (!(im <= A.Nx && im > 0)) continue;
        if (im != n) continue;
        const double wgt = k ? A.r_wc[i] : A.r_wf[i];
        const double ch[3] = {wgt * A.r_e[i], wgt * A.r_m[i], wgt * 0.0};
        p1_merge(g, ch);
    }
}
/* State at node n (1-based): the charges that reach it, merged in the reference's order — particles 1..Nx, each
   its floor corner then its ceil corner.  R = the step's reach (max |corner - home| over all deposits): only
   particles within R nodes can land here; their indices form one ascending run, or two when the window wraps. */
P1_HD void p1_gather_node(const Arrays& A, int periodic, int R, int64_t n, double* g) {
    g[0] = g[1] = g[2] = 0.0;
    const int64_t Nx = A.Nx;
    if (2 * (int64_t)R + 1 >= Nx) {
        for (int64_t i = 0; i < Nx; i++) p1_merge_candidate(A, periodic, i, n, g);
        return;
    }
    const int64_t lo = n - R, hi = n + R; /* 1-based particle numbers, unwrapped */
    if (!periodic) {
        for (int64_t q = (lo < 1 ? 1 : lo); q <= (hi > Nx ? Nx : hi); q++) p1_merge_candidate(A, periodic, q - 1, n, g);
    } else if (lo < 1) {
        for (int64_t q = 1; q <= hi; q++) p1_merge_candidate(A, periodic, q - 1, n, g);
        for (int64_t q = lo + Nx; q <= Nx; q++) p1_merge_candidate(A, periodic, q - 1, n, g);
    } else if (hi > Nx) {
        for (int64_t q = 1; q <= hi - Nx; q++) p1_merge_candidate(A, periodic, q - 1, n, g);
        for (int64_t q = lo; q <= Nx; q++) p1_merge_candidate(A, periodic, q - 1, n, g);
    } else {
        for (int64_t q = lo; q <= hi; q++) p1_merge_candidate(A, periodic, q - 1, n, g);
    }
}

/* ---- NodeToParticle! ---------------------------------------------------------------------------------------------- */
P1_HD void p1_remesh(const Arrays& A, const picles_params_t& P, int i, double DT, const double* s, double u_wind, Particle1& p,
                     Tally1& c) {
    const bool boundary = (p.flags & PICLES_PF_BOUNDARY) != 0;
    if (!boundary && (s[0] >= P.minimal_state[0]) && (s[1] * s[1] >= P.minimal_state[1])) {
        p1_vertex(s[0], s[1], A.xn[i], p.u);
        p.flags |= PICLES_PF_DT_RESET | PICLES_PF_ON;
        c.A++;
    } else if (!boundary && (u_wind * u_wind >= P.wind_min_squared)) {
        p1_reset_values(A, i, u_wind, DT, p.u);
        p.qold = 1e-4; p.iter = 0; p.status = 0;
        p.flags |= PICLES_PF_DT_RESET | PICLES_PF_ON;
        c.B++;
    } else {
        p.flags &= (uint8_t)~PICLES_PF_ON;
        c.D++;
    }
}

/* ---- SeedParticle! -------------------------------------------------------------------------------------------------- */
P1_HD void p1_seed(const Arrays& A, const picles_params_t& P, int i, double u, Particle1& p, double* s) {
    const double x = A.xn[i];
    bool on;
    if (fabs(u) > sqrt(2.0)) {
        p1_windsea2x(u, 0.0, P.seed_timescale, p.u[0], p.u[1]);
        on = true;
    } else {
        const double U = (u == 0.0) ? 1.0 : u, V = 1.0;
        const double Uamp = sqrt(U * U + V * V);
        p1_windsea2x(1.0 * U / Uamp, 1.0 * V / Uamp, P.seed_timescale, p.u[0], p.u[1]);
        on = false;
    }
    p.u[2] = x;
    const bool boundary = P.periodic_boundary ? false : (i == 0 || i == A.Nx - 1);
    p.flags = (uint8_t)(PICLES_PF_ACTIVE | (on ? PICLES_PF_ON : 0) | (boundary ? PICLES_PF_BOUNDARY : 0));
    s[0] = s[1] = s[2] = 0.0;
    if (on) { p1_charge(p.u, s[0], s[1]); s[2] = 0.0; }
    p.t = 0.0; p.dt = P.dt; p.qold = 1e-4; p.iter = 0; p.status = 0;
}

} /* namespace picles1d */
#endif /* PICLES_PHYSICS1D_H */
