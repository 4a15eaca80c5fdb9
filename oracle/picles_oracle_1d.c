/*
 * picles_oracle_1d.c — CPU restatement of the reference's ONE-DIMENSIONAL model (SURVEY §8f-4:
 * WaveGrowth1D), the checker of the picles1d_* CUDA path.  TEST INFRASTRUCTURE ONLY: nothing of the
 * product links, loads or calls this file.  PARITY UNPINNED against the real reference (no Julia here, and
 * the reference holds no fixtures for the 1-D path either).
 *
 * What it follows (paths relative to /root/reference/src):
 *   rhs1                particle_equations(u_wind; ...)          ParticleSystems/particle_waves_v5.jl:584-650
 *   windsea1 / seeding  get_initial_windsea(U10, T)              FetchRelations.jl:254-287
 *                       InitParticleValues / SeedParticle!       Operators/core_1D.jl:178-222, 270-330
 *   charge / vertex     GetParticleEnergyMomentum, GetVariablesAtVertex   Operators/core_1D.jl:98-124
 *   weights             compute_weights_and_index(::OneDGrid, x) ParticleInCell.jl:163-172, 33-49
 *   deposit             push_to_grid! (1-D, "version B") + merge! V1      ParticleInCell.jl:562-590, 228-252
 *   advance1            advance!                                 Operators/mapping_1D.jl:84-190
 *   remesh1             remesh! / NodeToParticle!                Operators/mapping_1D.jl:197-283
 *   step                State .= 0; time_step!                   Simulations/run.jl:72-80, Operators/TimeSteppers.jl:51-92
 *   integrate1          step!(integ, DT, true), auto_dt_reset!   [OrdinaryDiffEq, un-vendored: restated as in
 *                                                                 picles_oracle.c / SURVEY A.2, three components]
 *
 * As-written behaviour that is reproduced on purpose:
 *   - the deposit is NOT a sum: merge!(grid_point, charge) adds the charge only while the node is empty or
 *     m_grid * m_charge >= 0.5 (the un-normalised "cos theta" of the formula as typed, with m_y = 0); for
 *     oceanic momenta (1e-4 .. 1) that means the FIRST charge to arrive, in particle order, owns the node;
 *   - particles keep their own clock: one that is off does not advance it, so after k steps off it lags by
 *     k*DT and advance! tests the wind at its own t + DT.  With staged wind levels the oracle (and the
 *     device) use the model-clock levels instead; the closure mode below follows the reference literally;
 *   - the wind is sampled at the particle's CURRENT position u_wind(x, t) (2-D: at the home node).  Staged
 *     mode interpolates the node values linearly in x and t (exact for winds that are linear between nodes
 *     and over a step); closure mode calls the closure;
 *   - the NaN branch of advance! references an undefined variable (`@show winds_start`) and would throw;
 *     the evident intent (reseed from the wind at t_end) is implemented;
 *   - OneDGridNotes.x starts at 0 whatever grid.xmin is (ParticleMesh.jl:131) while the weights use
 *     (x - xmin)/dx: both are taken as given (node coordinates and xmin, dx are separate inputs).
 * Not modelled: ParticleDefaults seeding (every particle would sit at defaults.x), AutoTsit5's switch to
 * Rosenbrock23 (runs as Tsit5), layers > 1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/picles_b200.h"
#include "../picles_b200/csrc/pmath.h"

#define QOLDINIT 1e-4

typedef struct {
    double u[3]; /* lne, c̄_x, x */
    double t, dt, qold;
    int64_t iter;
    uint8_t on, boundary, dt_reset;
    int32_t status;
} p1_t;

typedef struct oracle1d {
    int Nx;
    double xmin, dx; /* OneDGrid.xmin, OneDGrid.dx: the weights' frame */
    double* xn;      /* OneDGridNotes.x: node coordinates, the particles' frame */
    picles_params_t P;
    p1_t* part;
    double* S; /* (Nx, 3) column-major */
    picles_counters_t C;
    double (*wind_fn)(double x, double t);
    /* staged winds of the current step */
    const double *w0, *w1;
    double t_clock, inv_DT;
} oracle1d_t;

/* ---- FetchRelations ---------------------------------------------------------------------- */
/* get_initial_windsea(U10, time_scale), FetchRelations.jl:254-287 ("JONSWAP") */
static void windsea1(double U10, double time_scale, double* lne, double* cg_bar) {
    time_scale = fabs(time_scale);
    double tau = 9.81 * time_scale / fabs(U10);
    double sgn = (U10 > 0.0) ? 1.0 : ((U10 < 0.0) ? -1.0 : 0.0);
    double X_tilde = pm_pow(tau / (22.8013 * 2.4097), 1.0 / (1.0 - 0.2748)); /* X_tilde_from_tau :128-130 */
    double f_m = 3.5 * (9.81 / fabs(U10)) * pm_pow(X_tilde, -0.33);         /* :165-167 */
    double a_j = 0.033 * pm_pow(f_m * fabs(U10) / 9.81, 0.67);             /* :184-186 */
    double w = f_m * 2.0 * 3.141592653589793;                              /* E_JONSWAP :201-203, ^(-4) = inv(x)^4 */
    double iw = 1.0 / w;
    double iw2 = iw * iw;
    double E = 0.31 * (9.81 * 9.81) * a_j * (iw2 * iw2);
    double f_peak = f_m * 9.81 / fabs(U10);
    double T_bar = 0.9 * (1.0 / f_peak);
    double cg_amp = 9.81 * T_bar / (4.0 * 3.141592653589793);
    *cg_bar = sgn * cg_amp;
    *lne = pm_log(E);
}
/* get_initial_windsea(U10, V10, time_scale), FetchRelations.jl:314-359: lne and cg_bar_x */
static void windsea2x(double U10, double V10, double time_scale, double* lne, double* cgx) {
    double U_amp = sqrt(U10 * U10 + V10 * V10);
    U_amp = (U_amp < 0.1) ? 0.1 : U_amp;
    time_scale = fabs(time_scale);
    double tau = 9.81 * time_scale / fabs(U_amp);
    double X_tilde = pm_pow(tau / (22.8013 * 2.4097), 1.0 / (1.0 - 0.2748));
    double f_m = 3.5 * (9.81 / U_amp) * pm_pow(X_tilde, -0.33);
    double a_j = 0.033 * pm_pow(f_m * U_amp / 9.81, 0.67);
    double w = f_m * 2.0 * 3.141592653589793;
    double iw = 1.0 / w;
    double iw2 = iw * iw;
    double E = 0.31 * (9.81 * 9.81) * a_j * (iw2 * iw2);
    double f_peak = f_m * 9.81 / U_amp;
    double T_bar = 0.9 * (1.0 / f_peak);
    double cg_amp = 9.81 * T_bar / (4.0 * 3.141592653589793);
    *cgx = cg_amp * U10 / U_amp;
    *lne = pm_log(E);
}

/* ---- core_1D ------------------------------------------------------------------------------- */
static void charge1(const double* u, double* ch) { /* core_1D.jl:98-107 */
    double e = pm_exp(u[0]);
    ch[0] = e;
    ch[1] = e / u[1] / 2.0;
    ch[2] = 0.0;
}
static void vertex1(const double* s, double x, double* u) { /* core_1D.jl:118-124 */
    u[0] = pm_log(s[0]);
    u[1] = s[0] / 2.0 / s[1];
    u[2] = x;
}

/* ---- right-hand side, particle_waves_v5.jl:597-646 ------------------------------------------ */
static void rhs1(const picles_params_t* P, const double* z, double u, double* dz) {
    double lne = z[0], cx = z[1];
    double r_g = P->r_g;
    double us = fabs(u);
    double c_gp = fabs(cx) / r_g; /* c_g_conversions_vector(abs(c̄)), :281-287 */
    double kp = 9.81 / (4.0 * pm_max(c_gp * c_gp, 1e-2));
    double wp = 9.81 / (2.0 * pm_max(fabs(c_gp), 0.1));
    double a = us / (2.0 * c_gp); /* α_func :215-225 */
    double alpha = (a > 500.0) ? 500.0 : a;
    double Hp = 0.5 * (1.0 + pm_tanh(P->p * (alpha - 0.85))); /* H_β(α, p) :274 */
    double sch = pm_sech(10.0 * (alpha - 0.85));               /* Δ_β(α) :275 */
    double Dp = 1.0 - 1.25 * (sch * sch);
    double It = 0.0, Dt = 0.0, Scg = 0.0;
    if (P->input) It = P->C_e * Hp * (alpha * alpha);
    if (P->dissipation) {
        double r = kp / P->e_T, pw;
        double twon = 2.0 * P->n;
        if (twon == 4.0) { double r2 = r * r; pw = r2 * r2; }
        else if (twon == 2.0) pw = r * r;
        else pw = pm_pow(r, twon);
        Dt = pm_exp(P->n * lne) * pw;
    }
    if (P->peak_shift) {
        double k2 = kp * kp;
        Scg = P->C_alpha * Dp * (k2 * k2) * pm_exp(2.0 * lne);
    }
    dz[0] = wp * r_g * Scg + wp * (It - Dt);
    dz[1] = -cx * wp * r_g * Scg;
    dz[2] = P->propagation ? cx : 0.0;
}

/* wind at (x, ts).  Closure mode: the closure.  Staged mode: the two node levels of the step, linear in time
   (fraction of the particle's own step: ts - t_start over DT), then linear in x between the two nodes that
   bracket x (constant beyond the ends of a non-periodic grid; the wrap cell between node Nx and node 1 on a
   periodic one). */
typedef struct {
    const oracle1d_t* o;
    double t_start;
} ctx1_t;
static double wind_at(const ctx1_t* c, double x, double ts) {
    const oracle1d_t* o = c->o;
    if (o->wind_fn) return o->wind_fn(x, ts);
    const int Nx = o->Nx;
    double s = (ts - c->t_start) * o->inv_DT;
    double dxn = o->xn[1] - o->xn[0];
    double xi = (x - o->xn[0]) / dxn;
    double fl = floor(xi);
    double fr = xi - fl;
    int64_t i0, i1;
    if (o->P.periodic_boundary) {
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
    if (!(xi == xi)) { i0 = 0; i1 = 1; fr = xi; } /* NaN position: NaN wind */
    double a = fma(o->w1[i0] - o->w0[i0], s, o->w0[i0]);
    double b = fma(o->w1[i1] - o->w0[i1], s, o->w0[i1]);
    return fma(b - a, fr, a);
}
static void f1(const ctx1_t* c, const double* z, double ts, double* dz, int64_t* nrhs) {
    rhs1(&c->o->P, z, wind_at(c, z[2], ts), dz);
    (*nrhs)++;
}

/* ---- OrdinaryDiffEq restatement, three components ------------------------------ */
typedef struct {
    double c[7], a[8][7], bt[8], beta1, beta2;
} tab_t;
static const tab_t TSIT5 = {
    {0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0},
    {{0}, {0},
     {0, 0.161},
     {0, -0.008480655492356989, 0.335480655492357},
     {0, 2.8971530571054935, -6.359448489975075, 4.3622954328695815},
     {0, 5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525},
     {0, 5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383},
     {0, 0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774}},
    {0, -0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
     0.5823571654525552, -0.45808210592918697, 0.015151515151515152},
    0.14, 0.08};
static const tab_t DP5 = {
    {0, 0.2, 0.3, 0.8, 8.0 / 9.0, 1.0, 1.0},
    {{0}, {0},
     {0, 0.2},
     {0, 3.0 / 40.0, 9.0 / 40.0},
     {0, 44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0},
     {0, 19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0},
     {0, 9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0},
     {0, 35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}},
    {0, -71.0 / 57600.0, 0.0, 71.0 / 16695.0, -71.0 / 1920.0, 17253.0 / 339200.0, -22.0 / 525.0, 1.0 / 40.0},
    0.17, 0.04};

static double rms3(const double* x) {
    double s = 0.0;
    for (int i = 0; i < 3; i++) s += x[i] * x[i];
    return sqrt(s / 3.0);
}
/* ode_determine_initdt */
static double initdt1(const ctx1_t* c, const double* u0, double t, const double* f0, int64_t* nrhs) {
    const picles_params_t* P = &c->o->P;
    double dtmin = pm_nextfloat_pos(P->dtmin);
    const double smalldt = 1e-6;
    double sk[3], tmp[3];
    for (int i = 0; i < 3; i++) sk[i] = fma(fabs(u0[i]), P->reltol, P->abstol);
    for (int i = 0; i < 3; i++) tmp[i] = u0[i] / sk[i];
    double d0 = rms3(tmp);
    for (int i = 0; i < 3; i++) tmp[i] = f0[i] / sk[i];
    double d1 = rms3(tmp);
    if (d1 != d1) return dtmin;
    double dt0 = ((d0 < 1e-5) | (d1 < 1e-5)) ? smalldt : (d0 / d1) / 100.0;
    dt0 = pm_min(dt0, P->dtmax);
    if (dt0 < 10.0 * 2.220446049250313e-16) return pm_max(smalldt, dtmin);
    double u1[3], f1v[3];
    for (int i = 0; i < 3; i++) u1[i] = fma(dt0, f0[i], u0[i]);
    f1(c, u1, t + dt0, f1v, nrhs);
    int same = 1;
    for (int i = 0; i < 3; i++) same &= (f0[i] == f1v[i]);
    if (same) return pm_max(dtmin, 100.0 * dt0);
    for (int i = 0; i < 3; i++) tmp[i] = (f1v[i] - f0[i]) / sk[i];
    double d2 = rms3(tmp) / dt0;
    double mx = pm_max(d1, d2);
    double dt1;
    if (mx <= 1e-15) dt1 = pm_max(1e-6, dt0 * 1e-3);
    else dt1 = pm_exp10(-(2.0 + pm_log10(mx)) / 5.0);
    return pm_max(dtmin, pm_min(pm_min(100.0 * dt0, dt1), P->dtmax));
}

/* step!(integrator, DT, true) */
static void integrate1(oracle1d_t* o, p1_t* p, const ctx1_t* c, double DT) {
    const picles_params_t* P = &o->P;
    picles_counters_t* C = &o->C;
    const tab_t* T = (P->solver == PICLES_SOLVER_DP5) ? &DP5 : &TSIT5;
    if (p->status & (PICLES_PST_MAXITERS | PICLES_PST_DTMIN | PICLES_PST_UNSTABLE)) return;
    double t = p->t, tstop = t + DT;
    double u[3], k[8][3];
    int64_t nrhs = 0;
    memcpy(u, p->u, sizeof u);
    f1(c, u, t, k[1], &nrhs); /* u_modified -> reset_fsal! */
    double dt = p->dt;
    if (p->dt_reset) { dt = initdt1(c, u, t, k[1], &nrhs); p->dt_reset = 0; }
    double qold = p->qold;
    int64_t iter = p->iter;
    int attempts = 0;
    const double qmin = 0.2, qmax = 10.0, gamma = 0.9;
    while (t < tstop) {
        iter++;
        double dtmin_t = pm_max(pm_eps(t), P->dtmin);
        dt = pm_min(P->dtmax, dt);
        dt = pm_max(dt, dtmin_t);
        dt = pm_min(dt, tstop - t);
        if (dt != dt) { p->status |= PICLES_PST_UNSTABLE; C->n_failed++; break; }
        if (iter > P->maxiters) { p->status |= PICLES_PST_MAXITERS; C->n_failed++; break; }
        if (!P->force_dtmin && dt <= P->dtmin && (t + dt < tstop)) { p->status |= PICLES_PST_DTMIN; C->n_failed++; break; }
        attempts++;
        double tmp[3], un[3];
        {
            double a = dt * T->a[2][1];
            for (int i = 0; i < 3; i++) tmp[i] = fma(a, k[1][i], u[i]);
            f1(c, tmp, fma(T->c[1], dt, t), k[2], &nrhs);
        }
        for (int s = 3; s <= 7; s++) {
            for (int i = 0; i < 3; i++) {
                double inner = T->a[s][1] * k[1][i];
                for (int j = 2; j < s; j++)
                    if (T->a[s][j] != 0.0) inner = fma(T->a[s][j], k[j][i], inner);
                tmp[i] = fma(dt, inner, u[i]);
            }
            double ts = (s >= 6) ? (t + dt) : fma(T->c[s - 1], dt, t);
            f1(c, tmp, ts, k[s], &nrhs);
            if (s == 7) memcpy(un, tmp, sizeof un);
        }
        double r[3];
        for (int i = 0; i < 3; i++) {
            double inner = T->bt[1] * k[1][i];
            for (int j = 2; j <= 7; j++)
                if (T->bt[j] != 0.0) inner = fma(T->bt[j], k[j][i], inner);
            double ut = dt * inner;
            double sc = fma(pm_max(fabs(u[i]), fabs(un[i])), P->reltol, P->abstol);
            r[i] = ut / sc;
        }
        double EEst = rms3(r);
        double q, q11 = 1.0;
        if (EEst == 0.0) {
            q = 1.0 / qmax;
        } else {
            double t1 = T->beta1 * pm_log(EEst);
            q11 = pm_exp(t1);
            q = pm_exp(t1 - T->beta2 * pm_log(qold));
            q = pm_max(1.0 / qmax, pm_min(1.0 / qmin, q / gamma));
        }
        int accept = (EEst <= 1.0) || (P->force_dtmin && fabs(dt) <= dtmin_t);
        if (accept) {
            qold = pm_max(EEst, QOLDINIT);
            double dtnew = dt / q;
            double ttmp = t + dt;
            t = (fabs(ttmp - tstop) < 100.0 * pm_eps(tstop)) ? tstop : ttmp;
            double dtp = pm_min(P->dtmax, dtnew);
            dtp = pm_max(dtp, pm_max(pm_eps(t), P->dtmin));
            dt = dtp;
            memcpy(u, un, sizeof u);
            memcpy(k[1], k[7], sizeof k[1]);
            C->n_substeps++;
            if ((u[0] != u[0]) | (u[1] != u[1]) | (u[2] != u[2])) { p->status |= PICLES_PST_UNSTABLE; C->n_failed++; break; }
        } else {
            if (P->nan_eest_rejects && q11 != q11) dt = dt / (1.0 / qmin); /* fastpow reading of a NaN error estimate: see picles_b200.h */
            else dt = dt / pm_min(1.0 / qmin, q11 / gamma);
            C->n_rejects++;
        }
    }
    memcpy(p->u, u, sizeof u);
    p->t = t; p->dt = dt; p->qold = qold; p->iter = iter;
    C->n_rhs += nrhs;
    C->n_integrated++;
    if (attempts > C->max_attempts) C->max_attempts = attempts;
}

/* ---- ParticleInCell, 1-D ----------------------------------------------------------------------- */
static int64_t wrap_index(int64_t pos, int64_t N) { /* ParticleInCell.jl:444-454 */
    pos = pos % N;
    if (pos < 0) pos += N;
    else if (pos == 0) pos += N;
    return pos;
}
/* merge!(grid_point, charge), ParticleInCell.jl:228-252, as typed */
static void merge_v1(double* g, const double* c) {
    double dE = g[0] - c[0];
    double cosv;
    double ng = sqrt(g[1] * g[1] + g[2] * g[2]);
    if (ng == 0.0) cosv = 1.0;
    else {
        double nc = sqrt(c[1] * c[1] + c[2] * c[2]);
        cosv = g[1] * c[1] + g[2] * g[2] / (ng * nc);
    }
    if (cosv >= 0.5) { g[0] += c[0]; g[1] += c[1]; g[2] += c[2]; }
    else if ((cosv < 0.5) && (dE > 0.0)) { /* forget the charge */ }
    else if ((cosv > 0.5) && (dE <= 0.0)) { g[0] = c[0]; g[1] = c[1]; g[2] = c[2]; }
}
static void particle_to_node1(oracle1d_t* o, const p1_t* p) {
    const int Nx = o->Nx;
    double ch[3];
    charge1(p->u, ch);
    double xnrm = (p->u[2] - o->xmin) / o->dx; /* compute_weights_and_index :163-172 */
    if (!(fabs(xnrm) < 1.0e9)) return;         /* Int(...) would throw: nothing deposited (fork) */
    double base = floor(xnrm);                 /* get_absolute_i_and_w(zp) :33-49: no rounding to 6 digits here */
    int64_t ifl = (int64_t)(base + 1.0);
    double wc = xnrm - base, wf = 1.0 - wc;
    int64_t idx[2] = {ifl, ifl + 1};
    double w[2] = {wf, wc};
    int any = 0;
    for (int k = 0; k < 2; k++) {
        int64_t im = idx[k];
        if (o->P.periodic_boundary) im = wrap_index(im, Nx);
        else if (!(im <= Nx && im > 0)) continue;
        double c[3] = {w[k] * ch[0], w[k] * ch[1], w[k] * ch[2]};
        double g[3] = {o->S[im - 1], o->S[im - 1 + Nx], o->S[im - 1 + 2 * (int64_t)Nx]};
        merge_v1(g, c);
        o->S[im - 1] = g[0]; o->S[im - 1 + Nx] = g[1]; o->S[im - 1 + 2 * (int64_t)Nx] = g[2];
        any = 1;
    }
    if (any) o->C.n_deposited++;
}

/* ---- model ------------------------------------------------------------------------------------ */
oracle1d_t* oracle1d_create(int Nx, double xmin, double dx, const double* x_nodes, const picles_params_t* P) {
    if (Nx < 2 || !x_nodes || !P) return NULL;
    oracle1d_t* o = (oracle1d_t*)calloc(1, sizeof *o);
    o->Nx = Nx; o->xmin = xmin; o->dx = dx;
    o->xn = (double*)malloc(sizeof(double) * Nx);
    memcpy(o->xn, x_nodes, sizeof(double) * Nx);
    o->P = *P;
    o->part = (p1_t*)calloc(Nx, sizeof(p1_t));
    o->S = (double*)calloc((size_t)3 * Nx, sizeof(double));
    return o;
}
void oracle1d_destroy(oracle1d_t* o) {
    if (!o) return;
    free(o->xn); free(o->part); free(o->S); free(o);
}
void oracle1d_set_wind_closure(oracle1d_t* o, double (*fn)(double, double)) { o->wind_fn = fn; }

/* ResetParticleValues(nothing, PI, u, DT), core_1D.jl:237-262 */
static void reset_values1(const oracle1d_t* o, int i, double u, double DT, double* z) {
    windsea1(u, DT, &z[0], &z[1]);
    z[2] = o->xn[i];
}

/* init_particles! / SeedParticle!, core_1D.jl:270-330 */
void oracle1d_seed(oracle1d_t* o, const double* u0) {
    const picles_params_t* P = &o->P;
    memset(o->S, 0, sizeof(double) * 3 * o->Nx);
    memset(&o->C, 0, sizeof o->C);
    for (int i = 0; i < o->Nx; i++) {
        p1_t* p = &o->part[i];
        memset(p, 0, sizeof *p);
        double x = o->xn[i];
        double u = o->wind_fn ? o->wind_fn(x, 0.0) : u0[i];
        int on;
        if (fabs(u) > sqrt(2.0)) { /* InitParticleValues :196-206 */
            windsea2x(u, 0.0, P->seed_timescale, &p->u[0], &p->u[1]);
            on = 1;
        } else { /* MinimalParticle(u, 0.0, DT): rand_sign() -> +1 (B-9) */
            double U = (u == 0.0) ? 1.0 : u, V = 1.0;
            double Uamp = sqrt(U * U + V * V);
            windsea2x(1.0 * U / Uamp, 1.0 * V / Uamp, P->seed_timescale, &p->u[0], &p->u[1]);
            on = 0;
        }
        p->u[2] = x;
        p->boundary = P->periodic_boundary ? 0 : (uint8_t)(i == 0 || i == o->Nx - 1);
        p->on = (uint8_t)on;
        if (on) {
            double ch[3];
            charge1(p->u, ch);
            o->S[i] = ch[0]; o->S[i + o->Nx] = ch[1]; o->S[i + 2 * (int64_t)o->Nx] = ch[2];
        }
        p->t = 0.0; p->dt = P->dt; p->qold = QOLDINIT; p->iter = 0; p->dt_reset = 0;
    }
}

/* advance!, mapping_1D.jl:84-190 */
static void advance1(oracle1d_t* o, int i, double DT) {
    const picles_params_t* P = &o->P;
    p1_t* p = &o->part[i];
    ctx1_t c = {o, p->t};
    const double t_start = p->t;
    if (p->on && !p->boundary) {
        integrate1(o, p, &c, DT); /* a stopped integrator (maxiters, dtmin, NaN) is a status, not an exception:
                                     the checks and the deposit below still run, as on the 2-D path */
    } else if (!p->on && !p->boundary) {
        double wind_end = wind_at(&c, o->xn[i], t_start + DT);
        if (wind_end * wind_end >= P->wind_min_squared) {
            reset_values1(o, i, wind_end, DT, p->u);
            p->dt_reset = 1;
            p->on = 1;
            o->C.n_reseed_advance++;
        }
    } else {
        p->on = 0;
        return;
    }
    int isn = (p->u[0] != p->u[0]) | (p->u[1] != p->u[1]) | (p->u[2] != p->u[2]);
    int isi = pm_isinf(p->u[0]) | pm_isinf(p->u[1]) | pm_isinf(p->u[2]);
    if (isn) {
        reset_values1(o, i, wind_at(&c, o->xn[i], t_start + DT), DT, p->u);
        p->dt_reset = 1; p->status |= PICLES_PST_NAN_RESET; o->C.n_fixups++;
    } else if (isi) {
        reset_values1(o, i, wind_at(&c, o->xn[i], t_start), DT, p->u);
        p->dt_reset = 1; p->status |= PICLES_PST_INF_RESET; o->C.n_fixups++;
    } else if (p->u[0] > P->log_energy_maximum) {
        reset_values1(o, i, wind_at(&c, o->xn[i], t_start), DT, p->u);
        p->dt_reset = 1; p->status |= PICLES_PST_EMAX_CLAMP; o->C.n_fixups++;
    }
    if (p->on) particle_to_node1(o, p);
}

/* remesh! / NodeToParticle!, mapping_1D.jl:197-283; u_wind = winds(x_node, clock.time) (pre-tick) */
static void remesh1(oracle1d_t* o, int i, double DT, double u_wind) {
    const picles_params_t* P = &o->P;
    p1_t* p = &o->part[i];
    const int Nx = o->Nx;
    double s[3] = {o->S[i], o->S[i + Nx], o->S[i + 2 * (int64_t)Nx]};
    if (!p->boundary && (s[0] >= P->minimal_state[0]) && (s[1] * s[1] >= P->minimal_state[1])) {
        vertex1(s, o->xn[i], p->u); /* set_u_and_t!(ui, last_t); auto_dt_reset! */
        p->dt_reset = 1;
        p->on = 1;
        o->C.n_remesh_A++;
    } else if (!p->boundary && (u_wind * u_wind >= P->wind_min_squared)) {
        reset_values1(o, i, u_wind, DT, p->u); /* reinit!(...); set_t!(last_t); auto_dt_reset! */
        p->qold = QOLDINIT; p->iter = 0; p->status = 0;
        p->dt_reset = 1;
        p->on = 1;
        o->C.n_remesh_B++;
    } else {
        p->on = 0;
        o->C.n_remesh_D++;
    }
}

/* State .= 0; time_step!(model, DT) with the clock at t */
void oracle1d_step(oracle1d_t* o, double t, double DT, const double* u_t, const double* u_t1) {
    memset(o->S, 0, sizeof(double) * 3 * o->Nx);
    memset(&o->C, 0, sizeof o->C);
    o->w0 = u_t; o->w1 = u_t1; o->t_clock = t; o->inv_DT = 1.0 / DT;
    for (int i = 0; i < o->Nx; i++) advance1(o, i, DT);
    for (int i = 0; i < o->Nx; i++) {
        double uw = o->wind_fn ? o->wind_fn(o->xn[i], t) : u_t[i];
        remesh1(o, i, DT, uw);
    }
    o->C.n_active = o->Nx;
}

void oracle1d_get_state(const oracle1d_t* o, double* S) { memcpy(S, o->S, sizeof(double) * 3 * o->Nx); }
void oracle1d_get_particles(const oracle1d_t* o, double* z, double* t, double* dt, uint8_t* flags, int32_t* status) {
    for (int i = 0; i < o->Nx; i++) {
        const p1_t* p = &o->part[i];
        for (int k = 0; k < 3; k++) z[i + (int64_t)k * o->Nx] = p->u[k];
        if (t) t[i] = p->t;
        if (dt) dt[i] = p->dt;
        if (flags)
            flags[i] = (uint8_t)((p->on ? PICLES_PF_ON : 0) | (p->boundary ? PICLES_PF_BOUNDARY : 0) |
                                 (p->dt_reset ? PICLES_PF_DT_RESET : 0) | PICLES_PF_ACTIVE);
        if (status) status[i] = p->status;
    }
}
void oracle1d_get_counters(const oracle1d_t* o, picles_counters_t* c) { *c = o->C; }

/* hooks for the known-answer tests */
void oracle1d_rhs(const picles_params_t* P, const double* z, double u, double* dz) { rhs1(P, z, u, dz); }
void oracle1d_windsea(double u, double T, double* out2) { windsea1(u, T, &out2[0], &out2[1]); }
void oracle1d_merge(double* g, const double* c) { merge_v1(g, c); }
