/*
 * picles_oracle.c — CPU restatement of the PiCLES per-timestep particle-in-cell path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (picles_b200/) may import, link or
 * execute this file; it is the checker used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference (mochell/PiCLES) contains no assertions, golden
 * vectors or fixtures for this path (tests/runtests.jl:4-6 is an empty testset), Julia
 * is not installed here, and the adaptive Runge–Kutta arithmetic lives in
 * OrdinaryDiffEq.jl, a dependency with no [compat] bound and a git-ignored Manifest
 * (Project.toml:16,50-51; .gitignore:24).  This file restates the published algorithm
 * of OrdinaryDiffEq's Tsit5/DP5 + PIController + Hairer initial step
 * and follows the reference's own call sites line by line; the only reference-held
 * known-answer numbers (src/Grids/mask_utils_test.jl:28-30) are checked in tests/.
 *
 * Structure follows the reference: one heap "ParticleInstance" per node, an ordered
 * `ocean_points` list, a scatter `push_to_grid!` into a zeroed State in list order
 * (the single-thread order of movie_time_step!, TimeSteppers.jl:212-247), then remesh.
 *
 * Transcendentals come from picles_b200/csrc/pmath.h (deterministic, shared with the
 * device so accept/reject decisions cannot flip between CPU and GPU); build with
 * -DORACLE_LIBM to use the system libm instead (cross-check of pmath itself).
 * Compile with -ffp-contract=off: every fused multiply-add below is explicit.
 *
 * Deliberate, documented forks from the as-run reference:
 *   - winds are staged meshes at the step's t and t+DT (plus optional equally spaced
 *     intermediate levels), a polynomial in time through the levels in between (the
 *     reference calls the closure at every RK stage time; oracle_set_wind_closure runs
 *     exactly that, as the yardstick the staged-level runs are measured against);
 *   - auto_dt_reset! is evaluated lazily at the start of the next advance (same
 *     arithmetic, same inputs);
 *   - solver Tsit5: the AutoTsit5 stiffness trigger is only counted; solver AutoTsit5
 *     (PICLES_SOLVER_AUTOTSIT5) runs OrdinaryDiffEq's AutoSwitch + Rosenbrock23 as restated
 *     below (published algorithm, from the call sites; third-party arithmetic unpinned);
 *   - rand_sign() for an exactly-zero wind component (FetchRelations.jl:365) is +1.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/picles_b200.h"
#include "../picles_b200/csrc/pmath.h"
#include "../picles_b200/csrc/pmath_trig.h"
#include "../picles_b200/csrc/pmath_dual.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#ifdef ORACLE_LIBM
#define O_EXP exp
#define O_LOG log
#define O_TANH tanh
#define O_SECH(x) (1.0 / cosh(x))
#define O_POW pow
#define O_LOG10 log10
#define O_EXP10(x) pow(10.0, (x))
#else
#define O_EXP pm_exp
#define O_LOG pm_log
#define O_TANH pm_tanh
#define O_SECH pm_sech
#define O_POW pm_pow
#define O_LOG10 pm_log10
#define O_EXP10 pm_exp10
#endif

#define QOLDINIT 1e-4 /* OrdinaryDiffEq default qoldinit for adaptive algorithms */

#if defined(ORACLE_CTRL_POW)
static double ctrl_pow(double x, double y) {
    float r = powf((float)x, (float)y);
#if ORACLE_CTRL_POW >= 2
    union { float f; uint32_t u; } c;
    c.f = r; c.u &= 0xfffff800u; r = c.f; /* keep 12 of the 23 mantissa bits */
#endif
    return (double)r;
}
#endif

/* ------------------------------------------------------------------------ */
/* types                                                                      */
/* ------------------------------------------------------------------------ */
typedef struct {
    double u[5]; /* lne, c̄_x, c̄_y, x, y          (ODEIntegrator.u) */
    double t;    /* ODEIntegrator.t */
    double dt;   /* ODEIntegrator.dt / dtpropose */
    double qold; /* PI controller memory */
    int64_t iter;
    uint8_t on;       /* custom_structures.jl:18 */
    uint8_t boundary; /* custom_structures.jl:17 */
    uint8_t dt_reset; /* auto_dt_reset! pending */
    uint8_t active;   /* in ocean_points */
    uint8_t exists;   /* mask != 0 (a real integrator was built, core_2D.jl:455-458) */
    int32_t status;
    int32_t stiff_run; /* consecutive stiff-looking steps (AutoTsit5 monitor) */
    int32_t as_count;  /* AutoSwitchCache.count: +n consecutive stiff attempts, -n non-stiff */
    uint8_t as_stiff;  /* AutoSwitchCache.is_stiffalg: Rosenbrock23 is the current algorithm */
} particle_t;

typedef struct oracle {
    int Nx, Ny, bx, by;
    uint8_t* mask;   /* Nx*Ny total mask */
    double* M;       /* 4 planes or NULL */
    double Mc[4];    /* uniform kernel */
    double* pc;      /* plane or NULL */
    picles_params_t P;
    particle_t* part;  /* Nx*Ny */
    int64_t* ocean;    /* ocean_points as linear indices, canonical order */
    int64_t n_ocean;
    double* S;         /* 3 planes */
    picles_counters_t C;
    int64_t stiff_triggers; /* times AutoTsit5 would have switched */
    int nthreads;
    int accumulate; /* 0: State .= 0 before the step (run!); 1: bare time_step! */
    /* intermediate wind levels of the next step (consumed by it): n_mid planes of Nx*Ny */
    int n_mid;
    double *u_mid, *v_mid;
    /* reference semantics: the wind closures u(x,y,t), v(x,y,t) called at the home node and the
       stage time (particle_waves_v5.jl:489-495); x, y = grid.data.x / grid.data.y */
    void (*wind_fn)(double x, double y, double t, double* u, double* v);
    double *x, *y;
    /* B-1 as run (on_persist == 0): a particle seeded off never integrates, so its integrator clock
       stays at 0 and advance! tests the wind at t_end = integ.t + DT = DT on EVERY step
       (mapping_2D.jl:132,172-176).  With staged winds that level is the first step's t+DT level:
       kept here.  (The closure mode simply calls the closure at p->t + DT.) */
    int64_t steps_since_seed;
    double *u_lag, *v_lag;
} oracle_t;

/* ------------------------------------------------------------------------ */
/* FetchRelations.jl                                                          */
/* ------------------------------------------------------------------------ */
/* get_initial_windsea(U10,V10,time_scale; particle_state=true), FetchRelations.jl:314-359 */
static void windsea(double U10, double V10, double time_scale, double* lne, double* cgx, double* cgy,
                    double* E_out, double* cg_amp_out) {
    double U_amp = sqrt(U10 * U10 + V10 * V10); /* :315 */
    U_amp = (U_amp < 0.1) ? 0.1 : U_amp;        /* :316 */
    time_scale = fabs(time_scale);              /* :318 */
    double tau = 9.81 * time_scale / fabs(U_amp); /* :319 */
    /* X_tilde_from_tau, :128-130, constants :107-111 */
    double X_tilde = O_POW(tau / (22.8013 * 2.4097), 1.0 / (1.0 - 0.2748));
    /* fₘ_from_X_tilde, :165-167 */
    double f_m = 3.5 * (9.81 / U_amp) * O_POW(X_tilde, -0.33);
    /* alpha_j, :184-186 */
    double a_j = 0.033 * O_POW(f_m * U_amp / 9.81, 0.67);
    /* E_JONSWAP, :201-203:  0.31 * 9.81^2 * alpha_j * (f_m*2*pi)^(-4); literal -4 -> inv(x)^4 */
    double w = f_m * 2.0 * 3.141592653589793;
    double iw = 1.0 / w;
    double iw2 = iw * iw;
    double E = 0.31 * (9.81 * 9.81) * a_j * (iw2 * iw2);
    double f_peak = f_m * 9.81 / U_amp; /* :334 */
    double T_bar = 0.9 * (1.0 / f_peak); /* :343 */
    double cg_amp = 9.81 * T_bar / (4.0 * 3.141592653589793); /* :344 */
    *cgx = cg_amp * U10 / U_amp; /* :345 */
    *cgy = cg_amp * V10 / U_amp; /* :346 */
    *lne = O_LOG(E);
    if (E_out) *E_out = E;
    if (cg_amp_out) *cg_amp_out = cg_amp;
}

/* MinimalWindsea(U10,V10,T), FetchRelations.jl:381-386; rand_sign() -> +1 */
static void minimal_windsea(double U10, double V10, double T, double* lne, double* cgx, double* cgy,
                            double* E, double* cg_amp, double* ux, double* uy) {
    if (U10 == 0.0) U10 = 1.0;
    if (V10 == 0.0) V10 = 1.0;
    double Uamp = sqrt(U10 * U10 + V10 * V10);
    double a = 1.0 * U10 / Uamp, b = 1.0 * V10 / Uamp; /* u_min = 1.0 */
    windsea(a, b, T, lne, cgx, cgy, E, cg_amp);
    if (ux) *ux = a;
    if (uy) *uy = b;
}

/* ------------------------------------------------------------------------ */
/* core_2D.jl particle <-> node                                               */
/* ------------------------------------------------------------------------ */
/* GetParticleEnergyMomentum, core_2D.jl:69-78 */
static void particle_to_charge(const double* u, double* ch) {
    double e = O_EXP(u[0]);
    double cs = sqrt(u[1] * u[1] + u[2] * u[2]);
    ch[0] = e;
    ch[1] = u[1] * e / (cs * cs) / 2.0;
    ch[2] = u[2] * e / (cs * cs) / 2.0;
}
/* GetVariablesAtVertex, core_2D.jl:121-128 */
static void vertex_to_particle(const double* s, double x, double y, double* u) {
    double e = s[0], mx = s[1], my = s[2];
    double m_amp = sqrt(mx * mx + my * my);
    u[0] = O_LOG(e);
    u[1] = mx * e / (2.0 * (m_amp * m_amp));
    u[2] = my * e / (2.0 * (m_amp * m_amp));
    u[3] = x;
    u[4] = y;
}
/* ResetParticleValues, core_2D.jl:307-343 */
static void reset_particle_values(const picles_params_t* P, double wu, double wv, double DT, double* u) {
    if (!P->has_defaults) {
        windsea(wu, wv, DT, &u[0], &u[1], &u[2], NULL, NULL); /* :321 */
    } else {
        u[0] = P->defaults[0]; u[1] = P->defaults[1]; u[2] = P->defaults[2]; /* :333 */
    }
    u[3] = 0.0; u[4] = 0.0; /* xy = (0,0) on mesh grids, mapping_2D.jl:138,292 */
}

/* ------------------------------------------------------------------------ */
/* particle_waves_v5.jl RHS                                                   */
/* ------------------------------------------------------------------------ */
static inline double alpha_func(double us, double cgp) { /* :215-225 */
    double a = us / (2.0 * cgp);
    return (a > 500.0) ? 500.0 : a;
}

/* particle_system(dz,z,params,t), particle_waves_v5.jl:479-556; (u,v) is the wind
   at the HOME node (:489-495), M row-major [M11 M12; M21 M22], pc = great-circle coef */
static void rhs(const picles_params_t* P, const double* z, double u, double v,
                const double* M, double pc, double* dz) {
    double lne = z[0], cx = z[1], cy = z[2];
    double r_g = P->r_g;
    double cbar = sqrt(cx * cx + cy * cy); /* speed, :297,500 */
    double us = sqrt(u * u + v * v);       /* :501 */
    /* c_g_conversions_vector(abs(c̄)), :281-287 (g = 9.81 default) */
    double c_gp = fabs(cbar) / r_g;
    double kp = 9.81 / (4.0 * pm_max(c_gp * c_gp, 1e-2));
    double wp = 9.81 / (2.0 * pm_max(fabs(c_gp), 0.1));
    double gx = cx / r_g, gy = cy / r_g; /* :505-506 */
    double alpha = alpha_func(us, c_gp); /* :509 */
    /* αₚ, :212 */
    double sg = sqrt(gx * gx + gy * gy);
    double msg = pm_max(sg, 1e-4);
    double alpha_p = (u * gx + v * gy) / (2.0 * (msg * msg));
    /* H_β, Δ_β :274-275 */
    double Hp = 0.5 * (1.0 + O_TANH(P->p * (alpha_p - 0.85)));
    double sch = O_SECH(10.0 * (alpha_p - 0.85));
    double Dp = 1.0 - 1.25 * (sch * sch);
    /* source terms :514-517 */
    double It = 0.0, Dt = 0.0, Scg = 0.0, Sdir = 0.0;
    if (P->input) It = P->C_e * Hp * (alpha * alpha); /* :317-321 */
    if (P->dissipation) {                              /* :331-335 */
        double r = kp / P->e_T, pw;
        double twon = 2.0 * P->n;
        if (twon == 4.0) { double r2 = r * r; pw = r2 * r2; }
        else if (twon == 2.0) pw = r * r;
        else pw = O_POW(r, twon);
        Dt = O_EXP(P->n * lne) * pw;
    }
    if (P->peak_shift) { /* :340  C_α * Δₚ * kₚ^4 * exp(2*lne) */
        double k2 = kp * kp;
        Scg = P->C_alpha * Dp * (k2 * k2) * O_EXP(2.0 * lne);
    }
    if (P->direction) { /* S_dir :345-346, sin2_a_min_b :242-249 */
        double a2 = alpha_func(us, sg);
        double prod = us * sg;
        double s2 = 0.0;
        if (!(prod == 0.0)) {
            s2 = (2.0 / (prod * prod)) *
                 (u * v * (2.0 * (gy * gy) - sg * sg) - gx * gy * (2.0 * (v * v) - us * us));
        }
        Sdir = a2 * a2 * P->C_varphi * Hp * s2;
    }
    double Ssph = cx * pc; /* :521, spherical_grid_corrections.jl:13-18 */
    /* :526-530 */
    dz[0] = wp * r_g * Scg + wp * (It - Dt);
    dz[1] = -cx * wp * r_g * Scg + cy * Sdir + cy * Ssph;
    dz[2] = -cy * wp * r_g * Scg - cx * Sdir - cx * Ssph;
    if (P->propagation) { /* :536 */
        dz[3] = M[0] * cx + M[1] * cy;
        dz[4] = M[2] * cx + M[3] * cy;
    } else {
        dz[3] = 0.0; dz[4] = 0.0;
    }
}

/* The same right-hand side on dual numbers (partials w.r.t. lne, c̄_x, c̄_y and the wind
   components): what ForwardDiff evaluates for Rosenbrock23's Jacobian and time gradient
   (autodiff = true, the default of Rosenbrock23()).  Same operation sequence as rhs(). */
static pmd_t alpha_func_dual(pmd_t us, pmd_t cgp) { /* IfElse.ifelse(a > 500, 500, a), :215-225 */
    pmd_t a = pmd_div(us, pmd_scale(2.0, cgp));
    return (a.v > 500.0) ? pmd_const(500.0) : a;
}
static void rhs_dual(const picles_params_t* P, pmd_t lne, pmd_t cx, pmd_t cy, pmd_t u, pmd_t v, const double* M,
                     double pc, pmd_t* dz) {
    pmd_t r_g = pmd_const(P->r_g);
    pmd_t cbar = pmd_sqrt(pmd_add(pmd_sqr(cx), pmd_sqr(cy)));
    pmd_t us = pmd_sqrt(pmd_add(pmd_sqr(u), pmd_sqr(v)));
    pmd_t c_gp = pmd_div(pmd_abs(cbar), r_g);
    pmd_t kp = pmd_cdiv(9.81, pmd_scale(4.0, pmd_maxc(pmd_sqr(c_gp), 1e-2)));
    pmd_t wp = pmd_cdiv(9.81, pmd_scale(2.0, pmd_maxc(pmd_abs(c_gp), 0.1)));
    pmd_t gx = pmd_div(cx, r_g), gy = pmd_div(cy, r_g);
    pmd_t alpha = alpha_func_dual(us, c_gp);
    pmd_t sg = pmd_sqrt(pmd_add(pmd_sqr(gx), pmd_sqr(gy)));
    pmd_t msg = pmd_maxc(sg, 1e-4);
    pmd_t alpha_p = pmd_div(pmd_add(pmd_mul(u, gx), pmd_mul(v, gy)), pmd_scale(2.0, pmd_sqr(msg)));
    pmd_t Hp = pmd_scale(0.5, pmd_addc(pmd_tanh(pmd_scale(P->p, pmd_addc(alpha_p, -0.85))), 1.0));
    pmd_t sch = pmd_sech(pmd_scale(10.0, pmd_addc(alpha_p, -0.85)));
    pmd_t Dp = pmd_addc(pmd_scale(-1.25, pmd_sqr(sch)), 1.0);
    pmd_t It = pmd_const(0.0), Dt = pmd_const(0.0), Scg = pmd_const(0.0), Sdir = pmd_const(0.0);
    if (P->input) It = pmd_mul(pmd_scale(P->C_e, Hp), pmd_sqr(alpha));
    if (P->dissipation) {
        pmd_t r = pmd_div(kp, pmd_const(P->e_T)), pw;
        double twon = 2.0 * P->n;
        if (twon == 4.0) pw = pmd_sqr(pmd_sqr(r));
        else if (twon == 2.0) pw = pmd_sqr(r);
        else pw = pmd_pow_given(r, twon, pm_pow(r.v, twon), pm_pow(r.v, twon - 1.0));
        Dt = pmd_mul(pmd_exp(pmd_scale(P->n, lne)), pw);
    }
    if (P->peak_shift) Scg = pmd_mul(pmd_mul(pmd_scale(P->C_alpha, Dp), pmd_sqr(pmd_sqr(kp))), pmd_exp(pmd_scale(2.0, lne)));
    if (P->direction) {
        pmd_t a2 = alpha_func_dual(us, sg);
        pmd_t prod = pmd_mul(us, sg);
        pmd_t s2 = pmd_const(0.0);
        if (!(prod.v == 0.0)) {
            pmd_t t1 = pmd_mul(pmd_mul(u, v), pmd_sub(pmd_scale(2.0, pmd_sqr(gy)), pmd_sqr(sg)));
            pmd_t t2 = pmd_mul(pmd_mul(gx, gy), pmd_sub(pmd_scale(2.0, pmd_sqr(v)), pmd_sqr(us)));
            s2 = pmd_mul(pmd_cdiv(2.0, pmd_sqr(prod)), pmd_sub(t1, t2));
        }
        Sdir = pmd_mul(pmd_mul(pmd_scale(P->C_varphi, pmd_sqr(a2)), Hp), s2);
    }
    pmd_t Ssph = pmd_scale(pc, cx);
    pmd_t wrs = pmd_mul(pmd_mul(wp, r_g), Scg);
    dz[0] = pmd_add(wrs, pmd_mul(wp, pmd_sub(It, Dt)));
    dz[1] = pmd_add(pmd_add(pmd_neg(pmd_mul(cx, wrs)), pmd_mul(cy, Sdir)), pmd_mul(cy, Ssph));
    dz[2] = pmd_sub(pmd_sub(pmd_neg(pmd_mul(cy, wrs)), pmd_mul(cx, Sdir)), pmd_mul(cx, Ssph));
    if (P->propagation) {
        dz[3] = pmd_add(pmd_scale(M[0], cx), pmd_scale(M[1], cy));
        dz[4] = pmd_add(pmd_scale(M[2], cx), pmd_scale(M[3], cy));
    } else {
        dz[3] = pmd_const(0.0); dz[4] = pmd_const(0.0);
    }
}

/* ------------------------------------------------------------------------ */
/* OrdinaryDiffEq restatement                                 */
/* ------------------------------------------------------------------------ */
typedef struct {
    double c[7];      /* c[1..6] used: stage times of k2..k7 */
    double a[8][7];   /* a[s][j], s = 2..7, j = 1..s-1 */
    double bt[8];     /* error weights 1..7 */
    double beta1, beta2;
    int order;
} tableau_t;

static const tableau_t TSIT5 = {
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
    0.14, 0.08, 5};

static const tableau_t DP5 = {
    {0, 0.2, 0.3, 0.8, 8.0 / 9.0, 1.0, 1.0},
    {{0}, {0},
     {0, 0.2},
     {0, 3.0 / 40.0, 9.0 / 40.0},
     {0, 44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0},
     {0, 19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0},
     {0, 9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0},
     {0, 35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}},
    {0, -71.0 / 57600.0, 0.0, 71.0 / 16695.0, -71.0 / 1920.0, 17253.0 / 339200.0, -22.0 / 525.0, 1.0 / 40.0},
    0.17, 0.04, 5};

#define WIND_SEG_MAX (PICLES_WIND_MID_MAX + 1)
typedef struct {
    const oracle_t* o;
    const double* M;
    double pc;
    int nseg;                                   /* time segments of the staged wind: levels - 1 */
    double cu[WIND_SEG_MAX + 1], cv[WIND_SEG_MAX + 1]; /* Newton coefficients of the levels in time */
    double t_start, inv_DT;
    double x, y;                                /* home-node coordinates (closure mode) */
} rhs_ctx_t;

/* Newton forward-difference form of the polynomial through nseg+1 equally spaced levels w[k]
   (k = 0: the step's t level, k = nseg: t+DT):  c[m] = Delta^m w_0 / m!.  Two levels: c[1] = w1-w0. */
static void wind_coefficients(const double* w, int nseg, double* c) {
    static const double inv_fact[WIND_SEG_MAX + 1] = {1.0, 1.0, 0.5, 1.0 / 6.0, 1.0 / 24.0};
    double d[WIND_SEG_MAX + 1] = {0};
    for (int k = 0; k <= nseg; k++) d[k] = w[k];
    c[0] = d[0];
    for (int m = 1; m <= nseg; m++) {
        for (int k = 0; k <= nseg - m; k++) d[k] = d[k + 1] - d[k];
        c[m] = (m >= 2) ? d[0] * inv_fact[m] : d[0];
    }
}

/* wind at stage time ts: the polynomial in time through the staged levels (documented fork:
   linear for the two levels t, t+DT), or the closure itself when one is set */
static inline void f_eval(const rhs_ctx_t* c, const double* z, double ts, double* dz, int64_t* nrhs) {
    double u, v;
    if (c->o->wind_fn) {
        c->o->wind_fn(c->x, c->y, ts, &u, &v);
    } else {
        double sg = (ts - c->t_start) * ((double)c->nseg * c->inv_DT);
        double pu = c->cu[c->nseg], pv = c->cv[c->nseg];
        for (int m = c->nseg - 1; m >= 0; m--) {
            double a = sg - (double)m;
            pu = fma(pu, a, c->cu[m]);
            pv = fma(pv, a, c->cv[m]);
        }
        u = pu; v = pv;
    }
    rhs(&c->o->P, z, u, v, c->M, c->pc, dz);
    (*nrhs)++;
}

static inline double rms5(const double* x) { /* ODE_DEFAULT_NORM */
    double s = 0.0;
    for (int i = 0; i < 5; i++) s += x[i] * x[i];
    return sqrt(s / 5.0);
}

/* ode_determine_initdt (Hairer), in-place branch; f0 = f(u0,t) is supplied */
static double initdt(const rhs_ctx_t* c, const double* u0, double t, const double* f0, int64_t* nrhs, double order) {
    const picles_params_t* P = &c->o->P;
    double dtmin = pm_nextfloat_pos(P->dtmin);
    const double smalldt = 1e-6;
    double sk[5], tmp[5];
    for (int i = 0; i < 5; i++) sk[i] = fma(fabs(u0[i]), P->reltol, P->abstol);
    for (int i = 0; i < 5; i++) tmp[i] = u0[i] / sk[i];
    double d0 = rms5(tmp);
    for (int i = 0; i < 5; i++) tmp[i] = f0[i] / sk[i];
    double d1 = rms5(tmp);
    if (d1 != d1) return dtmin;
    double dt0 = ((d0 < 1e-5) | (d1 < 1e-5)) ? smalldt : (d0 / d1) / 100.0;
    dt0 = pm_min(dt0, P->dtmax);
    if (dt0 < 10.0 * 2.220446049250313e-16) return pm_max(smalldt, dtmin);
    double u1[5], f1[5];
    for (int i = 0; i < 5; i++) u1[i] = fma(dt0, f0[i], u0[i]);
    f_eval(c, u1, t + dt0, f1, nrhs);
    int same = 1;
    for (int i = 0; i < 5; i++) same &= (f0[i] == f1[i]);
    if (same) return pm_max(dtmin, 100.0 * dt0);
    for (int i = 0; i < 5; i++) tmp[i] = (f1[i] - f0[i]) / sk[i];
    double d2 = rms5(tmp) / dt0;
    double mx = pm_max(d1, d2);
    double dt1;
    if (mx <= 1e-15) dt1 = pm_max(1e-6, dt0 * 1e-3);
    else dt1 = O_EXP10(-(2.0 + O_LOG10(mx)) / order); /* get_current_alg_order: 5 (Tsit5, DP5), 2 (Rosenbrock23) */
    return pm_max(dtmin, pm_min(pm_min(100.0 * dt0, dt1), P->dtmax));
}

/* ---- Rosenbrock23 (the stiff algorithm of AutoTsit5(Rosenbrock23())) -----------------------
 * OrdinaryDiffEq's Rosenbrock23 (Shampine's ode23s): d = 1/(2+sqrt 2), c32 = 6+sqrt 2,
 *   W = I - dt*d*J,  k1 = W\(f0 + dt*d*dT),  f1 = f(u + dt/2 k1, t + dt/2),
 *   k2 = W\(f1 - k1) + k1,  u1 = u + dt k2,  f2 = f(u1, t + dt),
 *   k3 = W\(f2 - c32 (k2 - f1) - 2 (k1 - f0) + dt dT),  utilde = dt/6 (k1 - 2 k2 + k3),
 * J = df/du and dT = df/dt by automatic differentiation at (u, t) (rhs_dual), the 5x5 system by
 * LU with partial pivoting.  In a composite algorithm the step also leaves eigen_est = opnorm(J, Inf)
 * for the AutoSwitch test.  Third-party arithmetic: the operation order inside OrdinaryDiffEq /
 * LinearSolve is not pinned by the reference. */
static void wind_value_and_rate(const rhs_ctx_t* c, double ts, double* u, double* v, double* ut, double* vt) {
    if (c->o->wind_fn) { /* closure mode: central difference for the time gradient */
        double up, vp, um, vm, h = 1.0;
        c->o->wind_fn(c->x, c->y, ts, u, v);
        c->o->wind_fn(c->x, c->y, ts + h, &up, &vp);
        c->o->wind_fn(c->x, c->y, ts - h, &um, &vm);
        *ut = (up - um) / (2.0 * h); *vt = (vp - vm) / (2.0 * h);
        return;
    }
    double scale = (double)c->nseg * c->inv_DT;
    double sg = (ts - c->t_start) * scale;
    double pu = c->cu[c->nseg], pv = c->cv[c->nseg], du = 0.0, dv = 0.0;
    for (int m = c->nseg - 1; m >= 0; m--) {
        double a = sg - (double)m;
        du = fma(du, a, pu); dv = fma(dv, a, pv);
        pu = fma(pu, a, c->cu[m]); pv = fma(pv, a, c->cv[m]);
    }
    *u = pu; *v = pv;
    *ut = du * scale; *vt = dv * scale;
}

/* LU with partial pivoting, in place; returns 0 if a pivot is exactly zero */
static int lu5(double A[5][5], int* piv) {
    for (int k = 0; k < 5; k++) {
        int p = k;
        double big = fabs(A[k][k]);
        for (int i = k + 1; i < 5; i++)
            if (fabs(A[i][k]) > big) { big = fabs(A[i][k]); p = i; }
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < 5; j++) { double tmp = A[k][j]; A[k][j] = A[p][j]; A[p][j] = tmp; }
        if (A[k][k] == 0.0) return 0;
        for (int i = k + 1; i < 5; i++) {
            A[i][k] = A[i][k] / A[k][k];
            for (int j = k + 1; j < 5; j++) A[i][j] = A[i][j] - A[i][k] * A[k][j];
        }
    }
    return 1;
}
static void lu5_solve(double A[5][5], const int* piv, double* b) {
    for (int k = 0; k < 5; k++) {
        double tmp = b[k]; b[k] = b[piv[k]]; b[piv[k]] = tmp;
        for (int i = k + 1; i < 5; i++) b[i] = b[i] - A[i][k] * b[k];
    }
    for (int k = 4; k >= 0; k--) {
        for (int j = k + 1; j < 5; j++) b[k] = b[k] - A[k][j] * b[j];
        b[k] = b[k] / A[k][k];
    }
}

/* one Rosenbrock23 attempt from (u, t) with f0 = f(u, t); fills unew, f2 = f(unew, t+dt), the
   scaled error estimate and eigen_est; returns 0 when W is singular (EEst = NaN: rejected) */
static void rosenbrock23_attempt(const rhs_ctx_t* c, const double* u, double t, double dt, const double* f0,
                                 double* unew, double* f2, double* EEst, double* eigen_est, int64_t* nrhs) {
    const picles_params_t* P = &c->o->P;
    const double d = 1.0 / (2.0 + sqrt(2.0)), c32 = 6.0 + sqrt(2.0);
    double gam = dt * d, dto2 = dt / 2.0, dto6 = dt / 6.0;
    double wu, wv, wut, wvt;
    wind_value_and_rate(c, t, &wu, &wv, &wut, &wvt);
    pmd_t dz[5];
    rhs_dual(P, pmd_var(u[0], 0), pmd_var(u[1], 1), pmd_var(u[2], 2), pmd_var(wu, 3), pmd_var(wv, 4), c->M, c->pc, dz);
    (*nrhs)++;
    double J[5][5], dT[5], W[5][5];
    double nrm = 0.0;
    for (int i = 0; i < 5; i++) {
        double row = 0.0;
        for (int j = 0; j < 5; j++) {
            J[i][j] = (j < 3) ? dz[i].d[j] : 0.0; /* the system does not depend on the particle position */
            row += fabs(J[i][j]);
        }
        nrm = pm_max(nrm, row);
        dT[i] = dz[i].d[3] * wut + dz[i].d[4] * wvt;
    }
    *eigen_est = nrm;
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) W[i][j] = ((i == j) ? 1.0 : 0.0) - gam * J[i][j];
    int piv[5];
    if (!lu5(W, piv)) {
        for (int i = 0; i < 5; i++) { unew[i] = u[i]; f2[i] = f0[i]; }
        *EEst = pm_nan();
        return;
    }
    double k1[5], k2[5], k3[5], f1[5], tmp[5];
    for (int i = 0; i < 5; i++) k1[i] = f0[i] + gam * dT[i];
    lu5_solve(W, piv, k1);
    for (int i = 0; i < 5; i++) tmp[i] = u[i] + dto2 * k1[i];
    f_eval(c, tmp, t + dto2, f1, nrhs);
    for (int i = 0; i < 5; i++) k2[i] = f1[i] - k1[i];
    lu5_solve(W, piv, k2);
    for (int i = 0; i < 5; i++) k2[i] = k2[i] + k1[i];
    for (int i = 0; i < 5; i++) unew[i] = u[i] + dt * k2[i];
    f_eval(c, unew, t + dt, f2, nrhs);
    for (int i = 0; i < 5; i++) k3[i] = f2[i] - c32 * (k2[i] - f1[i]) - 2.0 * (k1[i] - f0[i]) + dt * dT[i];
    lu5_solve(W, piv, k3);
    double r[5];
    for (int i = 0; i < 5; i++) {
        double ut = dto6 * (k1[i] - 2.0 * k2[i] + k3[i]);
        double sc = fma(pm_max(fabs(u[i]), fabs(unew[i])), P->reltol, P->abstol);
        r[i] = ut / sc;
    }
    *EEst = rms5(r);
}

/* step!(integrator, DT, true): integrate one particle from t to t+DT */
static void integrate(oracle_t* o, particle_t* p, const rhs_ctx_t* c, double DT, picles_counters_t* C,
                      int64_t* stiff_triggers) {
    const picles_params_t* P = &o->P;
    const tableau_t* T = (P->solver == PICLES_SOLVER_DP5) ? &DP5 : &TSIT5;
    const int autosw = (P->solver == PICLES_SOLVER_AUTOTSIT5);
    if (p->status & (PICLES_PST_MAXITERS | PICLES_PST_DTMIN | PICLES_PST_UNSTABLE)) return; /* dead retcode */
    double t = p->t;
    double tstop = t + DT;
    double u[5], k[8][5];
    int64_t nrhs = 0;
    memcpy(u, p->u, sizeof u);
    int stiff = autosw && p->as_stiff;
    /* u_modified -> reset_fsal!: k1 = f(u,t) */
    f_eval(c, u, t, k[1], &nrhs);
    double dt = p->dt;
    if (p->dt_reset) {
        dt = initdt(c, u, t, k[1], &nrhs, stiff ? 2.0 : 5.0);
        p->dt_reset = 0;
    }
    double qold = p->qold;
    int64_t iter = p->iter;
    int attempts = 0;
    const double qmin = 0.2, qmax = 10.0, gamma = 0.9;
    while (t < tstop) {
        /* loopheader! */
        iter++;
        double dtmin_t = pm_max(pm_eps(t), P->dtmin);
        dt = pm_min(P->dtmax, dt);
        dt = pm_max(dt, dtmin_t);
        dt = pm_min(dt, tstop - t);
        /* check_error! */
        if (dt != dt) { p->status |= PICLES_PST_UNSTABLE; C->n_failed++; break; }
        if (iter > P->maxiters) { p->status |= PICLES_PST_MAXITERS; C->n_failed++; break; }
        /* DtLessThanMin; the final sliver onto the tstop is exempt (DiffEqBase check_error) */
        if (!P->force_dtmin && dt <= P->dtmin && (t + dt < tstop)) {
            p->status |= PICLES_PST_DTMIN; C->n_failed++; break;
        }
        attempts++;
        /* perform_step! */
        double tmp[5], un[5], g6[5], fnew[5];
        double EEst, eig = 0.0;
        /* PIController exponents of the current algorithm (reset_alg_dependent_opts!):
           beta2 = 2/(5 order), beta1 = 7/(10 order) */
        double beta1 = stiff ? 0.35 : T->beta1, beta2 = stiff ? 0.2 : T->beta2;
        if (stiff) {
            rosenbrock23_attempt(c, u, t, dt, k[1], un, fnew, &EEst, &eig, &nrhs);
            C->n_stiff_attempts++;
        } else {
            {
                double a = dt * T->a[2][1];
                for (int i = 0; i < 5; i++) tmp[i] = fma(a, k[1][i], u[i]);
                f_eval(c, tmp, fma(T->c[1], dt, t), k[2], &nrhs);
            }
            for (int s = 3; s <= 7; s++) {
                for (int i = 0; i < 5; i++) {
                    double inner = T->a[s][1] * k[1][i];
                    for (int j = 2; j < s; j++)
                        if (T->a[s][j] != 0.0) inner = fma(T->a[s][j], k[j][i], inner);
                    tmp[i] = fma(dt, inner, u[i]);
                }
                double ts = (s >= 6) ? (t + dt) : fma(T->c[s - 1], dt, t);
                f_eval(c, tmp, ts, k[s], &nrhs);
                if (s == 6) memcpy(g6, tmp, sizeof g6); /* argument of k6: g6 of the stiffness monitor */
                if (s == 7) memcpy(un, tmp, sizeof un);
            }
            memcpy(fnew, k[7], sizeof fnew);
            /* error estimate */
            double r[5];
            for (int i = 0; i < 5; i++) {
                double inner = T->bt[1] * k[1][i];
                for (int j = 2; j <= 7; j++)
                    if (T->bt[j] != 0.0) inner = fma(T->bt[j], k[j][i], inner);
                double ut = dt * inner;
                double sc = fma(pm_max(fabs(u[i]), fabs(un[i])), P->reltol, P->abstol);
                r[i] = ut / sc;
            }
            EEst = rms5(r);
            /* Tsit5 in a composite algorithm: eigen_est = max_i |k7-k6| / |g7-g6| (Hairer II, p. 22) */
            if (T == &TSIT5)
                for (int i = 0; i < 5; i++) eig = pm_max(eig, fabs((k[7][i] - k[6][i]) / (un[i] - g6[i])));
        }
        /* stepsize_controller! (PIController): q = EEst^beta1 / qold^beta2, evaluated as
           exp(beta1*log(EEst) - beta2*log(qold)); OrdinaryDiffEq evaluates the two powers
           with an approximate `fastpow`, so the last bits are not pinned by the reference */
        double q, q11 = 1.0;
        if (EEst == 0.0) {
            q = 1.0 / qmax;
        } else {
#if defined(ORACLE_CTRL_POW)
            /* sensitivity builds only (oracle/_build/libpicles_oracle_fastpow*.so, tests/test_sensitivity.py): older
               OrdinaryDiffEq versions form the two powers with `fastpow`, a Float32 approximation.  1: Float32
               pow (6e-8 relative); 2: its result cut to 12 mantissa bits (2e-4 relative, the accuracy of the
               fastlog2 polynomial behind fastpow).  Shows what an accept/reject or step-size perturbation of that
               size does to the fields; never the checker of the CUDA path. */
            q11 = ctrl_pow(EEst, beta1);
            q = q11 / ctrl_pow(qold, beta2);
            q = pm_max(1.0 / qmax, pm_min(1.0 / qmin, q / gamma));
#else
            double t1 = beta1 * O_LOG(EEst);
            q11 = O_EXP(t1);
            q = O_EXP(t1 - beta2 * O_LOG(qold));
            q = pm_max(1.0 / qmax, pm_min(1.0 / qmin, q / gamma));
#endif
        }
        int accept = (EEst <= 1.0) || (P->force_dtmin && fabs(dt) <= dtmin_t);
        if (accept) {
            /* step_accept_controller! (qsteady_min = qsteady_max = 1: the band is a no-op) */
            qold = pm_max(EEst, QOLDINIT);
            double dtnew = dt / q;
            double ttmp = t + dt;
            /* fixed_t_for_floatingpoint_error! */
            t = (fabs(ttmp - tstop) < 100.0 * pm_eps(tstop)) ? tstop : ttmp;
            /* calc_dt_propose! */
            double dtp = pm_min(P->dtmax, dtnew);
            dtp = pm_max(dtp, pm_max(pm_eps(t), P->dtmin));
            dt = dtp;
            memcpy(u, un, sizeof u);
            memcpy(k[1], fnew, sizeof k[1]); /* FSAL: k7 (Tsit5/DP5), f(u1, t+dt) (Rosenbrock23) */
            C->n_substeps++;
            int bad = 0;
            for (int i = 0; i < 5; i++) bad |= (u[i] != u[i]);
            if (bad) { p->status |= PICLES_PST_UNSTABLE; C->n_failed++; break; } /* unstable_check */
        } else {
            /* step_reject_controller! */
            /* q11 = NaN (EEst = NaN: the stages of the trial step overflowed): Julia's min propagates it and the integrator
               ends with DtNaN at the next check_error!, unless the powers are OrdinaryDiffEq's fastpow / fastpower, which
               turn the NaN bit pattern into a large finite number: rejected by 1/qmin (picles_params_t::nan_eest_rejects) */
            if (P->nan_eest_rejects && q11 != q11) dt = dt / (1.0 / qmin);
            else dt = dt / pm_min(1.0 / qmin, q11 / gamma);
            C->n_rejects++;
        }
        /* AutoSwitch (AutoTsit5 defaults: maxstiffstep 10, maxnonstiffstep 3, nonstifftol = stifftol
           = 9//10, dtfac 2, stability_size(Tsit5) = 3.5068), evaluated with the eigenvalue estimate
           of the attempt just made and the step size the controller proposes:
           solver Tsit5 only counts how often the switch would have happened */
        if (T == &TSIT5) {
            double stiffness = fabs(eig * dt / 3.5068);
            if (autosw) {
                int is = stiffness > 0.9;
                int32_t cnt = p->as_count;
                cnt = is ? ((cnt < 0) ? 1 : cnt + 1) : ((cnt > 0) ? -1 : cnt - 1);
                if (cnt > 60) cnt = 60;
                if (cnt < -60) cnt = -60;
                int switched = 0;
                if (!stiff && cnt > 10) {
                    dt = dt * 2.0; stiff = 1; switched = 1; C->n_stiff_switches++;
#ifdef _OPENMP
#pragma omp atomic
#endif
                    (*stiff_triggers)++;
                } else if (stiff && cnt < -3) {
                    dt = dt / 2.0; stiff = 0; switched = 1;
                }
                p->as_count = cnt;
                /* choose_algorithm!: initialize!(integrator, new cache) re-evaluates fsalfirst = f(uprev, t) */
                if (switched && t < tstop) f_eval(c, u, t, k[1], &nrhs);
            } else {
                if (stiffness > 0.9) p->stiff_run = (p->stiff_run < 0) ? 1 : p->stiff_run + 1;
                else p->stiff_run = (p->stiff_run > 0) ? -1 : p->stiff_run - 1;
                if (p->stiff_run > 10) {
#ifdef _OPENMP
#pragma omp atomic
#endif
                    (*stiff_triggers)++;
                    p->stiff_run = 0;
                }
            }
        }
    }
    memcpy(p->u, u, sizeof u);
    p->t = t;
    p->dt = dt;
    p->qold = qold;
    p->iter = iter;
    p->as_stiff = (uint8_t)stiff;
    C->n_rhs += nrhs;
    C->n_integrated++;
    if (attempts > C->max_attempts) C->max_attempts = attempts;
}

/* ------------------------------------------------------------------------ */
/* ParticleInCell.jl                                                          */
/* ------------------------------------------------------------------------ */
/* wrap_index!, ParticleInCell.jl:444-454 (1-based) */
static inline int64_t wrap_index(int64_t pos, int64_t N) {
    pos = pos % N;
    if (pos < 0) pos += N;
    else if (pos == 0) pos += N;
    return pos;
}
static inline int in_domain(int64_t pos, int64_t N) { return pos > 0 && pos <= N; } /* :464-466 */

/* push_to_grid!(grid, charge, index_pos, weights, Nx::AbstractBoundary, Ny::AbstractBoundary)
   ParticleInCell.jl:341-376; i, j 1-based */
static void push_corner(oracle_t* o, const double* ch, int64_t i, int64_t j, double wx, double wy) {
    int64_t Nx = o->Nx, Ny = o->Ny;
    if (((o->bx == PICLES_BND_NONPERIODIC) && !in_domain(i, Nx)) ||
        ((o->by == PICLES_BND_NONPERIODIC) && !in_domain(j, Ny)) ||
        ((o->by == PICLES_BND_TRIPOLAR_NORTH) && (j < 1)))
        return; /* :351-355 */
    int64_t ii, jj;
    if ((o->by == PICLES_BND_TRIPOLAR_NORTH) && (j > Ny)) {
        /* TripolarNorthBoundary, :409-428 (dispatches on Nx::N_Periodic only; any
           other Nx type raises a MethodError that push_to_grid! swallows, :360-365) */
        if (o->bx != PICLES_BND_PERIODIC) return;
        if (i < 0) ii = Nx - (Nx + i % Nx);
        else ii = Nx - i % Nx;
        jj = 2 * Ny - j + 1;
        if (ii < 1 || ii > Nx || jj < 1 || jj > Ny) return; /* BoundsError in the reference */
    } else {
        ii = wrap_index(i, Nx);
        jj = wrap_index(j, Ny);
    }
    int64_t idx = (ii - 1) + (jj - 1) * Nx;
    int64_t plane = Nx * Ny;
    double w = wx * wy;
    o->S[idx] += w * ch[0];
    o->S[idx + plane] += w * ch[1];
    o->S[idx + 2 * plane] += w * ch[2];
}

/* get_absolute_i_and_w(zp, i_node), ParticleInCell.jl:58-71 */
static inline int weights_1d(double zp, int64_t i_node, int64_t* idx, double* w) {
    if (!(fabs(zp) < 1.0e9)) return 0; /* Int(NaN/Inf) would throw in the reference */
    double base = floor(zp);
    int64_t f = (int64_t)base;
    double wc = rint((zp - base) * 1e6) / 1e6; /* round(x, digits=6) */
    double wf = 1.0 - wc;
    idx[0] = f + i_node; idx[1] = f + i_node + 1;
    w[0] = wf; w[1] = wc;
    return 1;
}

/* ParticleToNode!(PI, S, G::MeshGrids, periodic), mapping_2D.jl:59-73 */
static void particle_to_node(oracle_t* o, const particle_t* p, int64_t i1, int64_t j1, picles_counters_t* C) {
    int64_t xi[2], yi[2];
    double xw[2], yw[2];
    if (!weights_1d(p->u[3], i1, xi, xw) || !weights_1d(p->u[4], j1, yi, yw)) return;
    double ch[3];
    particle_to_charge(p->u, ch);
    /* construct_loop order, ParticleInCell.jl:504-508 */
    push_corner(o, ch, xi[0], yi[0], xw[0], yw[0]);
    push_corner(o, ch, xi[1], yi[0], xw[1], yw[0]);
    push_corner(o, ch, xi[0], yi[1], xw[0], yw[1]);
    push_corner(o, ch, xi[1], yi[1], xw[1], yw[1]);
    C->n_deposited++;
    int64_t r = 0, d;
    d = llabs(xi[0] - i1); if (d > r) r = d;
    d = llabs(xi[1] - i1); if (d > r) r = d;
    d = llabs(yi[0] - j1); if (d > r) r = d;
    d = llabs(yi[1] - j1); if (d > r) r = d;
    if (r > C->reach) C->reach = (int32_t)(r > 2147483647 ? 2147483647 : r);
}

/* ------------------------------------------------------------------------ */
/* mask_utils.jl                                                              */
/* ------------------------------------------------------------------------ */
/* make_boundaries(mask, Nx, Ny), mask_utils.jl:14-22,38-55; ocean: 1 ocean / 0 land */
void oracle_make_boundaries(const uint8_t* ocean, int Nx, int Ny, int bx, int by, uint8_t* total) {
    for (int j = 0; j < Ny; j++)
        for (int i = 0; i < Nx; i++) {
            int self = ocean[i + (int64_t)j * Nx] != 0;
            int b = 0;
            if (!self) { /* circshift(mask, dims) .&& .!mask over the 4 neighbours (periodic shift) */
                int im = (i - 1 + Nx) % Nx, ip = (i + 1) % Nx, jm = (j - 1 + Ny) % Ny, jp = (j + 1) % Ny;
                b = (ocean[im + (int64_t)j * Nx] != 0) | (ocean[ip + (int64_t)j * Nx] != 0) |
                    (ocean[i + (int64_t)jm * Nx] != 0) | (ocean[i + (int64_t)jp * Nx] != 0);
            }
            total[i + (int64_t)j * Nx] = (uint8_t)(self + 2 * b);
        }
    if (bx == PICLES_BND_NONPERIODIC)
        for (int j = 0; j < Ny; j++) { total[(int64_t)j * Nx] = 3; total[Nx - 1 + (int64_t)j * Nx] = 3; }
    if (by == PICLES_BND_NONPERIODIC)
        for (int i = 0; i < Nx; i++) { total[i] = 3; total[i + (int64_t)(Ny - 1) * Nx] = 3; }
}

/* ------------------------------------------------------------------------ */
/* model                                                                      */
/* ------------------------------------------------------------------------ */
oracle_t* oracle_create(int Nx, int Ny, int bx, int by, const uint8_t* mask, const double* M,
                        const double* M_const, const double* pc, const picles_params_t* P) {
    oracle_t* o = (oracle_t*)calloc(1, sizeof *o);
    int64_t n = (int64_t)Nx * Ny;
    o->Nx = Nx; o->Ny = Ny; o->bx = bx; o->by = by;
    o->mask = (uint8_t*)malloc(n);
    memcpy(o->mask, mask, n);
    if (M) { o->M = (double*)malloc(4 * n * sizeof(double)); memcpy(o->M, M, 4 * n * sizeof(double)); }
    if (M_const) memcpy(o->Mc, M_const, 4 * sizeof(double));
    if (pc) { o->pc = (double*)malloc(n * sizeof(double)); memcpy(o->pc, pc, n * sizeof(double)); }
    o->P = *P;
    o->part = (particle_t*)calloc(n, sizeof(particle_t));
    o->S = (double*)calloc(3 * n, sizeof(double));
    /* make_boundary_lists + ocean_points, mask_utils.jl:71-82, WaveGrowthModels2D.jl:256-270:
       findall(mask .== 1) in column-major order, then (periodic_boundary only) findall(mask .== 3) */
    o->ocean = (int64_t*)malloc(n * sizeof(int64_t));
    int64_t m = 0;
    for (int64_t l = 0; l < n; l++) if (mask[l] == 1) o->ocean[m++] = l;
    if (P->periodic_boundary)
        for (int64_t l = 0; l < n; l++) if (mask[l] == 3) o->ocean[m++] = l;
    o->n_ocean = m;
    o->nthreads = 1;
    return o;
}

void oracle_destroy(oracle_t* o) {
    if (!o) return;
    free(o->mask); free(o->M); free(o->pc); free(o->part); free(o->ocean); free(o->S);
    free(o->u_mid); free(o->v_mid); free(o->x); free(o->y);
    free(o->u_lag); free(o->v_lag);
    free(o);
}

void oracle_set_threads(oracle_t* o, int n) { o->nthreads = n < 1 ? 1 : n; }
void oracle_set_accumulate(oracle_t* o, int on) { o->accumulate = on ? 1 : 0; }
/* intermediate wind levels at t + k*DT/(n_mid+1), k = 1..n_mid, for the next oracle_step only */
void oracle_set_wind_midlevels(oracle_t* o, int n_mid, const double* u_mid, const double* v_mid) {
    size_t n = (size_t)o->Nx * o->Ny;
    free(o->u_mid); free(o->v_mid);
    o->u_mid = o->v_mid = NULL;
    o->n_mid = n_mid;
    if (n_mid > 0) {
        o->u_mid = (double*)malloc(n_mid * n * sizeof(double));
        o->v_mid = (double*)malloc(n_mid * n * sizeof(double));
        memcpy(o->u_mid, u_mid, n_mid * n * sizeof(double));
        memcpy(o->v_mid, v_mid, n_mid * n * sizeof(double));
    }
}
/* reference semantics for the wind: call the closures at the home node and the stage time.
   fn == NULL returns to staged levels. */
void oracle_set_wind_closure(oracle_t* o, void (*fn)(double, double, double, double*, double*), const double* x,
                             const double* y) {
    size_t n = (size_t)o->Nx * o->Ny;
    free(o->x); free(o->y);
    o->x = o->y = NULL;
    o->wind_fn = fn;
    if (fn) {
        o->x = (double*)malloc(n * sizeof(double)); memcpy(o->x, x, n * sizeof(double));
        o->y = (double*)malloc(n * sizeof(double)); memcpy(o->y, y, n * sizeof(double));
    }
}

/* init_particles! / SeedParticle, run.jl:199-247, core_2D.jl:434-488 */
void oracle_seed(oracle_t* o, const double* u0, const double* v0) {
    const picles_params_t* P = &o->P;
    int64_t n = (int64_t)o->Nx * o->Ny;
    memset(o->S, 0, 3 * n * sizeof(double));
    memset(o->part, 0, n * sizeof(particle_t));
    for (int64_t l = 0; l < n; l++) {
        particle_t* p = &o->part[l];
        if (o->mask[l] == 0) continue; /* dummy instance, core_2D.jl:455-458 */
        p->exists = 1;
        /* InitParticleValues, core_2D.jl:247-288 */
        if (!P->has_defaults) {
            double wu = u0[l], wv = v0[l];
            if (sqrt(wu * wu + wv * wv) > sqrt(2.0)) {
                windsea(wu, wv, P->seed_timescale, &p->u[0], &p->u[1], &p->u[2], NULL, NULL);
                p->on = 1;
            } else {
                minimal_windsea(wu, wv, P->seed_timescale, &p->u[0], &p->u[1], &p->u[2], NULL, NULL, NULL, NULL);
                p->on = 0;
            }
        } else {
            p->u[0] = P->defaults[0]; p->u[1] = P->defaults[1]; p->u[2] = P->defaults[2];
            p->on = 1;
        }
        p->u[3] = 0.0; p->u[4] = 0.0;
        /* check_boundary_point, core_2D.jl:360-366 */
        p->boundary = P->periodic_boundary ? (o->mask[l] == 2) : (o->mask[l] >= 2);
        if (p->on) { /* init_z0_to_State!, initialize.jl:14-17 */
            double ch[3];
            particle_to_charge(p->u, ch);
            o->S[l] = ch[0]; o->S[l + n] = ch[1]; o->S[l + 2 * n] = ch[2];
        }
        p->t = 0.0; p->dt = P->dt; p->qold = QOLDINIT; p->iter = 0;
    }
    for (int64_t m = 0; m < o->n_ocean; m++) o->part[o->ocean[m]].active = 1;
    o->steps_since_seed = 0;
    free(o->u_lag); free(o->v_lag);
    o->u_lag = o->v_lag = NULL;
}

static inline void node_M(const oracle_t* o, int64_t l, double* M) {
    int64_t n = (int64_t)o->Nx * o->Ny;
    if (o->M) { M[0] = o->M[l]; M[1] = o->M[l + n]; M[2] = o->M[l + 2 * n]; M[3] = o->M[l + 3 * n]; }
    else memcpy(M, o->Mc, 4 * sizeof(double));
}

/* advance! without the deposit, mapping_2D.jl:118-235; returns the local `on` */
static int advance_particle(oracle_t* o, int64_t l, double DT, const double* u_t, const double* v_t,
                            const double* u_t1, const double* v_t1, picles_counters_t* C) {
    const picles_params_t* P = &o->P;
    particle_t* p = &o->part[l];
    double t_start = p->t; /* :132 */
    int on = p->on;
    /* winds.u/v(x, y, t_start) and (x, y, t_start + DT) of THIS particle (:173-176, :201-204, :214-216).
       The integrator clock of a particle that integrates every step is the model clock: the staged
       t / t+DT levels.  One that never integrates (seeded off, `on` frozen: B-1 as run) keeps
       t_start = 0, so its t_end is DT for ever: the level kept from the first step.  (Its t_start
       level is only read by the Inf fix-up, which a reseed from a finite wind cannot reach.) */
    double wu_start = u_t[l], wv_start = v_t[l], wu_end = u_t1[l], wv_end = v_t1[l];
    if (o->wind_fn) {
        o->wind_fn(o->x[l], o->y[l], t_start, &wu_start, &wv_start);
        o->wind_fn(o->x[l], o->y[l], t_start + DT, &wu_end, &wv_end);
    } else if (!on && !P->on_persist && o->u_lag) {
        wu_end = o->u_lag[l]; wv_end = o->v_lag[l];
    }
    if (on) { /* :149-170 */
        double M[4];
        node_M(o, l, M);
        rhs_ctx_t c;
        c.o = o; c.M = M; c.pc = o->pc ? o->pc[l] : 0.0;
        {
            int64_t n = (int64_t)o->Nx * o->Ny;
            double wu[WIND_SEG_MAX + 1], wv[WIND_SEG_MAX + 1];
            c.nseg = o->n_mid + 1;
            wu[0] = u_t[l]; wv[0] = v_t[l];
            for (int k = 1; k <= o->n_mid; k++) { wu[k] = o->u_mid[(k - 1) * n + l]; wv[k] = o->v_mid[(k - 1) * n + l]; }
            wu[c.nseg] = u_t1[l]; wv[c.nseg] = v_t1[l];
            wind_coefficients(wu, c.nseg, c.cu);
            wind_coefficients(wv, c.nseg, c.cv);
        }
        c.t_start = t_start; c.inv_DT = 1.0 / DT;
        c.x = o->x ? o->x[l] : 0.0; c.y = o->y ? o->y[l] : 0.0;
        integrate(o, p, &c, DT, C, &o->stiff_triggers);
    } else { /* :172-185: wind at t_end = t_start + DT, t_start = the particle's own integrator clock */
        if (wu_end * wu_end + wv_end * wv_end >= P->wind_min_squared) {
            reset_particle_values(P, wu_end, wv_end, DT, p->u);
            p->dt_reset = 1; /* reset_PI_u!, :91-96 */
            on = 1;
            C->n_reseed_advance++;
        }
    }
    /* fix-ups, :196-235 */
    int anynan = (p->u[0] != p->u[0]) | (p->u[1] != p->u[1]) | (p->u[2] != p->u[2]);
    int anyinf = pm_isinf(p->u[0]) | pm_isinf(p->u[1]) | pm_isinf(p->u[2]);
    if (anynan) {
        reset_particle_values(P, wu_end, wv_end, DT, p->u); /* wind at t_end */
        p->dt_reset = 1; p->status |= PICLES_PST_NAN_RESET; C->n_fixups++;
    } else if (anyinf) {
        reset_particle_values(P, wu_start, wv_start, DT, p->u); /* wind at t_start */
        p->dt_reset = 1; p->status |= PICLES_PST_INF_RESET; C->n_fixups++;
    } else if (p->u[0] > P->log_energy_maximum) {
        p->u[0] = P->log_energy_maximum;
        p->dt_reset = 1; p->status |= PICLES_PST_EMAX_CLAMP; C->n_fixups++;
    }
    if (P->on_persist) p->on = (uint8_t)on;
    return on;
}

/* remesh! / NodeToParticle!, mapping_2D.jl:250-356 */
static void remesh_particle(oracle_t* o, int64_t l, double DT, const double* u_t, const double* v_t,
                            picles_counters_t* C) {
    const picles_params_t* P = &o->P;
    particle_t* p = &o->part[l];
    int64_t n = (int64_t)o->Nx * o->Ny;
    double s[3] = {o->S[l], o->S[l + n], o->S[l + 2 * n]}; /* Get_u_FromShared, core_2D.jl:132 */
    double wu = u_t[l], wv = v_t[l];                       /* wind at the pre-tick clock, :258 */
    int on = p->on;
    if (!p->boundary && (s[0] >= P->minimal_state[0]) &&
        (s[1] * s[1] + s[2] * s[2] >= P->minimal_state[1])) { /* :306-312 */
        vertex_to_particle(s, 0.0, 0.0, p->u);
        p->dt_reset = 1; /* reset_PI_ut!: t unchanged, qold/iter kept */
        on = 1;
        C->n_remesh_A++;
    } else if (!p->boundary && (wu * wu + wv * wv >= P->wind_min_squared)) { /* :328-336 */
        reset_particle_values(P, wu, wv, DT, p->u);
        p->qold = QOLDINIT; p->iter = 0; p->status = 0; /* reinit! */
        p->as_count = 0; p->as_stiff = 0;                /* ... with a fresh AutoSwitch state */
        p->dt_reset = 1;
        on = 1;
        C->n_remesh_B++;
    } else if (p->boundary && (wu * wu + wv * wv >= P->wind_min_squared)) { /* :338-344 */
        reset_particle_values(P, wu, wv, DT, p->u);
        p->qold = QOLDINIT; p->iter = 0; p->status = 0;
        p->as_count = 0; p->as_stiff = 0;
        p->dt_reset = 1;
        on = 1;
        C->n_remesh_C++;
    } else { /* :347-353 */
        on = 0;
        C->n_remesh_D++;
    }
    if (P->on_persist) p->on = (uint8_t)on;
}

/* State .= 0 ; time_step!  (run.jl:75-82, TimeSteppers.jl:109-166) */
void oracle_step(oracle_t* o, double t, double DT, const double* u_t, const double* v_t,
                 const double* u_t1, const double* v_t1) {
    (void)t;
    int64_t n = (int64_t)o->Nx * o->Ny;
    if (!o->accumulate) memset(o->S, 0, 3 * n * sizeof(double)); /* run.jl:75-79 */
    if (!o->P.on_persist && o->steps_since_seed == 0 && !o->wind_fn) {
        /* the level at integrator time 0 + DT, read on every later step by the particles seeded off */
        free(o->u_lag); free(o->v_lag);
        o->u_lag = (double*)malloc(n * sizeof(double)); memcpy(o->u_lag, u_t1, n * sizeof(double));
        o->v_lag = (double*)malloc(n * sizeof(double)); memcpy(o->v_lag, v_t1, n * sizeof(double));
    }
    o->steps_since_seed++;
    picles_counters_t C;
    memset(&C, 0, sizeof C);
    C.n_active = o->n_ocean;
    uint8_t* on_local = (uint8_t*)malloc((size_t)(o->n_ocean > 0 ? o->n_ocean : 1));
    /* advance: the ODE phase is independent per particle (threads allowed); the
       deposit below stays serial in ocean_points order = movie_time_step! order */
#ifdef _OPENMP
    if (o->nthreads > 1) {
#pragma omp parallel num_threads(o->nthreads)
        {
            picles_counters_t Cl;
            memset(&Cl, 0, sizeof Cl);
#pragma omp for schedule(dynamic, 256)
            for (int64_t m = 0; m < o->n_ocean; m++)
                on_local[m] = (uint8_t)advance_particle(o, o->ocean[m], DT, u_t, v_t, u_t1, v_t1, &Cl);
#pragma omp critical
            {
                C.n_integrated += Cl.n_integrated; C.n_substeps += Cl.n_substeps; C.n_rejects += Cl.n_rejects;
                C.n_rhs += Cl.n_rhs; C.n_reseed_advance += Cl.n_reseed_advance; C.n_fixups += Cl.n_fixups;
                C.n_failed += Cl.n_failed;
                C.n_stiff_switches += Cl.n_stiff_switches; C.n_stiff_attempts += Cl.n_stiff_attempts;
                if (Cl.max_attempts > C.max_attempts) C.max_attempts = Cl.max_attempts;
            }
        }
    } else
#endif
    {
        for (int64_t m = 0; m < o->n_ocean; m++)
            on_local[m] = (uint8_t)advance_particle(o, o->ocean[m], DT, u_t, v_t, u_t1, v_t1, &C);
    }
    for (int64_t m = 0; m < o->n_ocean; m++) {
        if (!on_local[m]) continue; /* :238-240 */
        int64_t l = o->ocean[m];
        particle_to_node(o, &o->part[l], l % o->Nx + 1, l / o->Nx + 1, &C);
    }
    free(on_local);
    for (int64_t m = 0; m < o->n_ocean; m++) remesh_particle(o, o->ocean[m], DT, u_t, v_t, &C);
    o->C = C;
    o->n_mid = 0; /* intermediate levels are consumed by one step */
}

void oracle_get_state(const oracle_t* o, double* S) {
    memcpy(S, o->S, 3 * (size_t)o->Nx * o->Ny * sizeof(double));
}
void oracle_set_state(oracle_t* o, const double* S) {
    memcpy(o->S, S, 3 * (size_t)o->Nx * o->Ny * sizeof(double));
}
void oracle_get_particles(const oracle_t* o, double* z, double* t, double* dt, uint8_t* flags, int32_t* status) {
    int64_t n = (int64_t)o->Nx * o->Ny;
    for (int64_t l = 0; l < n; l++) {
        const particle_t* p = &o->part[l];
        if (z) for (int k = 0; k < 5; k++) z[l + k * n] = p->u[k];
        if (t) t[l] = p->t;
        if (dt) dt[l] = p->dt;
        if (flags) flags[l] = (uint8_t)((p->on ? PICLES_PF_ON : 0) | (p->boundary ? PICLES_PF_BOUNDARY : 0) |
                                        (p->dt_reset ? PICLES_PF_DT_RESET : 0) | (p->active ? PICLES_PF_ACTIVE : 0));
        if (status) status[l] = p->status;
    }
}
void oracle_get_aux(const oracle_t* o, double* qold, int64_t* iter) {
    int64_t n = (int64_t)o->Nx * o->Ny;
    for (int64_t l = 0; l < n; l++) { if (qold) qold[l] = o->part[l].qold; if (iter) iter[l] = o->part[l].iter; }
}
void oracle_get_counters(const oracle_t* o, picles_counters_t* c) { *c = o->C; }
/* AutoSwitch state as picles_get_solver_state packs it */
void oracle_get_solver_state(const oracle_t* o, int8_t* as) {
    int64_t n = (int64_t)o->Nx * o->Ny;
    for (int64_t l = 0; l < n; l++) as[l] = (int8_t)(o->part[l].as_count + (o->part[l].as_stiff ? 64 : 0));
}
int64_t oracle_n_ocean(const oracle_t* o) { return o->n_ocean; }
/* derived output fields: Hs = 4*sqrt(e) (movie_2D.jl:50), GetGroupVelocity (core_2D.jl:138-147) */
void oracle_fields(const oracle_t* o, double* Hs, double* cx, double* cy) {
    int64_t n = (int64_t)o->Nx * o->Ny;
    for (int64_t l = 0; l < n; l++) {
        double e = o->S[l], mx = o->S[n + l], my = o->S[2 * n + l];
        double m_amp = sqrt(mx * mx + my * my);
        Hs[l] = 4.0 * sqrt(e);
        cx[l] = mx * e / (2.0 * (m_amp * m_amp));
        cy[l] = my * e / (2.0 * (m_amp * m_amp));
    }
}
int64_t oracle_stiff_triggers(const oracle_t* o) { return o->stiff_triggers; }
void oracle_get_ocean_points(const oracle_t* o, int64_t* idx) { memcpy(idx, o->ocean, o->n_ocean * sizeof(int64_t)); }

/* ------------------------------------------------------------------------ */
/* gridded winds: Interpolations.LinearInterpolation((x,y,t), U, extrapolation_bc=Periodic())   */
/* tests/T03_PIC_tripolar_realistic.jl:61-73, src/Utils/WindEmulator.jl:18-43.  Interpolations.jl  */
/* is a third-party dependency absent from /root/reference (Project.toml, no [compat] bound); its  */
/* published rule is restated: periodic(y,l,u) = mod(y-l, u-l) + l on every axis, knot interval  */
/* i = clamp(searchsortedfirst(knots, y) - 1, 1, n-1), weights (1-d, d) with d = (y-k_i)/(k_i+1-k_i), */
/* value = nested weighted sum, first axis outermost.                                             */
/* ------------------------------------------------------------------------ */
static double julia_mod(double x, double y) { /* Base.mod(::Float64, ::Float64) */
    double r = fmod(x, y);
    if (r == 0.0) return copysign(r, y);
    if ((r > 0.0) != (y > 0.0)) return r + y;
    return r;
}
static void knot_interval(const double* k, int n, double y, int* i, double* d) {
    int first = 0; /* searchsortedfirst: number of knots < y */
    while (first < n && k[first] < y) first++;
    int idx = first - 1; /* 0-based */
    if (idx < 0) idx = 0;
    if (idx > n - 2) idx = n - 2;
    *i = idx;
    *d = (y - k[idx]) / (k[idx + 1] - k[idx]);
}
static double trilinear(const double* A, int nx, int ny, int ix, int iy, int it, double dx, double dy, double dt) {
    double acc = 0.0;
    const double wx[2] = {1.0 - dx, dx}, wy[2] = {1.0 - dy, dy}, wt[2] = {1.0 - dt, dt};
    double sx[2];
    for (int a = 0; a < 2; a++) {
        double sy[2];
        for (int b = 0; b < 2; b++) {
            const double* p = A + (ix + a) + (int64_t)nx * (iy + b) + (int64_t)nx * ny * it;
            sy[b] = wt[0] * p[0] + wt[1] * p[(int64_t)nx * ny];
        }
        sx[a] = wy[0] * sy[0] + wy[1] * sy[1];
    }
    acc = wx[0] * sx[0] + wx[1] * sx[1];
    return acc;
}
void oracle_wind_mesh_sample(int nx, int ny, int nt, const double* xw, const double* yw, const double* tw,
                             const double* U, const double* V, int64_t n, const double* x, const double* y, double t,
                             double* u_out, double* v_out) {
    int it;
    double dt;
    knot_interval(tw, nt, julia_mod(t - tw[0], tw[nt - 1] - tw[0]) + tw[0], &it, &dt);
    for (int64_t l = 0; l < n; l++) {
        int ix, iy;
        double dx, dy;
        knot_interval(xw, nx, julia_mod(x[l] - xw[0], xw[nx - 1] - xw[0]) + xw[0], &ix, &dx);
        knot_interval(yw, ny, julia_mod(y[l] - yw[0], yw[ny - 1] - yw[0]) + yw[0], &iy, &dy);
        u_out[l] = trilinear(U, nx, ny, ix, iy, it, dx, dy, dt);
        v_out[l] = trilinear(V, nx, ny, ix, iy, it, dx, dy, dt);
    }
}

/* ------------------------------------------------------------------------ */
/* unit hooks for tests                                                       */
/* ------------------------------------------------------------------------ */
void oracle_rhs(const picles_params_t* P, const double* z, double u, double v, const double* M, double pc, double* dz) {
    rhs(P, z, u, v, M, pc, dz);
}
void oracle_windsea(double u, double v, double T, double* out5, double* E, double* cg_amp) {
    windsea(u, v, T, &out5[0], &out5[1], &out5[2], E, cg_amp);
    out5[3] = 0.0; out5[4] = 0.0;
}
/* MinimalState(U,V,T) = [E, m_x^2+m_y^2], FetchRelations.jl:412-415 (mom = (U/Uamp)*E/(2*cg), :353-354) */
void oracle_minimal_state(double U, double V, double T, double* out2, double* part5) {
    double lne, cgx, cgy, E, cg, ux, uy;
    minimal_windsea(U, V, T, &lne, &cgx, &cgy, &E, &cg, &ux, &uy);
    double Ua = sqrt(ux * ux + uy * uy);
    Ua = (Ua < 0.1) ? 0.1 : Ua;
    double mx = (ux / Ua) * E / (2.0 * cg), my = (uy / Ua) * E / (2.0 * cg);
    out2[0] = E; out2[1] = mx * mx + my * my;
    if (part5) { part5[0] = lne; part5[1] = cgx; part5[2] = cgy; part5[3] = 0.0; part5[4] = 0.0; }
}
void oracle_particle_to_charge(const double* u, double* ch) { particle_to_charge(u, ch); }
void oracle_vertex_to_particle(const double* s, double* u) { vertex_to_particle(s, 0.0, 0.0, u); }
void oracle_weights(double zp, int64_t i_node, int64_t* idx, double* w) {
    if (!weights_1d(zp, i_node, idx, w)) { idx[0] = idx[1] = 0; w[0] = w[1] = 0.0; }
}
int64_t oracle_wrap_index(int64_t pos, int64_t N) { return wrap_index(pos, N); }
/* maps one deposit corner (1-based i,j) to its linear 0-based target or -1 (dropped) */
int64_t oracle_corner_target(int Nx, int Ny, int bx, int by, int64_t i, int64_t j) {
    if (((bx == PICLES_BND_NONPERIODIC) && !in_domain(i, Nx)) || ((by == PICLES_BND_NONPERIODIC) && !in_domain(j, Ny)) ||
        ((by == PICLES_BND_TRIPOLAR_NORTH) && (j < 1)))
        return -1;
    int64_t ii, jj;
    if ((by == PICLES_BND_TRIPOLAR_NORTH) && (j > Ny)) {
        if (bx != PICLES_BND_PERIODIC) return -1;
        if (i < 0) ii = Nx - (Nx + i % Nx); else ii = Nx - i % Nx;
        jj = 2 * (int64_t)Ny - j + 1;
        if (ii < 1 || ii > Nx || jj < 1 || jj > Ny) return -1;
    } else { ii = wrap_index(i, Nx); jj = wrap_index(j, Ny); }
    return (ii - 1) + (jj - 1) * (int64_t)Nx;
}
/* single-particle driver: integrate u over DT with the wind linear between two levels;
   as2 = {AutoSwitch count, is_stiffalg} in/out (NULL: a fresh non-stiff integrator) */
void oracle_integrate_one_as(const picles_params_t* P, const double* M, double pc, double* u5, double* t,
                             double* dt, double* qold, int64_t* iter, int dt_reset, double wu0, double wv0,
                             double wu1, double wv1, double DT, picles_counters_t* C_out, int32_t* status,
                             int32_t* as2) {
    oracle_t o;
    memset(&o, 0, sizeof o);
    o.P = *P;
    particle_t p;
    memset(&p, 0, sizeof p);
    memcpy(p.u, u5, sizeof p.u);
    p.t = *t; p.dt = *dt; p.qold = *qold; p.iter = *iter; p.dt_reset = (uint8_t)dt_reset; p.on = 1;
    p.status = *status;
    if (as2) { p.as_count = as2[0]; p.as_stiff = (uint8_t)(as2[1] != 0); }
    rhs_ctx_t c;
    c.o = &o; c.M = M; c.pc = pc;
    {
        double wu[2] = {wu0, wu1}, wv[2] = {wv0, wv1};
        c.nseg = 1;
        wind_coefficients(wu, 1, c.cu);
        wind_coefficients(wv, 1, c.cv);
    }
    c.t_start = p.t; c.inv_DT = 1.0 / DT; c.x = 0.0; c.y = 0.0;
    picles_counters_t C;
    memset(&C, 0, sizeof C);
    integrate(&o, &p, &c, DT, &C, &o.stiff_triggers);
    memcpy(u5, p.u, sizeof p.u);
    *t = p.t; *dt = p.dt; *qold = p.qold; *iter = p.iter; *status = p.status;
    if (as2) { as2[0] = p.as_count; as2[1] = p.as_stiff; }
    if (C_out) *C_out = C;
}
void oracle_integrate_one(const picles_params_t* P, const double* M, double pc, double* u5, double* t,
                          double* dt, double* qold, int64_t* iter, int dt_reset, double wu0, double wv0,
                          double wu1, double wv1, double DT, picles_counters_t* C_out, int32_t* status) {
    oracle_integrate_one_as(P, M, pc, u5, t, dt, qold, iter, dt_reset, wu0, wv0, wu1, wv1, DT, C_out, status, NULL);
}
/* Jacobian (5x5, row-major) and time gradient of the right-hand side by dual numbers */
void oracle_rhs_jacobian(const picles_params_t* P, const double* z, double u, double v, double ut, double vt,
                         const double* M, double pc, double* J25, double* dT5) {
    pmd_t dz[5];
    rhs_dual(P, pmd_var(z[0], 0), pmd_var(z[1], 1), pmd_var(z[2], 2), pmd_var(u, 3), pmd_var(v, 4), M, pc, dz);
    for (int i = 0; i < 5; i++) {
        for (int j = 0; j < 5; j++) J25[i * 5 + j] = (j < 3) ? dz[i].d[j] : 0.0;
        dT5[i] = dz[i].d[3] * ut + dz[i].d[4] * vt;
    }
}

#define VEC_HOOK(name, expr)                                         \
    void name(int64_t n, const double* x, double* out) {             \
        for (int64_t i = 0; i < n; i++) { double a = x[i]; out[i] = (expr); } \
    }
VEC_HOOK(oracle_pm_exp, pm_exp(a))
VEC_HOOK(oracle_pm_log, pm_log(a))
VEC_HOOK(oracle_pm_tanh, pm_tanh(a))
VEC_HOOK(oracle_pm_sech, pm_sech(a))
VEC_HOOK(oracle_pm_cosh, pm_cosh(a))
VEC_HOOK(oracle_pm_eps, pm_eps(a))
static double hook_sin(double a) { double s, c; pm_sincos(a, &s, &c); return s; }
static double hook_cos(double a) { double s, c; pm_sincos(a, &s, &c); return c; }
static double hook_sind(double a) { double s, c; pm_sincosd(a, &s, &c); return s; }
static double hook_cosd(double a) { double s, c; pm_sincosd(a, &s, &c); return c; }
VEC_HOOK(oracle_pm_sin, hook_sin(a))
VEC_HOOK(oracle_pm_cos, hook_cos(a))
VEC_HOOK(oracle_pm_sind, hook_sind(a))
VEC_HOOK(oracle_pm_cosd, hook_cosd(a))
VEC_HOOK(oracle_pm_tand, pm_tand(a))
/* grid metric of n nodes: ProjetionKernel(Gi, stats) (TripolarGridMOM6.jl:448-459) and
   SphericalPropagationCorrection(ij_mesh, stats) (spherical_grid_corrections.jl:13,49-51);
   M: 4 planes (M11, M12, M21, M22) of n */
void oracle_grid_metric(int64_t n, const double* dx, const double* dy, const double* angle_dx, const double* lat,
                        double R_earth, double* M, double* pc) {
    for (int64_t l = 0; l < n; l++)
        pm_grid_metric_node(dx[l], dy[l], angle_dx[l], lat[l], R_earth, &M[l], &M[n + l], &M[2 * n + l], &M[3 * n + l], &pc[l]);
}
void oracle_pm_pow(int64_t n, const double* x, const double* y, double* out) {
    for (int64_t i = 0; i < n; i++) out[i] = pm_pow(x[i], y[i]);
}
int oracle_uses_libm(void) {
#ifdef ORACLE_LIBM
    return 1;
#else
    return 0;
#endif
}
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
