/*
 * stiff.h — the stiff branch of ODESettings' default solver AutoTsit5(Rosenbrock23())
 * (src/ParticleSystems/particle_waves_v5.jl:47): OrdinaryDiffEq's AutoSwitch hands a particle to
 * Rosenbrock23 when the Tsit5 stiffness estimate stays above its threshold for more than ten
 * consecutive attempts.  This happens for a handful of particles at the calm foot of a wind ramp
 * and never on the homogeneous-box or tripolar configurations, so everything here is COLD code:
 * one out-of-line function per kernel, local arrays, IEEE operators — the hot loop of physics.h
 * only carries the monitor (one eigenvalue estimate per attempt when the solver id is
 * PICLES_SOLVER_AUTOTSIT5).
 *
 * OrdinaryDiffEq is a third-party dependency absent from /root/reference (Project.toml:16, no
 * [compat] bound): its published algorithm is restated — Rosenbrock23 = Shampine's ode23s
 * (d = 1/(2+sqrt 2), c32 = 6+sqrt 2) with Jacobian and time gradient by automatic differentiation
 * (pmath_dual.h), LU with partial pivoting, eigen_est = opnorm(J, Inf) for the switch back; the
 * PI controller exponents follow the current algorithm's order (beta2 = 2/(5*2), beta1 = 7/(10*2)).
 * The CPU checker carries its own restatement of the same rules; both agree bit for bit.
 */
#ifndef PICLES_STIFF_H
#define PICLES_STIFF_H

#include "physics.h"
#include "pmath_dual.h"

namespace picles {

PM_HD_NOINLINE_DECL void wind_poly_eval(const WindPoly* W, double ts, double* u, double* v, double* ut, double* vt) {
    double sg = (ts - W->t0) * W->scale;
    double pu = W->cu[W->nseg], pv = W->cv[W->nseg], du = 0.0, dv = 0.0;
    for (int m = W->nseg - 1; m >= 0; m--) {
        double a = sg - (double)m;
        du = fma(du, a, pu); dv = fma(dv, a, pv);
        pu = fma(pu, a, W->cu[m]); pv = fma(pv, a, W->cv[m]);
    }
    *u = pu; *v = pv;
    *ut = du * W->scale; *vt = dv * W->scale;
}

/* IfElse.ifelse(a > 500, 500, a), particle_waves_v5.jl:215-225, on dual numbers */
PM_HD pmd_t alpha_func_dual(pmd_t us, pmd_t cgp) {
    pmd_t a = pmd_div(us, pmd_scale(2.0, cgp));
    return (a.v > 500.0) ? pmd_const(500.0) : a;
}

/* particle_system(dz, z, params, t) on dual numbers: the operation sequence of rhs3/prop with the
   partials of (lne, c̄_x, c̄_y, wind u, wind v) carried along — what ForwardDiff evaluates for
   Rosenbrock23 (particle_waves_v5.jl:479-556) */
PM_HD_NOINLINE_DECL void rhs_dual(const picles_params_t* Pp, const double* z, double wu, double wv, const double* M,
                                  double pc, pmd_t* dz) {
    const picles_params_t& P = *Pp;
    pmd_t lne = pmd_var(z[0], 0), cx = pmd_var(z[1], 1), cy = pmd_var(z[2], 2), u = pmd_var(wu, 3), v = pmd_var(wv, 4);
    pmd_t r_g = pmd_const(P.r_g);
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
    pmd_t Hp = pmd_scale(0.5, pmd_addc(pmd_tanh(pmd_scale(P.p, pmd_addc(alpha_p, -0.85))), 1.0));
    pmd_t sch = pmd_sech(pmd_scale(10.0, pmd_addc(alpha_p, -0.85)));
    pmd_t Dp = pmd_addc(pmd_scale(-1.25, pmd_sqr(sch)), 1.0);
    pmd_t It = pmd_const(0.0), Dt = pmd_const(0.0), Scg = pmd_const(0.0), Sdir = pmd_const(0.0);
    if (P.input) It = pmd_mul(pmd_scale(P.C_e, Hp), pmd_sqr(alpha));
    if (P.dissipation) {
        pmd_t r = pmd_div(kp, pmd_const(P.e_T)), pw;
        double twon = 2.0 * P.n;
        if (twon == 4.0) pw = pmd_sqr(pmd_sqr(r));
        else if (twon == 2.0) pw = pmd_sqr(r);
        else pw = pmd_pow_given(r, twon, pm_pow_safe(r.v, twon), pm_pow_safe(r.v, twon - 1.0));
        Dt = pmd_mul(pmd_exp(pmd_scale(P.n, lne)), pw);
    }
    if (P.peak_shift) Scg = pmd_mul(pmd_mul(pmd_scale(P.C_alpha, Dp), pmd_sqr(pmd_sqr(kp))), pmd_exp(pmd_scale(2.0, lne)));
    if (P.direction) {
        pmd_t a2 = alpha_func_dual(us, sg);
        pmd_t prod = pmd_mul(us, sg);
        pmd_t s2 = pmd_const(0.0);
        if (!(prod.v == 0.0)) {
            pmd_t t1 = pmd_mul(pmd_mul(u, v), pmd_sub(pmd_scale(2.0, pmd_sqr(gy)), pmd_sqr(sg)));
            pmd_t t2 = pmd_mul(pmd_mul(gx, gy), pmd_sub(pmd_scale(2.0, pmd_sqr(v)), pmd_sqr(us)));
            s2 = pmd_mul(pmd_cdiv(2.0, pmd_sqr(prod)), pmd_sub(t1, t2));
        }
        Sdir = pmd_mul(pmd_mul(pmd_scale(P.C_varphi, pmd_sqr(a2)), Hp), s2);
    }
    pmd_t Ssph = pmd_scale(pc, cx);
    pmd_t wrs = pmd_mul(pmd_mul(wp, r_g), Scg);
    dz[0] = pmd_add(wrs, pmd_mul(wp, pmd_sub(It, Dt)));
    dz[1] = pmd_add(pmd_add(pmd_neg(pmd_mul(cx, wrs)), pmd_mul(cy, Sdir)), pmd_mul(cy, Ssph));
    dz[2] = pmd_sub(pmd_sub(pmd_neg(pmd_mul(cy, wrs)), pmd_mul(cx, Sdir)), pmd_mul(cx, Ssph));
    if (P.propagation) {
        dz[3] = pmd_add(pmd_scale(M[0], cx), pmd_scale(M[1], cy));
        dz[4] = pmd_add(pmd_scale(M[2], cx), pmd_scale(M[3], cy));
    } else {
        dz[3] = pmd_const(0.0); dz[4] = pmd_const(0.0);
    }
}

/* f(u, t): all five components with the IEEE right-hand side */
PM_HD_NOINLINE_DECL void f5_cold(const picles_params_t* Pp, const WindPoly* W, const double* M, double pc, const double* z,
                                 double ts, double* dz) {
    double wu, wv, ut, vt;
    wind_poly_eval(W, ts, &wu, &wv, &ut, &vt);
    D3 r = f3_cold(Pp, wu, wv, pc, z[0], z[1], z[2]);
    dz[0] = r.d0; dz[1] = r.d1; dz[2] = r.d2;
    prop(*Pp, M, z[1], z[2], dz[3], dz[4]);
}

/* LU with partial pivoting, in place, row-major 5x5; false if a pivot is exactly zero */
PM_HD_NOINLINE_DECL bool lu5(double* A, int* piv) {
    for (int k = 0; k < 5; k++) {
        int p = k;
        double big = fabs(A[k * 5 + k]);
        for (int i = k + 1; i < 5; i++)
            if (fabs(A[i * 5 + k]) > big) { big = fabs(A[i * 5 + k]); p = i; }
        piv[k] = p;
        if (p != k)
            for (int j = 0; j < 5; j++) { double tmp = A[k * 5 + j]; A[k * 5 + j] = A[p * 5 + j]; A[p * 5 + j] = tmp; }
        if (A[k * 5 + k] == 0.0) return false;
        for (int i = k + 1; i < 5; i++) {
            A[i * 5 + k] = A[i * 5 + k] / A[k * 5 + k];
            for (int j = k + 1; j < 5; j++) A[i * 5 + j] = A[i * 5 + j] - A[i * 5 + k] * A[k * 5 + j];
        }
    }
    return true;
}
PM_HD_NOINLINE_DECL void lu5_solve(const double* A, const int* piv, double* b) {
    for (int k = 0; k < 5; k++) {
        double tmp = b[k]; b[k] = b[piv[k]]; b[piv[k]] = tmp;
        for (int i = k + 1; i < 5; i++) b[i] = b[i] - A[i * 5 + k] * b[k];
    }
    for (int k = 4; k >= 0; k--) {
        for (int j = k + 1; j < 5; j++) b[k] = b[k] - A[k * 5 + j] * b[j];
        b[k] = b[k] / A[k * 5 + k];
    }
}

PM_HD double rms5_arr(const double* x) {
    double s = 0.0;
    for (int i = 0; i < 5; i++) s += x[i] * x[i];
    return sqrt(s / 5.0);
}

/* one Rosenbrock23 attempt from (u, t) with f0 = f(u, t): unew, f2 = f(unew, t + dt), EEst, eigen_est */
PM_HD_NOINLINE_DECL void rosenbrock23_attempt(const picles_params_t* Pp, const WindPoly* W, const double* M, double pc,
                                              const double* u, double t, double dt, const double* f0, double* unew,
                                              double* f2, double* EEst, double* eigen_est, int32_t* nrhs) {
    const picles_params_t& P = *Pp;
    const double d = 1.0 / (2.0 + sqrt(2.0)), c32 = 6.0 + sqrt(2.0);
    double gam = dt * d, dto2 = dt / 2.0, dto6 = dt / 6.0;
    double wu, wv, wut, wvt;
    wind_poly_eval(W, t, &wu, &wv, &wut, &wvt);
    pmd_t dz[5];
    rhs_dual(Pp, u, wu, wv, M, pc, dz);
    (*nrhs)++;
    double J[25], dT[5], Wm[25];
    double nrm = 0.0;
    for (int i = 0; i < 5; i++) {
        double row = 0.0;
        for (int j = 0; j < 5; j++) {
            J[i * 5 + j] = (j < 3) ? dz[i].d[j] : 0.0; /* the system does not depend on the particle position */
            row += fabs(J[i * 5 + j]);
        }
        nrm = pm_max(nrm, row);
        dT[i] = dz[i].d[3] * wut + dz[i].d[4] * wvt;
    }
    *eigen_est = nrm;
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) Wm[i * 5 + j] = ((i == j) ? 1.0 : 0.0) - gam * J[i * 5 + j];
    int piv[5];
    if (!lu5(Wm, piv)) {
        for (int i = 0; i < 5; i++) { unew[i] = u[i]; f2[i] = f0[i]; }
        *EEst = pm_nan();
        return;
    }
    double k1[5], k2[5], k3[5], f1[5], tmp[5];
    for (int i = 0; i < 5; i++) k1[i] = f0[i] + gam * dT[i];
    lu5_solve(Wm, piv, k1);
    for (int i = 0; i < 5; i++) tmp[i] = u[i] + dto2 * k1[i];
    f5_cold(Pp, W, M, pc, tmp, t + dto2, f1);
    (*nrhs)++;
    for (int i = 0; i < 5; i++) k2[i] = f1[i] - k1[i];
    lu5_solve(Wm, piv, k2);
    for (int i = 0; i < 5; i++) k2[i] = k2[i] + k1[i];
    for (int i = 0; i < 5; i++) unew[i] = u[i] + dt * k2[i];
    f5_cold(Pp, W, M, pc, unew, t + dt, f2);
    (*nrhs)++;
    for (int i = 0; i < 5; i++) k3[i] = f2[i] - c32 * (k2[i] - f1[i]) - 2.0 * (k1[i] - f0[i]) + dt * dT[i];
    lu5_solve(Wm, piv, k3);
    double r[5];
    for (int i = 0; i < 5; i++) {
        double ut = dto6 * (k1[i] - 2.0 * k2[i] + k3[i]);
        double sc = fma(pm_max(fabs(u[i]), fabs(unew[i])), P.reltol, P.abstol);
        r[i] = ut / sc;
    }
    *EEst = rms5_arr(r);
}

/*
 * The attempts of one particle while Rosenbrock23 is the current algorithm: from p.t towards
 * tstop, until the step is complete, the integrator stops (status bits), or AutoSwitch hands the
 * particle back to Tsit5 (as_stiff cleared).  Entry evaluates f0 = f(u, t) — reset_fsal! at the
 * start of a model step, initialize! of the Rosenbrock23 cache after a switch — and, when an
 * auto_dt_reset! is pending, the initial-step heuristic with the order of Rosenbrock23 (2).
 */
PM_HD_NOINLINE_DECL void stiff_integrate(const picles_params_t* Pp, const WindPoly* W, const double* M, double pc,
                                         double tstop, Particle* pp, int* as_count, int* as_stiff, int* attempts_io,
                                         Tally* cp) {
    const picles_params_t& P = *Pp;
    Particle& p = *pp;
    Tally& c = *cp;
    double u[5] = {p.u0, p.u1, p.u2, p.u3, p.u4};
    double t = p.t, dt = p.dt, qold = p.qold;
    int32_t iter = p.iter, nrhs = 0;
    int attempts = *attempts_io;
    int cnt = *as_count;
    bool stiff = true;
    double f0[5];
    f5_cold(Pp, W, M, pc, u, t, f0);
    nrhs++;
    if (p.flags & PICLES_PF_DT_RESET) { /* ode_determine_initdt with get_current_alg_order = 2 */
        p.flags &= (uint8_t)~PICLES_PF_DT_RESET;
        double dtr, dt0, d1n;
        if (initdt_a_cold(P, u[0], u[1], u[2], u[3], u[4], f0[0], f0[1], f0[2], f0[3], f0[4], dtr, dt0, d1n)) {
            dt = dtr;
        } else {
            double u1[5], f1[5];
            for (int i = 0; i < 5; i++) u1[i] = fma(dt0, f0[i], u[i]);
            f5_cold(Pp, W, M, pc, u1, t + dt0, f1);
            nrhs++;
            dt = initdt_b_cold(P, u[0], u[1], u[2], u[3], u[4], f0[0], f0[1], f0[2], f0[3], f0[4], f1[0], f1[1], f1[2], f1[3],
                               f1[4], dt0, d1n, 2.0);
        }
    }
    const double qmin = 0.2, qmax = 10.0, gamma = 0.9;
    const double beta1 = 0.35, beta2 = 0.2;
    while (t < tstop) {
        iter++;
        const double dtmin_t = pm_max(pm_eps(t), P.dtmin);
        dt = pm_min(P.dtmax, dt);
        dt = pm_max(dt, dtmin_t);
        dt = pm_min(dt, tstop - t);
        if (dt != dt) { p.status |= PICLES_PST_UNSTABLE; c.failed++; break; }
        if ((int64_t)iter > P.maxiters) { p.status |= PICLES_PST_MAXITERS; c.failed++; break; }
        if (!P.force_dtmin && dt <= P.dtmin && (t + dt < tstop)) { p.status |= PICLES_PST_DTMIN; c.failed++; break; }
        attempts++;
        double un[5], fnew[5], EEst, eig;
        rosenbrock23_attempt(Pp, W, M, pc, u, t, dt, f0, un, fnew, &EEst, &eig, &nrhs);
        c.stiff_attempts++;
        double q, q11 = 1.0;
        if (EEst == 0.0) {
            q = 1.0 / qmax;
        } else {
            double t1 = beta1 * pm_log_safe(EEst);
            q11 = pm_exp(t1);
            q = pm_exp(t1 - beta2 * pm_log_safe(qold));
            q = pm_max(1.0 / qmax, pm_min(1.0 / qmin, q / gamma));
        }
        bool accept = (EEst <= 1.0) || (P.force_dtmin && fabs(dt) <= dtmin_t);
        if (accept) {
            qold = pm_max(EEst, PH_QOLDINIT);
            double dtnew = dt / q;
            double ttmp = t + dt;
            t = (fabs(ttmp - tstop) < 100.0 * pm_eps(tstop)) ? tstop : ttmp;
            double dtp = pm_min(P.dtmax, dtnew);
            dtp = pm_max(dtp, pm_max(pm_eps(t), P.dtmin));
            dt = dtp;
            bool bad = false;
            for (int i = 0; i < 5; i++) { u[i] = un[i]; f0[i] = fnew[i]; bad = bad || (un[i] != un[i]); }
            c.substeps++;
            if (bad) { p.status |= PICLES_PST_UNSTABLE; c.failed++; break; }
        } else {
            dt = dt / pm_reject_factor(P.nan_eest_rejects, q11, qmin, gamma);
            c.rejects++;
        }
        /* AutoSwitch with eigen_est = opnorm(J, Inf) */
        double stiffness = fabs(eig * dt / 3.5068);
        bool is = stiffness > 0.9;
        cnt = is ? ((cnt < 0) ? 1 : cnt + 1) : ((cnt > 0) ? -1 : cnt - 1);
        if (cnt > 60) cnt = 60;
        if (cnt < -60) cnt = -60;
        if (cnt < -3) { /* back to Tsit5: dt / dtfac; its initialize! re-evaluates fsalfirst in the caller */
            dt = dt / 2.0;
            stiff = false;
            break;
        }
    }
    p.u0 = u[0]; p.u1 = u[1]; p.u2 = u[2]; p.u3 = u[3]; p.u4 = u[4];
    p.t = t; p.dt = dt; p.qold = qold; p.iter = iter;
    c.rhs += nrhs;
    *as_count = cnt;
    *as_stiff = stiff ? 1 : 0;
    *attempts_io = attempts;
}

/* declared in physics.h */
PM_HD_NOINLINE_DECL bool stiff_phase(const picles_params_t* Pp, StiffArgs* a, double pc, double tstop) {
    stiff_integrate(Pp, &a->W, a->M, pc, tstop, &a->p, &a->as_count, &a->as_stiff, &a->attempts, &a->c);
    return !a->as_stiff && (a->p.t < tstop) &&
           !(a->p.status & (PICLES_PST_MAXITERS | PICLES_PST_DTMIN | PICLES_PST_UNSTABLE));
}

} /* namespace picles */
#endif /* PICLES_STIFF_H */
