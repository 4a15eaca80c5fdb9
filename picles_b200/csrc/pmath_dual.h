/*
 * pmath_dual.h — forward-mode dual numbers over the deterministic pmath functions.
 *
 * Used only by the stiff branch of AutoTsit5(Rosenbrock23()) (stiff.h): OrdinaryDiffEq's
 * Rosenbrock23 takes the Jacobian df/du and the time derivative df/dt of the particle system
 * by automatic differentiation (ForwardDiff: autodiff = true is its default), i.e. the exact
 * derivative of the same operation sequence.  Like pmath.h this is a math library shared by the
 * device code and the CPU checker, so both differentiate with bit-identical arithmetic: plain
 * C99, straight-line, IEEE operators only (the stiff branch is rare — it is never on the hot
 * path — so no fast paths here).  Derivative rules follow ForwardDiff/DiffRules:
 *   max(a, c): the derivative of a where a >= c (ties to a), else 0;  abs: sign(a) a';
 *   sqrt: a' / (2 sqrt a);  a / b: (a' - (a/b) b') / b;  tanh: (1 - tanh^2) a';
 *   sech: -tanh sech a';  exp: exp a'.
 */
#ifndef PICLES_PMATH_DUAL_H
#define PICLES_PMATH_DUAL_H

#include "pmath.h"

#define PMD_N 5 /* partials: lne, c_x, c_y, wind u, wind v */

typedef struct {
    double v;
    double d[PMD_N];
} pmd_t;

PM_HD pmd_t pmd_const(double c) {
    pmd_t r;
    r.v = c;
    for (int i = 0; i < PMD_N; i++) r.d[i] = 0.0;
    return r;
}
PM_HD pmd_t pmd_var(double v, int k) {
    pmd_t r = pmd_const(v);
    r.d[k] = 1.0;
    return r;
}
PM_HD pmd_t pmd_add(pmd_t a, pmd_t b) {
    pmd_t r;
    r.v = a.v + b.v;
    for (int i = 0; i < PMD_N; i++) r.d[i] = a.d[i] + b.d[i];
    return r;
}
PM_HD pmd_t pmd_sub(pmd_t a, pmd_t b) {
    pmd_t r;
    r.v = a.v - b.v;
    for (int i = 0; i < PMD_N; i++) r.d[i] = a.d[i] - b.d[i];
    return r;
}
PM_HD pmd_t pmd_neg(pmd_t a) {
    pmd_t r;
    r.v = -a.v;
    for (int i = 0; i < PMD_N; i++) r.d[i] = -a.d[i];
    return r;
}
PM_HD pmd_t pmd_addc(pmd_t a, double c) {
    pmd_t r = a;
    r.v = a.v + c;
    return r;
}
/* c * a */
PM_HD pmd_t pmd_scale(double c, pmd_t a) {
    pmd_t r;
    r.v = c * a.v;
    for (int i = 0; i < PMD_N; i++) r.d[i] = c * a.d[i];
    return r;
}
PM_HD pmd_t pmd_mul(pmd_t a, pmd_t b) {
    pmd_t r;
    r.v = a.v * b.v;
    for (int i = 0; i < PMD_N; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
PM_HD pmd_t pmd_sqr(pmd_t a) {
    pmd_t r;
    const double t = 2.0 * a.v;
    r.v = a.v * a.v;
    for (int i = 0; i < PMD_N; i++) r.d[i] = t * a.d[i];
    return r;
}
PM_HD pmd_t pmd_div(pmd_t a, pmd_t b) {
    pmd_t r;
    const double q = a.v / b.v;
    r.v = q;
    for (int i = 0; i < PMD_N; i++) r.d[i] = (a.d[i] - q * b.d[i]) / b.v;
    return r;
}
/* c / b */
PM_HD pmd_t pmd_cdiv(double c, pmd_t b) { return pmd_div(pmd_const(c), b); }
PM_HD pmd_t pmd_sqrt(pmd_t a) {
    pmd_t r;
    const double s = sqrt(a.v);
    const double t = 2.0 * s;
    r.v = s;
    for (int i = 0; i < PMD_N; i++) r.d[i] = a.d[i] / t;
    return r;
}
PM_HD pmd_t pmd_abs(pmd_t a) { return (pm_d2i(a.v) < 0) ? pmd_neg(a) : a; }
/* max(a, c) for a constant c */
PM_HD pmd_t pmd_maxc(pmd_t a, double c) { return (a.v >= c || a.v != a.v) ? a : pmd_const(c); }
PM_HD pmd_t pmd_exp(pmd_t a) {
    pmd_t r;
    const double e = pm_exp(a.v);
    r.v = e;
    for (int i = 0; i < PMD_N; i++) r.d[i] = e * a.d[i];
    return r;
}
PM_HD pmd_t pmd_tanh(pmd_t a) {
    pmd_t r;
    const double t = pm_tanh_safe(a.v);
    const double g = 1.0 - t * t;
    r.v = t;
    for (int i = 0; i < PMD_N; i++) r.d[i] = g * a.d[i];
    return r;
}
PM_HD pmd_t pmd_sech(pmd_t a) {
    pmd_t r;
    const double s = pm_sech_safe(a.v);
    const double g = -(pm_tanh_safe(a.v) * s);
    r.v = s;
    for (int i = 0; i < PMD_N; i++) r.d[i] = g * a.d[i];
    return r;
}
/* a^y for a real exponent y: value as exp(y log a) would lose the exactness of the small
   integer powers the path uses, so the caller passes the value and a^(y-1) */
PM_HD pmd_t pmd_pow_given(pmd_t a, double y, double value, double value_ym1) {
    pmd_t r;
    const double g = y * value_ym1;
    r.v = value;
    for (int i = 0; i < PMD_N; i++) r.d[i] = g * a.d[i];
    return r;
}

#endif /* PICLES_PMATH_DUAL_H */
