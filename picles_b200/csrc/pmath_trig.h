/*
 * pmath_trig.h — deterministic sin/cos (radians) and tand (degrees) for the grid-metric
 * lookup: the per-node projection kernel [cosα/dx sinα/dy; −sinα/dx cosα/dy]
 * (src/Grids/TripolarGridMOM6.jl:448-459) and the great-circle coefficient
 * sign(φ)·min(sign(φ)·tand(φ),60)/R (src/Grids/spherical_grid_corrections.jl:13) feed every
 * right-hand side, so the device (k_grid_metric) and the CPU oracle must form them with the
 * same bits.  Only IEEE + − × ÷ and fma are used (no libm): identical on sm_100a
 * (--fmad=false) and on the host (-ffp-contract=off).
 *
 * Algorithms: Cody–Waite reduction by π/2 in three parts (valid for |x| < 2^20·π/2, far
 * beyond any angle in degrees·π/180) and the fdlibm/musl minimax kernels on [−π/4, π/4]
 * with a tail term.  Accuracy < 1 ulp against libm on the tested ranges
 * (tests/test_pmath.py).
 */
#ifndef PICLES_PMATH_TRIG_H
#define PICLES_PMATH_TRIG_H

#include "pmath.h"

/* sin(x + y) on |x| <= pi/4, y the tail of x */
PM_HD double pm_ksin(double x, double y) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = x * x;
    double w = z * z;
    double r = S2 + z * (S3 + z * S4) + z * w * (S5 + z * S6);
    double v = z * x;
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
/* cos(x + y) on |x| <= pi/4 */
PM_HD double pm_kcos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double w = z * z;
    double r = z * (C1 + z * (C2 + z * C3)) + (w * w) * (C4 + z * (C5 + z * C6));
    double hz = 0.5 * z;
    w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}

/* quadrant selection: n mod 4 */
PM_HD void pm_quadrant(int n, double ks, double kc, double* s, double* c) {
    int q = n & 3;
    double ss = (q & 1) ? kc : ks;
    double cc = (q & 1) ? ks : kc;
    ss = (q & 2) ? -ss : ss;
    cc = ((q == 1) || (q == 2)) ? -cc : cc;
    *s = ss;
    *c = cc + 0.0; /* -0 -> +0, as cosd/cos return for exact zeros */
}

/* sin and cos of x radians, |x| < 1e6 */
PM_HD void pm_sincos(double x, double* s, double* c) {
    const double invpio2 = 6.36619772367581382433e-01, toint = 6755399441055744.0;
    const double pio2_1 = 1.57079632673412561417e+00, pio2_2 = 6.07710050630396597660e-11,
                 pio2_2t = 2.02226624879595063154e-21;
    double fn = (x * invpio2 + toint) - toint;
    int n = (int)fn;
    /* two rounds of Cody–Waite: x - fn*(pio2_1 + pio2_2 + pio2_2t), 118 bits of pi/2 */
    double t = x - fn * pio2_1;
    double w = fn * pio2_2;
    double r = t - w;
    w = fn * pio2_2t - ((t - r) - w);
    double y0 = r - w;
    double y1 = (r - y0) - w;
    pm_quadrant(n, pm_ksin(y0, y1), pm_kcos(y0, y1), s, c);
    if (!(fabs(x) < 1.0e6)) { *s = pm_nan(); *c = pm_nan(); }
}

/* sind, cosd: exact reduction in degrees, then radians as a double-double */
PM_HD void pm_sincosd(double deg, double* s, double* c) {
    const double D_HI = 0.017453292519943295, D_LO = 2.9486522708701687e-19; /* pi/180 = D_HI + D_LO */
    const double toint = 6755399441055744.0;
    double fn = (deg / 90.0 + toint) - toint;
    int n = (int)fn;
    double r = fma(-90.0, fn, deg); /* exact: |deg| < 2^50 and r is a multiple of ulp(deg) */
    double hi = r * D_HI;
    double lo = fma(r, D_HI, -hi) + r * D_LO;
    pm_quadrant(n, pm_ksin(hi, lo), pm_kcos(hi, lo), s, c);
    if (!(fabs(deg) < 1.0e9)) { *s = pm_nan(); *c = pm_nan(); }
}
PM_HD double pm_tand(double deg) {
    double s, c;
    pm_sincosd(deg, &s, &c);
    return s / c;
}

/* ---- one node of the grid metric ------------------------------------------------------ */
/* M = [cosα/dx sinα/dy; −sinα/dx cosα/dy], α = angle_dx·π/180 (TripolarGridMOM6.jl:448-459);
   pc = sign(φ)·min(sign(φ)·tand(φ), 60)/R (spherical_grid_corrections.jl:13) */
PM_HD void pm_grid_metric_node(double dx, double dy, double angle_deg, double lat_deg, double R_earth, double* M11,
                               double* M12, double* M21, double* M22, double* pc) {
    double a = (angle_deg * 3.141592653589793) / 180.0;
    double s, c;
    pm_sincos(a, &s, &c);
    *M11 = c / dx;
    *M12 = s / dy;
    *M21 = -s / dx;
    *M22 = c / dy;
    double sg = (lat_deg > 0.0) ? 1.0 : ((lat_deg < 0.0) ? -1.0 : lat_deg); /* Julia sign(): ±0 and NaN pass through */
    double t = pm_tand(lat_deg);
    *pc = (sg * pm_min(sg * t, 60.0)) / R_earth;
}

#endif /* PICLES_PMATH_TRIG_H */
