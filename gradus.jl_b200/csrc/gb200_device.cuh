// gb200_device.cuh -- FP64 device (and shared host) building blocks of the trace kernel.
//
// Everything the reference evaluates per ray is written out in closed form here so the
// whole integrator state lives in registers:
//   * metric components + r/theta Jacobian (reference: ForwardDiff over
//     src/metrics/kerr-metric.jl:11-28 and johannsen-psaltis-ad.jl:4-26, via
//     metric_jacobian, src/tracing/method-implementations/auto-diff.jl:206-211)
//   * the static-axisymmetric geodesic acceleration (auto-diff.jl:59-76,115-141)
//   * LNRF (ZAMO) initial conditions (src/tracing/utility.jl:13-40 with the tetrad of
//     src/orthonormalization.jl:116-122 in closed form) and the v^t constraint
//     (auto-diff.jl:161-173)
//   * disc conditions (src/geometry/discs/*.jl) and the redshift point function
//     (src/redshift.jl, src/orbits/circular-orbits.jl)
#pragma once
#include <cstdint>
#include <math.h>
#include "../../include/gradus_b200.h"

#ifdef __CUDACC__
#define GB_HD __host__ __device__
#define GB_D __device__ __forceinline__
#else
#define GB_HD
#define GB_D inline
#endif

#include "jp_metric_generated.cuh"

#define GB_MAX_PF 4

// ---------------------------------------------------------------- kernel parameter block
struct GbParams {
    // metric
    int32_t metric_kind;
    double M, a, eps3;
    // integrator
    double lam0, lam1, abstol, reltol, dtmax, mu;
    int64_t maxiters;
    int32_t pow_mode;
    // termination
    int32_t geometry_kind;
    double gp0, gp1, gp2, gtol;
    double chart_inner, chart_outer;
    int32_t callback_kind;
    double callback_delta;
    // initial conditions
    int32_t ic_kind, grid_kind;
    int64_t width, height;
    double lo0, step0_hi, step0_lo; // alpha (or r) axis: value = lo + i*step, step in double-double
    double lo1, step1_hi, step1_lo; // beta (or theta) axis
    double geoK;                    // geometric grid ratio
    double inv_lo_hi, inv_step_hi, inv_step_lo; // inverse grid: range(1/max, 1/min, N)
    double xo[4];                   // observer position
    // observer LNRF constants: v^r = c_r p_r, v^th = c_th p_th, v^ph = c_ph0 + c_ph1 p_ph
    double c_r, c_th, c_ph0, c_ph1;
    double go[5];                   // metric components at the observer (for constrain_time and E_obs)
    const double* ex[4];            // explicit ICs (device SoA), indexed by global ray id
    const double* ev[4];
    // ray range
    int64_t first, count, stride;
    // outputs, indexed by slot n in [0, count)
    int32_t* o_status;
    double* o_lambda;
    double* o_x[4];
    double* o_v[4];
    double* o_x0[4];
    double* o_v0[4];
    int32_t* o_naccept;
    int32_t* o_nreject;
    int32_t* o_flags;
    int32_t npf;
    int32_t pf[GB_MAX_PF];
    double* o_img[GB_MAX_PF];
    // line profile: per-ray (g, f) written for the binning kernel
    double* o_g;
    double* o_f;
    double min_re, max_re;
    int32_t emis_kind, emis_n;
    double emis_index;
    const double* emis_r;
    const double* emis_eps;
    // redshift
    double r_isco;
    int32_t pl_n;
    const double* pl_r;
    const double* pl_ut;
    const double* pl_ur;
    const double* pl_uphi;
    // work queue + global counters
    unsigned long long* queue;    // next slot to hand out
    unsigned long long* counters; // [0] accepted, [1] rejected, [2] flagged rays
};

// ---------------------------------------------------------------- Tsit5 constants (OrdinaryDiffEq Tsit5ConstantCache)
#define GB_A21 0.161
#define GB_A31 -0.008480655492356989
#define GB_A32 0.335480655492357
#define GB_A41 2.8971530571054935
#define GB_A42 -6.359448489975075
#define GB_A43 4.3622954328695815
#define GB_A51 5.325864828439257
#define GB_A52 -11.748883564062828
#define GB_A53 7.4955393428898365
#define GB_A54 -0.09249506636175525
#define GB_A61 5.86145544294642
#define GB_A62 -12.92096931784711
#define GB_A63 8.159367898576159
#define GB_A64 -0.071584973281401
#define GB_A65 -0.028269050394068383
#define GB_A71 0.09646076681806523
#define GB_A72 0.01
#define GB_A73 0.4798896504144996
#define GB_A74 1.379008574103742
#define GB_A75 -3.290069515436081
#define GB_A76 2.324710524099774
#define GB_BT1 -0.00178001105222577714
#define GB_BT2 -0.0008164344596567469
#define GB_BT3 0.007880878010261995
#define GB_BT4 -0.1447110071732629
#define GB_BT5 0.5823571654525552
#define GB_BT6 -0.45808210592918697
#define GB_BT7 0.015151515151515152
// dense output: b_1 = Th (1 + Th (R12 + Th (R13 + Th R14))), b_j = Th^2 (Rj2 + Th (Rj3 + Th Rj4))
#define GB_R12 -2.763706197274826
#define GB_R13 2.9132554618219126
#define GB_R14 -1.0530884977290216
#define GB_R22 0.13169999999999998
#define GB_R23 -0.2234
#define GB_R24 0.1017
#define GB_R32 3.9302962368947516
#define GB_R33 -5.941033872131505
#define GB_R34 2.490627285651253
#define GB_R42 -12.411077166933676
#define GB_R43 30.33818863028232
#define GB_R44 -16.548102889244902
#define GB_R52 37.50931341651104
#define GB_R53 -88.1789048947664
#define GB_R54 47.37952196281928
#define GB_R62 -27.896526289197286
#define GB_R63 65.09189467479366
#define GB_R64 -34.87065786149661
#define GB_R72 1.5
#define GB_R73 -4.0
#define GB_R74 2.5

// ---------------------------------------------------------------- metrics
// Kerr: components and Jacobian from w = 2Mr/Sigma (see DESIGN.md for the derivation).
template <class S>
GB_HD inline void kerr_metric_jacobian(double M, double a, S r, S s, S c, S g[5], S dr[5], S dth[5]) {
    const double a2 = a * a;
    const S r2 = r * r, s2 = s * s, c2 = c * c;
    const S sin2 = 2.0 * s * c;
    const S Sig = r2 + a2 * c2;
    const S Del = r2 - 2.0 * M * r + a2;
    const S iSig = 1.0 / Sig, iDel = 1.0 / Del;
    const S w = 2.0 * M * r * iSig;
    const S w_r = 2.0 * (M - w * r) * iSig;
    const S w_t = w * a2 * sin2 * iSig;
    const S Sig_t = -a2 * sin2;
    const S B = r2 + a2 + a2 * s2 * w;
    const S q = sin2 * w + s2 * w_t;
    g[0] = w - 1.0;
    g[1] = Sig * iDel;
    g[2] = Sig;
    g[3] = s2 * B;
    g[4] = -a * s2 * w;
    dr[0] = w_r;
    dr[1] = (2.0 * r - g[1] * 2.0 * (r - M)) * iDel;
    dr[2] = 2.0 * r;
    dr[3] = s2 * (2.0 * r + a2 * s2 * w_r);
    dr[4] = -a * s2 * w_r;
    dth[0] = w_t;
    dth[1] = Sig_t * iDel;
    dth[2] = Sig_t;
    dth[3] = sin2 * B + s2 * a2 * q;
    dth[4] = -a * q;
}

template <int METRIC>
GB_HD inline void metric_jacobian(const GbParams& P, double r, double s, double c, double g[5], double dr[5], double dth[5]) {
    if (METRIC == GB200_METRIC_KERR) kerr_metric_jacobian(P.M, P.a, r, s, c, g, dr, dth);
    else jp_metric_jacobian(P.M, P.a, P.eps3, r, s, c, g, dr, dth);
}
GB_HD inline void metric_jacobian_rt(const GbParams& P, double r, double s, double c, double g[5], double dr[5], double dth[5]) {
    if (P.metric_kind == GB200_METRIC_KERR) kerr_metric_jacobian(P.M, P.a, r, s, c, g, dr, dth);
    else jp_metric_jacobian(P.M, P.a, P.eps3, r, s, c, g, dr, dth);
}

// a^mu = -g^{mu m} ( gdot_{mk} v^k - 1/2 S_m ),  gdot = v^r d_r g + v^th d_th g,  S_m = d_m g_{kl} v^k v^l
// (the contraction of auto-diff.jl:115-141 with d_t = d_phi = 0 written out)
GB_HD inline void geodesic_accel(const double g[5], const double dr[5], const double dth[5],
                                 double vt, double vr, double vth, double vph, double acc[4]) {
    const double D = g[0] * g[3] - g[4] * g[4];
    const double iD = 1.0 / D;
    const double gitt = g[3] * iD, giphph = g[0] * iD, gitph = -g[4] * iD;
    const double girr = 1.0 / g[1], githth = 1.0 / g[2];
    const double d0 = vr * dr[0] + vth * dth[0];
    const double d1 = vr * dr[1] + vth * dth[1];
    const double d2 = vr * dr[2] + vth * dth[2];
    const double d3 = vr * dr[3] + vth * dth[3];
    const double d4 = vr * dr[4] + vth * dth[4];
    const double Pt = d0 * vt + d4 * vph;
    const double Pp = d4 * vt + d3 * vph;
    const double vtt = vt * vt, vrr = vr * vr, vthth = vth * vth, vpp = vph * vph, vtp2 = 2.0 * vt * vph;
    const double Sr = dr[0] * vtt + dr[1] * vrr + dr[2] * vthth + dr[3] * vpp + dr[4] * vtp2;
    const double St = dth[0] * vtt + dth[1] * vrr + dth[2] * vthth + dth[3] * vpp + dth[4] * vtp2;
    acc[0] = -(gitt * Pt + gitph * Pp);
    acc[1] = -girr * (d1 * vr - 0.5 * Sr);
    acc[2] = -githth * (d2 * vth - 0.5 * St);
    acc[3] = -(gitph * Pt + giphph * Pp);
}

#ifdef __CUDACC__
template <int METRIC>
GB_D void rhs_accel(const GbParams& P, double r, double th, double vt, double vr, double vth, double vph,
                    double acc[4], double& s, double& c) {
    sincos(th, &s, &c);
    double g[5], dr[5], dth[5];
    metric_jacobian<METRIC>(P, r, s, c, g, dr, dth);
    geodesic_accel(g, dr, dth, vt, vr, vth, vph, acc);
}
#endif

GB_HD inline void metric_components_rt(const GbParams& P, double r, double th, double g[5]) {
    double dr[5], dth[5];
    metric_jacobian_rt(P, r, sin(th), cos(th), g, dr, dth);
}

// constrain_time, auto-diff.jl:161-173
GB_HD inline double constrain_vt(const double g[5], double vr, double vth, double vph, double mu) {
    const double disc = -g[0] * g[1] * vr * vr - g[0] * g[2] * vth * vth - g[0] * mu * mu - (g[0] * g[3] - g[4] * g[4]) * vph * vph;
    return -(g[4] * vph + sqrt(disc)) / g[0];
}

// ---------------------------------------------------------------- disc conditions (distance_to_disc)
template <int GEOM>
GB_HD inline double disc_condition(const GbParams& P, double r, double s, double c) {
    if (GEOM == GB200_GEOMETRY_THIN_DISC) { // thin-disc.jl:20-26 with _gtol_error = gtol*|r| (discs.jl:7)
        const double rho = r * fabs(s);
        if (rho < P.gp0 || rho > P.gp1) return 1.0;
        return r * fabs(c) - P.gtol * fabs(r);
    } else if (GEOM == GB200_GEOMETRY_SHAKURA_SUNYAEV) { // thick-disc.jl:57-63, shakura-sunyaev.jl:28-33
        const double rho = r * fabs(s);
        if (rho < P.gp2) return 1.0;
        const double h = 3.0 * P.gp1 * P.gp0 * (1.0 - sqrt(P.gp2 / rho));
        if (h <= 0.0) return 1.0;
        return r * fabs(c) - h;
    } else if (GEOM == GB200_GEOMETRY_DATUM_PLANE) { // datum-plane.jl:6-10
        return r * c - P.gp0;
    }
    return 1.0;
}

// ---------------------------------------------------------------- image-plane -> initial state
GB_HD inline double dd_axis(double lo, double shi, double slo, int64_t i) { // lo + i*step with a double-double step
    const double fi = (double)i;
#ifdef __CUDA_ARCH__
    const double p = fi * shi;
    const double e = fma(fi, shi, -p) + fi * slo;
#else
    const double p = fi * shi;
    const double e = fma(fi, shi, -p) + fi * slo;
#endif
    const double sum = lo + p;
    const double bb = sum - lo;
    const double err = (lo - (sum - bb)) + (p - bb);
    return sum + (err + e);
}

struct GbRayInit {
    double x[4];
    double v[4];
    double area;
};

// Initial state for global ray id `i`: velfunc(i) then constrain_all (constraints.jl:14-15).
GB_HD inline void ray_initial_state(const GbParams& P, int64_t i, GbRayInit& o) {
    o.area = 1.0;
    if (P.ic_kind == GB200_IC_EXPLICIT) {
        for (int k = 0; k < 4; ++k) { o.x[k] = P.ex[k][i]; o.v[k] = P.ev[k][i]; }
        double g[5];
        metric_components_rt(P, o.x[1], o.x[2], g);
        o.v[0] = constrain_vt(g, o.v[1], o.v[2], o.v[3], P.mu);
        return;
    }
    double alpha, beta;
    if (P.ic_kind == GB200_IC_RENDER_GRID) { // rendering.jl:150-159
        const int64_t col = i / P.height, row = i - col * P.height;
        alpha = dd_axis(P.lo0, P.step0_hi, P.step0_lo, col) + 1e-6;
        beta = dd_axis(P.lo1, P.step1_hi, P.step1_lo, row) + 1e-6;
    } else { // polar plane, planes.jl:93-115
        const int64_t j = i / P.width, k = i - j * P.width;
        double rr;
        if (P.grid_kind == GB200_GRID_GEOMETRIC) rr = P.lo0 * pow(P.geoK, (double)k);
        else if (P.grid_kind == GB200_GRID_INVERSE) rr = 1.0 / dd_axis(P.inv_lo_hi, P.inv_step_hi, P.inv_step_lo, P.width - 1 - k);
        else rr = dd_axis(P.lo0, P.step0_hi, P.step0_lo, k);
        const double th = dd_axis(P.lo1, P.step1_hi, P.step1_lo, j);
        alpha = rr * cos(th);
        beta = rr * sin(th);
        o.area = rr * rr;
    }
    for (int k = 0; k < 4; ++k) o.x[k] = P.xo[k];
    // local_momentum (utility.jl:13-20) and the closed-form LNRF transform
    const double b = beta / P.xo[1], a = alpha / P.xo[1];
    const double pr = -1.0 / sqrt(1.0 + a * a + b * b);
    const double pth = b * pr, pph = a * pr;
    o.v[1] = P.c_r * pr;
    o.v[2] = P.c_th * pth;
    o.v[3] = P.c_ph0 + P.c_ph1 * pph;
    o.v[0] = constrain_vt(P.go, o.v[1], o.v[2], o.v[3], P.mu);
}

// ---------------------------------------------------------------- endpoint point functions
// CircularOrbits.fourvelocity at (rho, pi/2): circular-orbits.jl:11-37,58-61,114-123
GB_HD inline void circular_fourvelocity(const GbParams& P, double rho, double& ut_up, double& uph_up) {
    double g[5], dr[5], dth[5];
    metric_jacobian_rt(P, rho, 1.0, 0.0, g, dr, dth);
    const double D = g[0] * g[3] - g[4] * g[4];
    const double iD = 1.0 / D;
    const double gitt = g[3] * iD, giphph = g[0] * iD, gitph = -g[4] * iD;
    const double disc = sqrt(dr[4] * dr[4] - dr[0] * dr[3]);
    const double Om = -(dr[4] - disc) / dr[3];
    const double A = -(Om * gitt - gitph);
    const double B = (Om * gitph - giphph);
    const double denom = B * B * gitt + 2.0 * A * B * gitph + A * A * giphph;
    const double sg = (denom > 0.0) ? 1.0 : ((denom < 0.0) ? -1.0 : 0.0);
    const double d = -sg * sqrt(1.0 / fabs(denom));
    const double ut = B * d, uph = A * d; // covariant
    ut_up = gitt * ut + gitph * uph;
    uph_up = gitph * ut + giphph * uph;
}

GB_HD inline double table_lerp(const double* xs, const double* ys, int n, double x) { // NaNLinearInterpolator, interpolations.jl:7-14 (clamped abscissa)
    x = fmin(fmax(x, xs[0]), xs[n - 1]);
    int lo = 0, hi = n - 1; // last index with xs[idx] <= x, clamped to [0, n-2]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (xs[mid] <= x) lo = mid; else hi = mid; }
    const double w = (x - xs[lo]) / (xs[lo + 1] - xs[lo]);
    return (1.0 - w) * ys[lo] + w * ys[lo + 1];
}

// redshift_function (src/redshift.jl:192-220; plunging region :93-164; generic: :246-276)
GB_HD inline double redshift_endpoint(const GbParams& P, const double x[4], const double v[4], const double v0[4], const double go[5]) {
    const double sth = sin(x[2]);
    const double rho = x[1] * fabs(sth);
    double u0, u1 = 0.0, u3;
    if (rho < P.r_isco) {
        if (P.metric_kind == GB200_METRIC_KERR) { // Cunningham (1975) plunging flow
            const double M = P.M, a = P.a, rms = P.r_isco, r = rho;
            const double sM = sqrt(M), srms = sqrt(rms);
            const double Le = sM * (rms * rms - 2.0 * a * sqrt(M * rms) + a * a) / (rms * srms - 2.0 * M * srms + a * sM);
            const double Del = r * r - 2.0 * M * r + a * a;
            const double H = (2.0 * M * r - a * Le) / Del;
            const double ge = sqrt(1.0 - (2.0 * M) / (3.0 * rms));
            const double qq = rms / r - 1.0;
            const double ur = -sqrt((2.0 * M) / (3.0 * rms)) * (qq * sqrt(qq));
            u3 = ge / (r * r) * (Le + a * H);
            u0 = ge * (1.0 + 2.0 * M * (1.0 + H) / r);
            u1 = -ur;
        } else {
            if (P.pl_n < 2) return nan("");
            u0 = table_lerp(P.pl_r, P.pl_ut, P.pl_n, rho);
            u1 = -table_lerp(P.pl_r, P.pl_ur, P.pl_n, rho);
            u3 = table_lerp(P.pl_r, P.pl_uphi, P.pl_n, rho);
        }
    } else {
        circular_fourvelocity(P, rho, u0, u3);
    }
    double g[5], dr[5], dth[5];
    metric_jacobian_rt(P, x[1], sth, cos(x[2]), g, dr, dth);
    const double Ed = (g[0] * v[0] + g[4] * v[3]) * u0 + (g[1] * v[1]) * u1 + (g[4] * v[0] + g[3] * v[3]) * u3;
    const double Eo = go[0] * v0[0] + go[4] * v0[3]; // metric at x_init
    return Eo / Ed;
}

GB_HD inline double emissivity_eval(const GbParams& P, double rho) {
    if (P.emis_kind == GB200_EMISSIVITY_POWERLAW) return pow(rho, -P.emis_index);
    return table_lerp(P.emis_r, P.emis_eps, P.emis_n, rho);
}
