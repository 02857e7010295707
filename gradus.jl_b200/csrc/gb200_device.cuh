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

// Branch-free FP64 reciprocal: hardware seed (rcp.approx, >= 20 good bits) + two Newton steps -> <= 1-2 ulp.
// Used for the well-conditioned denominators of the hot loop (Sigma, Delta, sin^2, error scales); an IEEE
// division costs ~25 instructions including a divergent slow-path check, this costs 5.
#ifndef GB_OPT_RCP3
#define GB_OPT_RCP3 1 /* one third-order correction y (1 + e + e^2) instead of two Newton steps (3 DFMA instead of 4) */
#endif
GB_HD inline double gb_rcp(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if GB_OPT_RCP3
    // 1/x = y / (1 - e) = y (1 + e + e^2 + e^3 + ...), e = 1 - x y: |e| <= 2^-20 for the hardware seed, so the
    // truncation e^3 is below 2^-60
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
#else
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#endif
#else
    return 1.0 / x;
#endif
}

// One Newton step only (relative error ~2^-40): for the error-norm scales, whose reciprocal feeds an estimate that
// already carries ~1e-7 of rounding noise.
GB_HD inline double gb_rcp_lo(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, e, y);
#else
    return 1.0 / x;
#endif
}

// max / min as one compare and a select.  fmax()/fmin() cost ~8 instructions each on sm_100a (IEEE nan quieting);
// these return b when either argument is a nan, and every use below either cannot see a nan or lets it propagate
// into the step-size, where the integrator's nan check flags the ray.
GB_HD inline double gb_max(double a, double b) { return a > b ? a : b; }
GB_HD inline double gb_min(double a, double b) { return a < b ? a : b; }

// Branch-free square root for x >= 1e-300 (callers clamp): rsqrt seed + two coupled Newton steps + a final residual
// correction, <= 1 ulp.  The library sqrt() keeps a slow-path call whose register save/restore spilled in the hot loop.
GB_HD inline double gb_sqrt_pos(double x) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double e = fma(-h, g, 0.5);
    g = fma(g, e, g); h = fma(h, e, h);
    e = fma(-h, g, 0.5);
    g = fma(g, e, g); h = fma(h, e, h);
    const double d = fma(-g, g, x);
    return fma(d, h, g);
#else
    return sqrt(x);
#endif
}

#include "jp_metric_generated.cuh"
#include "metrics_generated.cuh"

#define GB_MAX_PF 4

// ---------------------------------------------------------------- kernel parameter block
struct GbParams {
    // metric
    int32_t metric_kind;
    double M, a, eps3;
    double a2, twoM; // a^2 and 2M: read from the constant bank at every RHS evaluation instead of recomputed
    double jp_e;     // Johannsen-Psaltis: eps3 M^3
    double mp[8]; // all metric parameters in the order of include/gradus_b200.h (M = mp[0], a = mp[1])
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
    double lo0, step0_hi, step0_lo; // alpha (or r / x) axis: value = lo + i*step, step in double-double
    double lo1, step1_hi, step1_lo; // beta (or theta / y) axis
    double geoK, geoK1;             // geometric grid ratios of axis 0 / axis 1
    double inv_lo_hi, inv_step_hi, inv_step_lo;    // inverse grid of axis 0: range(1/max, 1/min, N)
    double inv1_lo_hi, inv1_step_hi, inv1_step_lo; // inverse grid of axis 1 (cartesian plane)
    int64_t n0, n1;                 // points per axis grid (cartesian plane: Nx/2, Ny/2)
    double xo[4];                   // observer position
    // observer LNRF constants: v^r = c_r p_r, v^th = c_th p_th, v^ph = c_ph0 + c_ph1 p_ph
    double c_r, c_th, c_ph0, c_ph1;
    double go[5];                   // metric components at the observer (for constrain_time and E_obs)
    const double* ex[4];            // explicit ICs (device SoA), indexed by global ray id
    const double* ev[4];
    // ray range: slot n <-> ray index first + (n / block) * stride * block + n % block
    int64_t first, count, stride, block;
    // work order: tickets are handed out so that a warp starts on a 2-D tile of GB_TILE_R x GB_TILE_C neighbouring
    // rays (rows x columns of the image / r x theta of the polar plane) instead of 32 consecutive rows; 0 = ticket order
    int64_t tile_h;
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
    double* o_closest; // GB200_GEOMETRY_TARGET_POINT: closest approach to the target
    int32_t npf;
    int32_t pf[GB_MAX_PF];
    double* o_img[GB_MAX_PF];
    // line profile: per-ray (g, f) written for the binning kernel
    double* o_g;
    double* o_f;
    // fused line-profile histogram (lp_bins != nullptr): Buckets.Simple bins in HBM (read through L1), per-CTA 128-bit
    // fixed-point bins in shared memory, flushed into lp_acc[2 * nbins] = (low words | high words) at the end of the CTA
    const double* lp_bins;
    int32_t lp_nbins, lp_right_closed;
    double lp_scale;            // f is accumulated as round(f * lp_scale) in 128-bit integers
    unsigned long long* lp_acc;
    double min_re, max_re;
    int32_t emis_kind, emis_n;
    double emis_index;
    const double* emis_r;
    const double* emis_eps;
    // tabulated cross-section of GB200_GEOMETRY_THICK_TABLE (device pointers, rho ascending)
    int32_t cs_n;
    const double* cs_rho;
    const double* cs_h;
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
#ifdef __CUDA_ARCH__
#define GB_TAB(i) gb_tab[i]
#else
#define GB_TAB(i) gb_tab_host[i]
#endif
#define GB_A21_V 0.161
#define GB_A21 GB_TAB(0)
#define GB_A31_V -0.008480655492356989
#define GB_A31 GB_TAB(1)
#define GB_A32_V 0.335480655492357
#define GB_A32 GB_TAB(2)
#define GB_A41_V 2.8971530571054935
#define GB_A41 GB_TAB(3)
#define GB_A42_V -6.359448489975075
#define GB_A42 GB_TAB(4)
#define GB_A43_V 4.3622954328695815
#define GB_A43 GB_TAB(5)
#define GB_A51_V 5.325864828439257
#define GB_A51 GB_TAB(6)
#define GB_A52_V -11.748883564062828
#define GB_A52 GB_TAB(7)
#define GB_A53_V 7.4955393428898365
#define GB_A53 GB_TAB(8)
#define GB_A54_V -0.09249506636175525
#define GB_A54 GB_TAB(9)
#define GB_A61_V 5.86145544294642
#define GB_A61 GB_TAB(10)
#define GB_A62_V -12.92096931784711
#define GB_A62 GB_TAB(11)
#define GB_A63_V 8.159367898576159
#define GB_A63 GB_TAB(12)
#define GB_A64_V -0.071584973281401
#define GB_A64 GB_TAB(13)
#define GB_A65_V -0.028269050394068383
#define GB_A65 GB_TAB(14)
#define GB_A71_V 0.09646076681806523
#define GB_A71 GB_TAB(15)
#define GB_A72_V 0.01
#define GB_A72 GB_TAB(16)
#define GB_A73_V 0.4798896504144996
#define GB_A73 GB_TAB(17)
#define GB_A74_V 1.379008574103742
#define GB_A74 GB_TAB(18)
#define GB_A75_V -3.290069515436081
#define GB_A75 GB_TAB(19)
#define GB_A76_V 2.324710524099774
#define GB_A76 GB_TAB(20)
#define GB_BT1_V -0.00178001105222577714
#define GB_BT1 GB_TAB(21)
#define GB_BT2_V -0.0008164344596567469
#define GB_BT2 GB_TAB(22)
#define GB_BT3_V 0.007880878010261995
#define GB_BT3 GB_TAB(23)
#define GB_BT4_V -0.1447110071732629
#define GB_BT4 GB_TAB(24)
#define GB_BT5_V 0.5823571654525552
#define GB_BT5 GB_TAB(25)
#define GB_BT6_V -0.45808210592918697
#define GB_BT6 GB_TAB(26)
#define GB_BT7_V 0.015151515151515152
#define GB_BT7 GB_TAB(27)
// dense output: b_1 = Th (1 + Th (R12 + Th (R13 + Th R14))), b_j = Th^2 (Rj2 + Th (Rj3 + Th Rj4))
#define GB_R12_V -2.763706197274826
#define GB_R12 GB_TAB(28)
#define GB_R13_V 2.9132554618219126
#define GB_R13 GB_TAB(29)
#define GB_R14_V -1.0530884977290216
#define GB_R14 GB_TAB(30)
#define GB_R22_V 0.13169999999999998
#define GB_R22 GB_TAB(31)
#define GB_R23_V -0.2234
#define GB_R23 GB_TAB(32)
#define GB_R24_V 0.1017
#define GB_R24 GB_TAB(33)
#define GB_R32_V 3.9302962368947516
#define GB_R32 GB_TAB(34)
#define GB_R33_V -5.941033872131505
#define GB_R33 GB_TAB(35)
#define GB_R34_V 2.490627285651253
#define GB_R34 GB_TAB(36)
#define GB_R42_V -12.411077166933676
#define GB_R42 GB_TAB(37)
#define GB_R43_V 30.33818863028232
#define GB_R43 GB_TAB(38)
#define GB_R44_V -16.548102889244902
#define GB_R44 GB_TAB(39)
#define GB_R52_V 37.50931341651104
#define GB_R52 GB_TAB(40)
#define GB_R53_V -88.1789048947664
#define GB_R53 GB_TAB(41)
#define GB_R54_V 47.37952196281928
#define GB_R54 GB_TAB(42)
#define GB_R62_V -27.896526289197286
#define GB_R62 GB_TAB(43)
#define GB_R63_V 65.09189467479366
#define GB_R63 GB_TAB(44)
#define GB_R64_V -34.87065786149661
#define GB_R64 GB_TAB(45)
#define GB_R72_V 1.5
#define GB_R72 GB_TAB(46)
#define GB_R73_V -4.0
#define GB_R73 GB_TAB(47)
#define GB_R74_V 2.5
#define GB_R74 GB_TAB(48)


// Tableau constants live in the constant bank so FP64 instructions read them as c[bank][offset] operands
// (literal doubles are otherwise materialised with two UMOVs per use: 17% of all executed instructions in v1).
#ifdef __CUDACC__
__device__ __constant__ double gb_tab[] = {
    0.161, // 0: GB_A21
    -0.008480655492356989, // 1: GB_A31
    0.335480655492357, // 2: GB_A32
    2.8971530571054935, // 3: GB_A41
    -6.359448489975075, // 4: GB_A42
    4.3622954328695815, // 5: GB_A43
    5.325864828439257, // 6: GB_A51
    -11.748883564062828, // 7: GB_A52
    7.4955393428898365, // 8: GB_A53
    -0.09249506636175525, // 9: GB_A54
    5.86145544294642, // 10: GB_A61
    -12.92096931784711, // 11: GB_A62
    8.159367898576159, // 12: GB_A63
    -0.071584973281401, // 13: GB_A64
    -0.028269050394068383, // 14: GB_A65
    0.09646076681806523, // 15: GB_A71
    0.01, // 16: GB_A72
    0.4798896504144996, // 17: GB_A73
    1.379008574103742, // 18: GB_A74
    -3.290069515436081, // 19: GB_A75
    2.324710524099774, // 20: GB_A76
    -0.00178001105222577714, // 21: GB_BT1
    -0.0008164344596567469, // 22: GB_BT2
    0.007880878010261995, // 23: GB_BT3
    -0.1447110071732629, // 24: GB_BT4
    0.5823571654525552, // 25: GB_BT5
    -0.45808210592918697, // 26: GB_BT6
    0.015151515151515152, // 27: GB_BT7
    -2.763706197274826, // 28: GB_R12
    2.9132554618219126, // 29: GB_R13
    -1.0530884977290216, // 30: GB_R14
    0.13169999999999998, // 31: GB_R22
    -0.2234, // 32: GB_R23
    0.1017, // 33: GB_R24
    3.9302962368947516, // 34: GB_R32
    -5.941033872131505, // 35: GB_R33
    2.490627285651253, // 36: GB_R34
    -12.411077166933676, // 37: GB_R42
    30.33818863028232, // 38: GB_R43
    -16.548102889244902, // 39: GB_R44
    37.50931341651104, // 40: GB_R52
    -88.1789048947664, // 41: GB_R53
    47.37952196281928, // 42: GB_R54
    -27.896526289197286, // 43: GB_R62
    65.09189467479366, // 44: GB_R63
    -34.87065786149661, // 45: GB_R64
    1.5, // 46: GB_R72
    -4.0, // 47: GB_R73
    2.5, // 48: GB_R74
    0.63661977236758134308, // 49: GB_SC_2OPI
    1.57079632673412561417e+00, // 50: GB_SC_PIO2_1
    6.07710050650619224932e-11, // 51: GB_SC_PIO2_2
    2.02226624879595063154e-21, // 52: GB_SC_PIO2_3
    1.58969099521155010221e-10, // 53: GB_SC_S6
    -2.50507602534068634195e-08, // 54: GB_SC_S5
    2.75573137070700676789e-06, // 55: GB_SC_S4
    -1.98412698298579493134e-04, // 56: GB_SC_S3
    8.33333333332248946124e-03, // 57: GB_SC_S2
    -1.66666666666666324348e-01, // 58: GB_SC_S1
    -1.13596475577881948265e-11, // 59: GB_SC_C6
    2.08757232129817482790e-09, // 60: GB_SC_C5
    -2.75573143513906633035e-07, // 61: GB_SC_C4
    2.48015872894767294178e-05, // 62: GB_SC_C3
    -1.38888888888741095749e-03, // 63: GB_SC_C2
    4.16666666666666019037e-02, // 64: GB_SC_C1
    6.93147180369123816490e-01, // 65: GB_LN2_HI
    1.90821492927058770002e-10, // 66: GB_LN2_LO
    6.666666666666735130e-01, // 67: GB_LG1
    3.999999999940941908e-01, // 68: GB_LG2
    2.857142874366239149e-01, // 69: GB_LG3
    2.222219843214978396e-01, // 70: GB_LG4
    1.818357216161805012e-01, // 71: GB_LG5
    1.531383769920937332e-01, // 72: GB_LG6
    1.479819860511658591e-01, // 73: GB_LG7
    1.44269504088896338700e+00, // 74: GB_INVLN2
    2.08767569878681e-09, // 75: GB_EX12
    2.505210838544172e-08, // 76: GB_EX11
    2.755731922398589e-07, // 77: GB_EX10
    2.7557319223985893e-06, // 78: GB_EX9
    2.48015873015873e-05, // 79: GB_EX8
    0.0001984126984126984, // 80: GB_EX7
    0.001388888888888889, // 81: GB_EX6
    0.008333333333333333, // 82: GB_EX5
    0.041666666666666664, // 83: GB_EX4
    0.16666666666666666, // 84: GB_EX3
    0.5, // 85: GB_EX2
};
#endif
#define GB_SC_2OPI GB_TAB(49)
#define GB_SC_PIO2_1 GB_TAB(50)
#define GB_SC_PIO2_2 GB_TAB(51)
#define GB_SC_PIO2_3 GB_TAB(52)
#define GB_SC_S6 GB_TAB(53)
#define GB_SC_S5 GB_TAB(54)
#define GB_SC_S4 GB_TAB(55)
#define GB_SC_S3 GB_TAB(56)
#define GB_SC_S2 GB_TAB(57)
#define GB_SC_S1 GB_TAB(58)
#define GB_SC_C6 GB_TAB(59)
#define GB_SC_C5 GB_TAB(60)
#define GB_SC_C4 GB_TAB(61)
#define GB_SC_C3 GB_TAB(62)
#define GB_SC_C2 GB_TAB(63)
#define GB_SC_C1 GB_TAB(64)
#define GB_LN2_HI GB_TAB(65)
#define GB_LN2_LO GB_TAB(66)
#define GB_LG1 GB_TAB(67)
#define GB_LG2 GB_TAB(68)
#define GB_LG3 GB_TAB(69)
#define GB_LG4 GB_TAB(70)
#define GB_LG5 GB_TAB(71)
#define GB_LG6 GB_TAB(72)
#define GB_LG7 GB_TAB(73)
#define GB_INVLN2 GB_TAB(74)
#define GB_EX12 GB_TAB(75)
#define GB_EX11 GB_TAB(76)
#define GB_EX10 GB_TAB(77)
#define GB_EX9 GB_TAB(78)
#define GB_EX8 GB_TAB(79)
#define GB_EX7 GB_TAB(80)
#define GB_EX6 GB_TAB(81)
#define GB_EX5 GB_TAB(82)
#define GB_EX4 GB_TAB(83)
#define GB_EX3 GB_TAB(84)
#define GB_EX2 GB_TAB(85)
static const double gb_tab_host[] = {0.161, -0.008480655492356989, 0.335480655492357, 2.8971530571054935, -6.359448489975075, 4.3622954328695815, 5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525, 5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383, 0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774, -0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 0.015151515151515152, -2.763706197274826, 2.9132554618219126, -1.0530884977290216, 0.13169999999999998, -0.2234, 0.1017, 3.9302962368947516, -5.941033872131505, 2.490627285651253, -12.411077166933676, 30.33818863028232, -16.548102889244902, 37.50931341651104, -88.1789048947664, 47.37952196281928, -27.896526289197286, 65.09189467479366, -34.87065786149661, 1.5, -4.0, 2.5, 0.63661977236758134308, 1.57079632673412561417e+00, 6.07710050650619224932e-11, 2.02226624879595063154e-21, 1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01, -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, 2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02, 6.93147180369123816490e-01, 1.90821492927058770002e-10, 6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01, 2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01, 1.479819860511658591e-01, 1.44269504088896338700e+00, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06, 2.48015873015873e-05, 0.0001984126984126984, 0.001388888888888889, 0.008333333333333333, 0.041666666666666664, 0.16666666666666666, 0.5};

// ---------------------------------------------------------------- metrics
// Kerr: components and Jacobian from w = 2Mr/Sigma (see DESIGN.md for the derivation).
template <class S>
GB_HD inline void kerr_metric_jacobian(double M, double a, S r, S s, S c, S g[5], S dr[5], S dth[5]) {
    const double a2 = a * a;
    const S r2 = r * r, s2 = s * s, c2 = c * c;
    const S sin2 = 2.0 * s * c;
    const S Sig = r2 + a2 * c2;
    const S Del = r2 - 2.0 * M * r + a2;
    const S iSig = gb_rcp(Sig), iDel = gb_rcp(Del);
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

// Morris-Thorne wormhole, src/metrics/morris-thorne-ad.jl:4-15; mp = (b).  The reference's phi-phi component carries a
// single power of sin(theta) (:11); restated as written there.
template <class S>
GB_HD inline void morris_thorne_metric_jacobian(const double* mp, S r, S s, S c, S g[5], S dr[5], S dth[5]) {
    const S rho2 = r * r + mp[0] * mp[0];
    const S zero = 0.0 * r;
    g[0] = zero - 1.0; g[1] = zero + 1.0; g[2] = rho2; g[3] = rho2 * s; g[4] = zero;
    dr[0] = zero; dr[1] = zero; dr[2] = 2.0 * r; dr[3] = 2.0 * r * s; dr[4] = zero;
    dth[0] = zero; dth[1] = zero; dth[2] = zero; dth[3] = rho2 * c; dth[4] = zero;
}

// components + Jacobian of metric `kind` with parameters mp[] (s = sin theta, c = cos theta)
template <class S>
GB_HD inline void metric_jacobian_kind(int kind, const double* mp, S r, S s, S c, S g[5], S dr[5], S dth[5]) {
    switch (kind) {
    case GB200_METRIC_KERR: kerr_metric_jacobian<S>(mp[0], mp[1], r, s, c, g, dr, dth); break;
    case GB200_METRIC_JOHANNSEN_PSALTIS: jp_metric_jacobian<S>(mp[0], mp[1], mp[2], r, s, c, g, dr, dth); break;
    case GB200_METRIC_JOHANNSEN: johannsen_metric_jacobian<S>(mp, r, s, c, g, dr, dth); break;
    case GB200_METRIC_BUMBLEBEE: bumblebee_metric_jacobian<S>(mp, r, s, c, g, dr, dth); break;
    case GB200_METRIC_MORRIS_THORNE: morris_thorne_metric_jacobian<S>(mp, r, s, c, g, dr, dth); break;
    case GB200_METRIC_DILATON_AXION: dilaton_axion_metric_jacobian<S>(mp, r, s, c, g, dr, dth); break;
    default: kerr_newman_metric_jacobian<S>(mp, r, s, c, g, dr, dth); break;
    }
}
template <int METRIC>
GB_HD inline void metric_jacobian(const GbParams& P, double r, double s, double c, double g[5], double dr[5], double dth[5]) {
    metric_jacobian_kind<double>(METRIC, P.mp, r, s, c, g, dr, dth); // METRIC is a compile-time constant: the switch folds
}
GB_HD inline void metric_jacobian_rt(const GbParams& P, double r, double s, double c, double g[5], double dr[5], double dth[5]) {
    metric_jacobian_kind<double>(P.metric_kind, P.mp, r, s, c, g, dr, dth);
}

// a^mu = -g^{mu m} ( gdot_{mk} v^k - 1/2 S_m ),  gdot = v^r d_r g + v^th d_th g,  S_m = d_m g_{kl} v^k v^l
// (the contraction of auto-diff.jl:115-141 with d_t = d_phi = 0 written out)
GB_HD inline void geodesic_accel(const double g[5], const double dr[5], const double dth[5],
                                 double vt, double vr, double vth, double vph, double acc[4], double* gi_out = nullptr) {
    const double D = g[0] * g[3] - g[4] * g[4];
    const double iD = gb_rcp(D);
    const double gitt = g[3] * iD, giphph = g[0] * iD, gitph = -g[4] * iD;
    const double girr = gb_rcp(g[1]), githth = gb_rcp(g[2]);
    if (gi_out) { gi_out[0] = gitt; gi_out[1] = girr; gi_out[2] = githth; gi_out[3] = giphph; gi_out[4] = gitph; }
    const double d0 = vr * dr[0] + vth * dth[0];
    const double d3 = vr * dr[3] + vth * dth[3];
    const double d4 = vr * dr[4] + vth * dth[4];
    const double Pt = d0 * vt + d4 * vph;
    const double Pp = d4 * vt + d3 * vph;
    const double vtt = vt * vt, vrr = vr * vr, vthth = vth * vth, vpp = vph * vph, vtp2 = 2.0 * vt * vph;
    // S_r / 2 - gdot_rr v^r and S_th / 2 - gdot_thth v^th with the (v^r)^2, (v^th)^2 terms merged (see kerr_rhs_accel_sq)
    const double X = vr * vth;
    const double Br = dr[0] * vtt - dr[1] * vrr + dr[2] * vthth + dr[3] * vpp + dr[4] * vtp2;
    const double Bt = dth[0] * vtt + dth[1] * vrr - dth[2] * vthth + dth[3] * vpp + dth[4] * vtp2;
    acc[0] = fma(-gitt, Pt, -gitph * Pp);
    acc[1] = girr * fma(-dth[1], X, 0.5 * Br);
    acc[2] = githth * fma(-dr[2], X, 0.5 * Bt);
    acc[3] = fma(-gitph, Pt, -giphph * Pp);
}

#ifndef GB_OPT_SIGNFLIP
#define GB_OPT_SIGNFLIP 1 /* quadrant signs of sincos through the integer pipe */
#endif
#ifndef GB_OPT_KERR_NOD12
#define GB_OPT_KERR_NOD12 1 /* Kerr a^r, a^theta without gdot_rr, gdot_thth */
#endif
#ifndef GB_OPT_MAGIC
#define GB_OPT_MAGIC 1 /* nearest-integer through the 1.5 * 2^52 shift instead of rint + double->int conversion */
#endif
#ifndef GB_OPT_CW2
#define GB_OPT_CW2 1 /* two-term Cody-Waite reduction (the third term is 2e-21 per quadrant) */
#endif
#ifndef GB_OPT_KERR1R
#define GB_OPT_KERR1R 1 /* Kerr RHS with a single reciprocal 1/(Sigma Delta sin^2) */
#endif
#ifndef GB_OPT_KERR_SQ
#define GB_OPT_KERR_SQ 1 /* Kerr RHS from sin^2, cos^2, sin 2theta of the reduced argument (signs of sin, cos only where needed) */
#endif
#ifdef __CUDACC__
// Branch-free sincos for the polar angle (|x| << 2^20 rad always holds: theta is bounded by the number of polar
// passages): Cody-Waite reduction by pi/2 in three parts, fdlibm __kernel_sin/__kernel_cos minimax polynomials on
// [-pi/4, pi/4] (< 1 ulp), quadrant fix-up with selects.  The CUDA library sincos() carries a Payne-Hanek slow
// path behind a divergent branch; this has none.
// x with its sign flipped when bit 31 of `signbit31` is set (integer pipe; a DADD negate + select costs FP64 issue slots)
GB_D double gb_flip_sign(double x, int signbit31) {
    return __hiloint2double(__double2hiint(x) ^ (signbit31 & (int)0x80000000), __double2loint(x));
}
// Reduced form: x = n pi/2 + rr, returns sn = sin(rr), cs = cos(rr) and the quadrant n.
struct GbSinCos { double sn, cs; int n; };
GB_D GbSinCos gb_sincos_reduced(double x) {
#if GB_OPT_MAGIC
    // round-to-nearest through the 1.5 * 2^52 shift: the integer lands in the low mantissa word (two's complement) and the
    // rounded value comes back with one subtraction; rint + double->int conversions are slow-rate instructions
    const double tq = fma(x, GB_SC_2OPI, 6755399441055744.0);
    const double q = tq - 6755399441055744.0;
#else
    const double q = rint(x * GB_SC_2OPI);
#endif
    double rr = fma(-q, GB_SC_PIO2_1, x);
    rr = fma(-q, GB_SC_PIO2_2, rr);
#if !GB_OPT_CW2
    rr = fma(-q, GB_SC_PIO2_3, rr); // 2.0e-21 |q|: below 1e-18 for the |theta| < 500 a geodesic can reach
#endif
    const double z = rr * rr;
    double ps = fma(z, GB_SC_S6, GB_SC_S5);
    ps = fma(z, ps, GB_SC_S4);
    ps = fma(z, ps, GB_SC_S3);
    ps = fma(z, ps, GB_SC_S2);
    ps = fma(z, ps, GB_SC_S1);
    const double sn = fma(rr * z, ps, rr);
    double pc = fma(z, GB_SC_C6, GB_SC_C5);
    pc = fma(z, pc, GB_SC_C4);
    pc = fma(z, pc, GB_SC_C3);
    pc = fma(z, pc, GB_SC_C2);
    pc = fma(z, pc, GB_SC_C1);
    const double cs = fma(z * z, pc, fma(z, -0.5, 1.0));
    GbSinCos o;
    o.sn = sn; o.cs = cs;
#if GB_OPT_MAGIC
    o.n = __double2loint(tq);
#else
    o.n = __double2int_rn(q);
#endif
    return o;
}
GB_D void gb_sincos(double x, double* sp, double* cp) {
    const GbSinCos o = gb_sincos_reduced(x);
    const bool swap = (o.n & 1) != 0;
    const double so = swap ? o.cs : o.sn;
    const double co = swap ? o.sn : o.cs;
#if GB_OPT_SIGNFLIP
    *sp = gb_flip_sign(so, o.n << 30);       // n & 2
    *cp = gb_flip_sign(co, (o.n + 1) << 30); // (n + 1) & 2
#else
    *sp = (o.n & 2) ? -so : so;
    *cp = ((o.n + 1) & 2) ? -co : co;
#endif
}

// Branch-free natural logarithm for the step controller (argument: the error estimate, a positive normal double or 0).
// fdlibm/musl __log: x = 2^k m with m in [sqrt(1/2), sqrt(2)), f = m - 1, s = f/(2+f), log m = f - f^2/2 + s (f^2/2 + R(s^2));
// < 1 ulp.  x = 0 returns a large negative number (about -745) instead of -inf, which the controller clamps the same way.
// The library log() carries denormal/inf/nan branches and ~20 literal coefficients that each cost two moves per use.
GB_D double gb_log_pos(double x) {
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    hx += 0x3ff00000 - 0x3fe6a09e;
    const int k = (hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffff) + 0x3fe6a09e;
    const double m = __hiloint2double(hx, lx);
    const double f = m - 1.0;
    const double hfsq = 0.5 * f * f;
    const double sq = f * gb_rcp(2.0 + f);
    const double z = sq * sq, w = z * z;
    const double t1 = w * fma(w, fma(w, GB_LG6, GB_LG4), GB_LG2);
    const double t2 = z * fma(w, fma(w, fma(w, GB_LG7, GB_LG5), GB_LG3), GB_LG1);
    const double R = t2 + t1;
    const double dk = (double)k;
    return fma(dk, GB_LN2_HI, fma(sq, hfsq + R, dk * GB_LN2_LO) - hfsq + f);
}
// Branch-free exp for |x| <= 8 (the controller clamps its argument first): x = k ln2 + rr, degree-12 Taylor polynomial on
// |rr| <= ln2/2 (truncation 1.7e-16 relative), scaled by 2^k through the exponent field.
GB_D double gb_exp_small(double x) {
#if GB_OPT_MAGIC
    const double tk = fma(x, GB_INVLN2, 6755399441055744.0);
    const double kf = tk - 6755399441055744.0;
#else
    const double kf = rint(x * GB_INVLN2);
#endif
    double rr = fma(-kf, GB_LN2_HI, x);
    rr = fma(-kf, GB_LN2_LO, rr);
    // Estrin-style split into even/odd halves keeps the dependency chain short
    const double r2 = rr * rr;
    double pe = fma(r2, GB_EX12, GB_EX10);
    double po = fma(r2, GB_EX11, GB_EX9);
    pe = fma(r2, pe, GB_EX8); po = fma(r2, po, GB_EX7);
    pe = fma(r2, pe, GB_EX6); po = fma(r2, po, GB_EX5);
    pe = fma(r2, pe, GB_EX4); po = fma(r2, po, GB_EX3);
    pe = fma(r2, pe, GB_EX2);
    // exp(rr) = 1 + rr + r2 * (pe + rr * po)
    const double p = fma(r2, fma(rr, po, pe), rr) + 1.0;
#if GB_OPT_MAGIC
    const int k = __double2loint(tk);
#else
    const int k = __double2int_rn(kf);
#endif
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// The controller's own log and exp (GB_OPT_CTRL_LO): the error estimate they act on carries a relative 2^-20 from its scales
// (GB_OPT_NORMRCP), and the reference evaluates this power at Float32 accuracy (FastPower), so 1e-7 is all that is asked of
// them: the reciprocal seed alone, four instead of seven terms of the log series (truncation 4e-9), a degree-7 instead of a
// degree-12 exponential (5e-9).  Twelve dependent FP64 instructions fewer at the serial end of every step attempt.
// 39.46 -> 39.01 ms on C2, 50.15 -> 49.67 ms on the Johannsen-Psaltis render (profiles/r02_tuning_log.md).
#ifndef GB_OPT_CTRL_LO
#define GB_OPT_CTRL_LO 1
#endif
GB_D double gb_log_pos_lo(double x) {
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    hx += 0x3ff00000 - 0x3fe6a09e;
    const int k = (hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffff) + 0x3fe6a09e;
    const double m = __hiloint2double(hx, lx);
    const double f = m - 1.0;
    const double hfsq = 0.5 * f * f;
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(2.0 + f));
    const double sq = f * y;
    const double z = sq * sq;
    const double R = z * fma(z, fma(z, fma(z, GB_LG4, GB_LG3), GB_LG2), GB_LG1);
    const double dk = (double)k;
    return fma(dk, GB_LN2_HI, fma(sq, hfsq + R, dk * GB_LN2_LO) - hfsq + f);
}
GB_D double gb_exp_small_lo(double x) {
    const double tk = fma(x, GB_INVLN2, 6755399441055744.0);
    const double kf = tk - 6755399441055744.0;
    double rr = fma(-kf, GB_LN2_HI, x);
    rr = fma(-kf, GB_LN2_LO, rr);
    const double r2 = rr * rr;
    const double pe = fma(r2, fma(r2, GB_EX6, GB_EX4), GB_EX2);
    const double po = fma(r2, fma(r2, GB_EX7, GB_EX5), GB_EX3);
    const double p = fma(r2, fma(rr, po, pe), rr) + 1.0;
    const int k = __double2loint(tk);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// Kerr right-hand side with everything folded: two reciprocals (1/(Sigma Delta) and 1/sin^2) and the identity
// g_tt g_phph - g_tph^2 = -Delta sin^2, so  g^tt = -B/Delta, g^tph = -a w/Delta, g^phph = (1 - w)/(Delta sin^2).
GB_D void kerr_rhs_accel_sq(double M, double a, double a2, double twoM, double r, double s2, double c2, double sin2, double vt, double vr, double vth, double vph, double acc[4]);
GB_D void kerr_rhs_accel(double M, double a, double r, double s, double c, double vt, double vr, double vth, double vph, double acc[4]) {
    kerr_rhs_accel_sq(M, a, a * a, 2.0 * M, r, s * s, c * c, 2.0 * s * c, vt, vr, vth, vph, acc);
}
#ifndef GB_OPT_KERR_LAG
#define GB_OPT_KERR_LAG 1 /* Kerr accelerations from the Euler-Lagrange form with A = tdot - a sin^2 phdot (v23: 41.13 -> 40.56 ms on C2) */
#endif
#if GB_OPT_KERR_LAG
// The same accelerations from 2L = -tdot^2 + (r^2 + a^2) s^2 phdot^2 + w A^2 + (Sigma / Delta) rdot^2 + Sigma thdot^2,
// w = 2 M r / Sigma, A = tdot - a s^2 phdot: p_t = -tdot + w A and p_phi = s^2 ((r^2 + a^2) phdot - a w A) are conserved,
// so tddot = d(w A)/dlambda =: Udot and phddot follows from p_phi; rddot and thddot are the Euler-Lagrange equations.
// 79 FP64 instructions against 83 and 15 three-register DFMAs against 21; a second reciprocal 1 / (r^2 + a^2) whose chain
// does not wait for the sine and cosine.
GB_D void kerr_rhs_accel_sq(double M, double a, double a2, double twoM, double r, double s2, double c2, double sin2, double vt, double vr, double vth, double vph, double acc[4]) {
    const double r2 = r * r;
    const double rho2 = r2 + a2;
    const double irho2 = gb_rcp(rho2);
    const double Sig = fma(a2, c2, r2);
    const double Del = fma(r, r - twoM, a2);
    const double Ds = Del * s2;
    const double R = gb_rcp(Sig * Ds);
    const double iSig = R * Ds, iDel_is2 = R * Sig, iDel = iDel_is2 * s2, is2 = iDel_is2 * Del;
    const double w = twoM * r * iSig;
    const double hw_r = fma(-w, r, M) * iSig; // w_r / 2
    const double a2sin2 = a2 * sin2;
    const double wS = w * iSig;
    const double as2 = a * s2;
    const double A = fma(-as2, vph, vt), U = w * A;
    const double h = a2sin2 * vth;
    const double wdot = fma(2.0 * hw_r, vr, wS * h);
    const double rr = r * vr;
    double x = fma(rho2 * wdot, A, -(w * h) * U);
    x = fma(2.0 * ((w * as2) * rr), vph, x);
    const double Udot = x * iDel;
    const double pp = vph * vph, tt = vth * vth, rr2 = vr * vr, X = vr * vth;
    acc[0] = Udot;
    {
        const double q3 = fma(hw_r, A * A, r * fma(s2, pp, tt));
        const double c2_ = fma(r, iSig, -((r - M) * iDel));
        acc[1] = fma(a2sin2 * iSig, X, fma(-c2_, rr2, (Del * iSig) * q3));
    }
    {
        const double e2 = fma(-(2.0 * a) * U, vph, rho2 * pp);
        const double e5 = fma(U * A, iSig, tt - rr2 * iDel);
        acc[2] = iSig * fma(-(2.0 * r), X, (0.5 * sin2) * fma(a2, e5, e2));
    }
    {
        const double m = (sin2 * is2) * vth;
        acc[3] = fma(-m, vph, fma(-(2.0 * rr), vph, a * fma(m, U, Udot)) * irho2);
    }
}
#else
GB_D void kerr_rhs_accel_sq(double M, double a, double a2, double twoM, double r, double s2, double c2, double sin2, double vt, double vr, double vth, double vph, double acc[4]) {
    const double r2 = r * r;
    const double Sig = fma(a2, c2, r2);
    const double Del = fma(r, r - twoM, a2);
#if GB_OPT_KERR1R
    const double Ds = Del * s2;
    const double R = gb_rcp(Sig * Ds);
    const double iSig = R * Ds, iDel_is2 = R * Sig, iDel = iDel_is2 * s2;
    const double a2sin2 = a2 * sin2;
    const double w = twoM * r * iSig;
    const double w_r = 2.0 * (M - w * r) * iSig;
    const double w_t = w * iSig * a2sin2;
    const double Sig_t = -a2sin2;
#else
    const double R = gb_rcp(Sig * Del), is2 = gb_rcp(s2);
    const double iSig = R * Del, iDel = R * Sig;
    const double iDel_is2 = iDel * is2;
    const double w = 2.0 * M * r * iSig;
    const double w_r = 2.0 * (M - w * r) * iSig;
    const double w_t = w * a2 * sin2 * iSig;
    const double Sig_t = -a2 * sin2;
#endif
    const double as2 = a2 * s2;
    const double B = fma(as2, w, r2 + a2);
    const double q = fma(s2, w_t, sin2 * w);
    const double grr = Sig * iDel;
    // Jacobian (d_r, d_theta) of (tt, rr, thth, phph, tph)
    const double r0 = w_r, t0 = w_t;
    const double r1 = (2.0 * r - grr * 2.0 * (r - M)) * iDel, t1 = Sig_t * iDel;
    const double r2_ = 2.0 * r, t2 = Sig_t;
    const double r3 = s2 * fma(as2, w_r, 2.0 * r), t3 = fma(sin2, B, as2 * q);
    const double r4 = -a * s2 * w_r, t4 = -a * q;
    // inverse metric
    // -g^tt, -g^tph, g^phph: the signs are carried by operand modifiers below, never by a separate negation
    const double mgitt = B * iDel, mgitph = a * w * iDel, giphph = (1.0 - w) * iDel_is2;
    const double girr = Del * iSig, githth = iSig;
    const double d0 = fma(vr, r0, vth * t0), d3 = fma(vr, r3, vth * t3), d4 = fma(vr, r4, vth * t4);
    const double Pt = fma(d0, vt, d4 * vph), Pp = fma(d4, vt, d3 * vph);
    const double vtt = vt * vt, vrr = vr * vr, vthth = vth * vth, vpp = vph * vph, vtp2 = 2.0 * vt * vph;
    acc[0] = fma(mgitt, Pt, mgitph * Pp);
    acc[3] = fma(mgitph, Pt, -giphph * Pp);
#if GB_OPT_KERR_NOD12
    // a^r = g^rr (S_r / 2 - gdot_rr v^r) with gdot_rr v^r = d_r g_rr v^r v^r + d_th g_rr v^r v^th: the d_r g_rr (v^r)^2 term
    // flips its sign inside the sum and only the cross term -d_th g_rr v^r v^th remains (same for a^theta), which
    // saves forming gdot_rr and gdot_thth
    const double X = vr * vth;
    const double Br = fma(r0, vtt, fma(-r1, vrr, fma(r2_, vthth, fma(r3, vpp, r4 * vtp2))));
    const double Bt = fma(t0, vtt, fma(t1, vrr, fma(-t2, vthth, fma(t3, vpp, t4 * vtp2))));
    acc[1] = girr * fma(-t1, X, 0.5 * Br);
    acc[2] = githth * fma(-r2_, X, 0.5 * Bt);
#else
    const double d1 = fma(vr, r1, vth * t1), d2 = fma(vr, r2_, vth * t2);
    const double Sr = fma(r0, vtt, fma(r1, vrr, fma(r2_, vthth, fma(r3, vpp, r4 * vtp2))));
    const double St = fma(t0, vtt, fma(t1, vrr, fma(t2, vthth, fma(t3, vpp, t4 * vtp2))));
    acc[1] = girr * fma(-d1, vr, 0.5 * Sr);
    acc[2] = githth * fma(-d2, vth, 0.5 * St);
#endif
}
#endif

// Lorentz force on a charged test particle in the Kerr-Newman field, q/mu F^mu_kappa v^kappa with
// F = g^-1 (dA - dA') (faraday_tensor, src/tracing/utility.jl:89-99; geodesic_ode_problem(::KerrNewmanMetric),
// src/metrics/kerr-newman-ad.jl:66-102) and A = (r Q / Sigma) (1, 0, 0, -a sin^2 theta) (:29-33), derivatives in closed form.
GB_HD inline void kerr_newman_lorentz(const double* mp, double r, double s, double c, const double gi[5],
                                      double vt, double vr, double vth, double vph, double acc[4]) {
    const double a = mp[1], Q = mp[2], q = mp[3];
    const double a2 = a * a, s2 = s * s, sc2 = 2.0 * s * c;
    const double Sig = fma(a2, c * c, r * r);
    const double iS = gb_rcp(Sig);
    const double At = Q * r * iS;
    const double At_r = Q * (Sig - 2.0 * r * r) * iS * iS; // d_r (r / Sigma) = (Sigma - 2 r^2) / Sigma^2
    const double At_t = At * a2 * sc2 * iS;                 // d_theta (1 / Sigma) = 2 a^2 sin cos / Sigma^2
    const double Ap_r = -a * s2 * At_r;
    const double Ap_t = -a * fma(sc2, At, s2 * At_t);
    // w_sigma = (d_kappa A_sigma - d_sigma A_kappa) v^kappa
    const double wt = fma(At_r, vr, At_t * vth), wp = fma(Ap_r, vr, Ap_t * vth);
    const double wr = -fma(At_r, vt, Ap_r * vph), wth = -fma(At_t, vt, Ap_t * vph);
    acc[0] = fma(q, fma(gi[0], wt, gi[4] * wp), acc[0]);
    acc[1] = fma(q * gi[1], wr, acc[1]);
    acc[2] = fma(q * gi[2], wth, acc[2]);
    acc[3] = fma(q, fma(gi[4], wt, gi[3] * wp), acc[3]);
}

// The right-hand side, inlined at each of the six stages.  (Measured alternatives, profiles/r01_tuning_log.md: one
// out-of-line copy removes instruction-fetch stalls but pays ~25% more instructions in call marshalling; with the
// CTA-synchronous stepping of gb200_trace.cu the inlined form is the faster one.)
#ifndef GB_OPT_JP_LAG
#define GB_OPT_JP_LAG 1 /* Johannsen-Psaltis accelerations from the Euler-Lagrange form (below) instead of the generated Jacobian + contraction */
#endif
// Johannsen-Psaltis accelerations in closed form.  With h = eps3 M^3 r / Sigma^2, k = 1 + h, w = 2 M r / Sigma, Z = k w,
// A = tdot - a s^2 phdot, B = tdot + a s^2 phdot the Lagrangian of johannsen-psaltis-ad.jl:4-26 is
//   2L = -tdot^2 + (r^2 + a^2) s^2 phdot^2 + Z A^2 - h A B + G rdot^2 + Sigma thdot^2,   G = Sigma k / Dh,  Dh = Delta + h a^2 s^2,
// and g_tt g_phph - g_tph^2 = -k s^2 Dh, so g^tt = -Phi / (k Dh), g^tph = -a w / Dh, g^phph = (1 - w) / (s^2 Dh) with
// Phi = r^2 + a^2 + a^2 s^2 (h + Z).  tddot, phddot: the 2 x 2 solve of d/dlambda (g_t. v) = d/dlambda (g_ph. v) = 0 with the
// total derivatives hdot, Zdot; rddot, thddot: the Euler-Lagrange equations with d ln G.  Two reciprocals (1 / Sigma, which
// h needs, and 1 / (k Dh s^2)) instead of five, 114 FP64 instructions instead of 156 with the generated Jacobian
// (tools/sass_count.py), equal to the oracle's dual-number right-hand side to 2e-13 (tests/test_gpu_parity.py).
GB_D void jp_rhs_accel_sq(double M, double a, double a2, double twoM, double e, double r, double s2, double c2, double sin2,
                          double vt, double vr, double vth, double vph, double acc[4]) {
    const double r2 = r * r;
    const double Sig = fma(a2, c2, r2);
    const double Del = fma(r, r - twoM, a2);
    const double rho2 = r2 + a2;
    const double a2s2 = a2 * s2;
    const double iSig = gb_rcp(Sig);
    const double iSig2 = iSig * iSig;
    const double q = r2 * iSig;
    const double mS = twoM * iSig, w = mS * r, w_r = mS * fma(-2.0, q, 1.0);
    const double eS = e * iSig2, h = eS * r, h_r = eS * fma(-4.0, q, 1.0);
    const double k = 1.0 + h, Z = k * w;
    const double Delh = fma(h, a2s2, Del);
    const double kD = k * Delh;
    const double R2 = gb_rcp(kD * s2);
    const double R2s = R2 * s2, ik = R2s * Delh, iDelh = R2s * k, is2 = R2 * kD;
    const double a2sin2 = a2 * sin2, T = a2sin2 * iSig, hT = h * T, w_t = w * T;
    const double Z_r = fma(h_r, w, k * w_r), Z_t = w_t * fma(3.0, h, 1.0);
    const double x = (a * s2) * vph, A = vt - x, AB = fma(vt, vt, -(x * x)), AA = A * A;
    const double hTv = hT * vth;
    const double hdot = fma(h_r, vr, hTv + hTv), Zdot = fma(Z_r, vr, Z_t * vth);
    const double m = (sin2 * is2) * vth, s2dot = sin2 * vth;
    const double zv = Z * vph;
    const double Pt = fma(Zdot, A, fma(-hdot, vt, -((a * s2dot) * zv)));
    const double rr = r * vr;
    const double Phi2 = fma(a2s2, fma(2.0, h, Z), rho2);
    const double inner2 = fma(a2s2, hdot, rr + rr);
    const double Pps = fma(vph, fma(m, Phi2, inner2), -((a * fma(m, Z, Zdot)) * A)); // (d/dlambda g_ph.) v / s^2
    const double Phi = fma(a2s2, h + Z, rho2);
    const double aw = a * w;
    acc[0] = iDelh * fma(Phi * ik, Pt, aw * (s2 * Pps));
    acc[3] = iDelh * fma(aw, Pt, -((1.0 - w) * Pps));
    const double pp = vph * vph, tt = vth * vth, rr2 = vr * vr, X = vr * vth;
    const double tr = r + r;
    {
        const double dQr = fma(Z_r, AA, fma(-h_r, AB, tr * (s2 * pp)));
        const double lnGr = fma(-fma(h_r, a2s2, tr - twoM), iDelh, fma(h_r, ik, tr * iSig));
        const double hi2 = 2.0 * (h * iSig);
        const double lnGt = a2sin2 * fma(-fma(hi2, a2s2, h), iDelh, fma(hi2, ik, -iSig));
        const double iG = (iSig * ik) * Delh;
        acc[1] = fma(-lnGt, X, fma(-(0.5 * lnGr), rr2, iG * fma(r, tt, 0.5 * dQr)));
        const double G = (Sig * k) * iDelh;
        double dQt = (sin2 * pp) * fma(2.0 * h, a2s2, rho2);
        dQt = fma(Z_t, AA, dQt);
        dQt = fma(-2.0 * ((a * sin2) * zv), A, dQt);
        dQt = fma(-2.0 * hT, AB, dQt);
        acc[2] = (0.5 * iSig) * fma(-4.0 * rr, vth, fma(a2sin2, tt, fma(G * lnGt, rr2, dQt)));
    }
}

struct GbAcc { double a0, a1, a2, a3, s, c; };
#ifdef GB_RHS_CALL /* tuning variant: one out-of-line copy (smaller code, but ~25% more instructions for call marshalling) */
#define GB_RHS_ATTR __device__ __noinline__
#else
#define GB_RHS_ATTR __device__ __forceinline__
#endif
template <int METRIC>
GB_RHS_ATTR GbAcc rhs_eval(const GbParams& P, double r, double th, double vt, double vr, double vth, double vph) {
    GbAcc o;
    double acc[4];
#if GB_OPT_KERR_SQ
    if (METRIC == GB200_METRIC_KERR) {
        // squares and sin(2 theta) need no quadrant signs beyond the parity of n; the signed sin, cos are formed with
        // integer operations and are dead code at the stages whose caller ignores them
        const GbSinCos q = gb_sincos_reduced(th);
        const bool swap = (q.n & 1) != 0;
        const double sa = swap ? q.cs : q.sn, ca = swap ? q.sn : q.cs;
        o.s = gb_flip_sign(sa, q.n << 30);
        o.c = gb_flip_sign(ca, (q.n + 1) << 30);
        kerr_rhs_accel_sq(P.M, P.a, P.a2, P.twoM, r, sa * sa, ca * ca, gb_flip_sign(2.0 * sa * ca, q.n << 31), vt, vr, vth, vph, acc);
        o.a0 = acc[0]; o.a1 = acc[1]; o.a2 = acc[2]; o.a3 = acc[3];
        return o;
    }
#endif
#if GB_OPT_JP_LAG
    if (METRIC == GB200_METRIC_JOHANNSEN_PSALTIS) {
        const GbSinCos q = gb_sincos_reduced(th);
        const bool swap = (q.n & 1) != 0;
        const double sa = swap ? q.cs : q.sn, ca = swap ? q.sn : q.cs;
        o.s = gb_flip_sign(sa, q.n << 30);
        o.c = gb_flip_sign(ca, (q.n + 1) << 30);
        jp_rhs_accel_sq(P.M, P.a, P.a2, P.twoM, P.jp_e, r, sa * sa, ca * ca, gb_flip_sign(2.0 * sa * ca, q.n << 31), vt, vr, vth, vph, acc);
        o.a0 = acc[0]; o.a1 = acc[1]; o.a2 = acc[2]; o.a3 = acc[3];
        return o;
    }
#endif
    gb_sincos(th, &o.s, &o.c);
    if (METRIC == GB200_METRIC_KERR) {
        kerr_rhs_accel(P.M, P.a, r, o.s, o.c, vt, vr, vth, vph, acc);
    } else {
        double g[5], dr[5], dth[5];
        metric_jacobian<METRIC>(P, r, o.s, o.c, g, dr, dth);
        if (METRIC == GB200_METRIC_KERR_NEWMAN) {
            double gi[5];
            geodesic_accel(g, dr, dth, vt, vr, vth, vph, acc, gi);
            if (P.mp[3] != 0.0) kerr_newman_lorentz(P.mp, r, o.s, o.c, gi, vt, vr, vth, vph, acc);
        } else geodesic_accel(g, dr, dth, vt, vr, vth, vph, acc);
    }
    o.a0 = acc[0]; o.a1 = acc[1]; o.a2 = acc[2]; o.a3 = acc[3];
    return o;
}
// The seventh stage sits at the same abscissa as the sixth (c_6 = c_7 = 1): the two polar angles are both approximations of
// theta(t + dt) and differ by about ten times the local error, so the seventh stage's sine and cosine follow from the sixth
// stage's by the addition theorem with the Taylor series of sin(delta), cos(delta) to third order.  The neglected
// delta^4 / 24 is below 1e-30 at the default tolerance and stays seven orders below the tolerance for any tolerance up to
// 1e-2 (an attempt whose delta is larger than that is far outside its tolerance and is rejected whatever its stage values
// are).  9 FP64 instructions instead of the 22 of a full evaluation: 39.9 -> 39.2 ms on C2 (profiles/r02_tuning_log.md).
#ifndef GB_OPT_SC7
#define GB_OPT_SC7 1
#endif
template <int METRIC>
GB_D void rhs_accel_near(const GbParams& P, double r, double th, double vt, double vr, double vth, double vph,
                         double th_prev, double s_prev, double c_prev, double acc[4], double& s, double& c) {
    if (METRIC == GB200_METRIC_KERR || (METRIC == GB200_METRIC_JOHANNSEN_PSALTIS && GB_OPT_JP_LAG)) {
        const double d = th - th_prev;
        const double d6 = d * (-1.0 / 6.0);
        // sin(th) = s + d (c + d (-s/2 + d (-c/6))),  cos(th) = c + d (-s + d (-c/2 + d (s/6)))
        const double sn = fma(d, fma(d, fma(d6, c_prev, -0.5 * s_prev), c_prev), s_prev);
        const double cn = fma(d, fma(d, fma(-d6, s_prev, -0.5 * c_prev), -s_prev), c_prev);
        s = sn; c = cn;
        const double sc = sn * cn;
        if (METRIC == GB200_METRIC_KERR) kerr_rhs_accel_sq(P.M, P.a, P.a2, P.twoM, r, sn * sn, cn * cn, sc + sc, vt, vr, vth, vph, acc);
        else jp_rhs_accel_sq(P.M, P.a, P.a2, P.twoM, P.jp_e, r, sn * sn, cn * cn, sc + sc, vt, vr, vth, vph, acc);
        return;
    }
    const GbAcc o = rhs_eval<METRIC>(P, r, th, vt, vr, vth, vph);
    acc[0] = o.a0; acc[1] = o.a1; acc[2] = o.a2; acc[3] = o.a3; s = o.s; c = o.c;
}
template <int METRIC>
GB_D void rhs_accel(const GbParams& P, double r, double th, double vt, double vr, double vth, double vph,
                    double acc[4], double& s, double& c) {
    const GbAcc o = rhs_eval<METRIC>(P, r, th, vt, vr, vth, vph);
    acc[0] = o.a0; acc[1] = o.a1; acc[2] = o.a2; acc[3] = o.a3; s = o.s; c = o.c;
}
#endif

template <int METRIC>
GB_HD inline void metric_components_t(const GbParams& P, double r, double s, double c, double g[5]) {
    double dr[5], dth[5];
    metric_jacobian<METRIC>(P, r, s, c, g, dr, dth);
}
GB_HD inline void metric_components_rt(const GbParams& P, double r, double th, double g[5]) {
    double dr[5], dth[5];
    metric_jacobian_rt(P, r, sin(th), cos(th), g, dr, dth);
}

// constrain_time, auto-diff.jl:161-173
GB_HD inline double constrain_vt(const double g[5], double vr, double vth, double vph, double mu) {
    const double disc = -g[0] * g[1] * vr * vr - g[0] * g[2] * vth * vth - g[0] * mu * mu - (g[0] * g[3] - g[4] * g[4]) * vph * vph;
    return -(g[4] * vph + sqrt(disc)) / g[0];
}

// cross_section(rho) of a tabulated ThickDisc: linear interpolation, -1 (no disc) outside the table
GB_HD inline double cross_section_table(const double* xs, const double* ys, int n, double x) {
    if (n < 2 || !(x >= xs[0]) || !(x <= xs[n - 1])) return -1.0;
    int lo = 0, hi = n - 1; // last index with xs[idx] <= x, clamped to [0, n-2]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (xs[mid] <= x) lo = mid; else hi = mid; }
    const double w = (x - xs[lo]) / (xs[lo + 1] - xs[lo]);
    return (1.0 - w) * ys[lo] + w * ys[lo + 1];
}

// ---------------------------------------------------------------- disc conditions (distance_to_disc)
// `hgt` is the datum-plane height of this ray (P.gp0 unless the IC carries per-ray heights); unused otherwise.
template <int GEOM>
GB_HD inline double disc_condition(const GbParams& P, double r, double s, double c, double hgt) {
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
        return r * c - hgt;
    } else if (GEOM == GB200_GEOMETRY_THICK_TABLE) { // thick-disc.jl:57-63 with the closure's cross-section tabulated
        const double h = cross_section_table(P.cs_rho, P.cs_h, P.cs_n, r * fabs(s));
        if (h <= 0.0) return 1.0;
        return r * fabs(c) - h;
    }
    return 1.0;
}

// ---------------------------------------------------------------- ray indexing
#ifndef GB_TILE_R
#define GB_TILE_R 8
#endif
#ifndef GB_TILE_C
#define GB_TILE_C 4
#endif
GB_HD inline int64_t ray_index_of_slot(const GbParams& P, int64_t n) {
    return P.first + (n / P.block) * (P.stride * P.block) + n % P.block;
}
// ticket (work-queue order) -> slot (output order).  Within every strip of GB_TILE_C columns (= GB_TILE_C * tile_h
// consecutive slots) tickets walk 8 x 4 tiles; neighbouring rays take similar step counts and reach the disc at the
// same steps, which keeps a warp's event scans, terminations and refills in phase.
GB_HD inline int64_t slot_of_ticket(const GbParams& P, int64_t k) {
    if (P.tile_h == 0) return k;
    const int64_t strip = GB_TILE_C * P.tile_h;
    const int64_t sidx = k / strip, w = k - sidx * strip;
    const int64_t t = w / (GB_TILE_R * GB_TILE_C), l = w - t * (GB_TILE_R * GB_TILE_C);
    const int64_t dc = l / GB_TILE_R, dr = l - dc * GB_TILE_R;
    return sidx * strip + dc * P.tile_h + t * GB_TILE_R + dr;
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
        if (P.mu == P.mu) { // mu = NaN: the caller's v^t is already constrained (for its own mass) and is kept as given
            double g[5];
            metric_components_rt(P, o.x[1], o.x[2], g);
            o.v[0] = constrain_vt(g, o.v[1], o.v[2], o.v[3], P.mu);
        }
        return;
    }
    double alpha, beta;
    if (P.ic_kind == GB200_IC_IMPACT_PARAMETERS) { // explicit (alpha, beta) lists: ex[0] = alpha, ex[1] = beta
        alpha = P.ex[0][i];
        beta = P.ex[1][i];
    } else if (P.ic_kind == GB200_IC_CARTESIAN_PLANE) { // planes.jl:152-171: grids mirrored about their first point
        const int64_t rows = 2 * P.n1 - 1;
        const int64_t col = i / rows, row = i - col * rows;
        const int64_t ka = (col < P.n0 - 1) ? (P.n0 - 1 - col) : (col - (P.n0 - 1));
        const int64_t kb = (row < P.n1 - 1) ? (P.n1 - 1 - row) : (row - (P.n1 - 1));
        double xa, yb;
        if (P.grid_kind == GB200_GRID_GEOMETRIC) { xa = P.lo0 * pow(P.geoK, (double)ka); yb = P.lo1 * pow(P.geoK1, (double)kb); }
        else if (P.grid_kind == GB200_GRID_INVERSE) {
            xa = 1.0 / dd_axis(P.inv_lo_hi, P.inv_step_hi, P.inv_step_lo, P.n0 - 1 - ka);
            yb = 1.0 / dd_axis(P.inv1_lo_hi, P.inv1_step_hi, P.inv1_step_lo, P.n1 - 1 - kb);
        } else { xa = dd_axis(P.lo0, P.step0_hi, P.step0_lo, ka); yb = dd_axis(P.lo1, P.step1_hi, P.step1_lo, kb); }
        alpha = (col < P.n0 - 1) ? -xa : xa;
        beta = (row < P.n1 - 1) ? -yb : yb;
    } else if (P.ic_kind == GB200_IC_RENDER_GRID) { // rendering.jl:150-159
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
template <int METRIC>
GB_HD inline void circular_fourvelocity(const GbParams& P, double rho, double& ut_up, double& uph_up) {
    double g[5], dr[5], dth[5];
    metric_jacobian<METRIC>(P, rho, 1.0, 0.0, g, dr, dth);
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

#ifdef __CUDACC__
__host__ __device__ __noinline__
#endif
static double table_lerp(const double* xs, const double* ys, int n, double x) { // NaNLinearInterpolator, interpolations.jl:7-14 (clamped abscissa)
    x = fmin(fmax(x, xs[0]), xs[n - 1]);
    int lo = 0, hi = n - 1; // last index with xs[idx] <= x, clamped to [0, n-2]
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (xs[mid] <= x) lo = mid; else hi = mid; }
    const double w = (x - xs[lo]) / (xs[lo + 1] - xs[lo]);
    return (1.0 - w) * ys[lo] + w * ys[lo + 1];
}

// redshift_function (src/redshift.jl:192-220; plunging region :93-164; generic: :246-276)
// E_obs = g_{mu nu}(x_init) v_init^mu (1,0,0,0)^nu is a per-ray constant fixed at launch of the ray.
template <int METRIC>
GB_HD inline double redshift_endpoint(const GbParams& P, const double x[4], const double v[4], double E_obs) {
    const double sth = sin(x[2]);
    double rho = x[1] * fabs(sth);
    double u0, u1 = 0.0, u3;
    // A disc whose inner edge is the ISCO itself is hit at the edge to rounding (the event is the discontinuity of
    // distance_to_disc there): without a plunging table such a hit takes the circular orbit at the ISCO, where the
    // plunging flow starts, instead of having no velocity field at all.
    if (METRIC != GB200_METRIC_KERR && P.pl_n < 2 && rho < P.r_isco && rho >= P.r_isco * (1.0 - 1e-9)) rho = P.r_isco;
    if (rho < P.r_isco) {
        if (METRIC == GB200_METRIC_KERR) { // Cunningham (1975) plunging flow
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
        circular_fourvelocity<METRIC>(P, rho, u0, u3);
    }
    double g[5], dr[5], dth[5];
    metric_jacobian<METRIC>(P, x[1], sth, cos(x[2]), g, dr, dth);
    const double Ed = (g[0] * v[0] + g[4] * v[3]) * u0 + (g[1] * v[1]) * u1 + (g[4] * v[0] + g[3] * v[3]) * u3;
    return E_obs / Ed;
}

GB_HD inline double emissivity_eval(const GbParams& P, double rho) {
    if (P.emis_kind == GB200_EMISSIVITY_POWERLAW) return pow(rho, -P.emis_index);
    return table_lerp(P.emis_r, P.emis_eps, P.emis_n, rho);
}
